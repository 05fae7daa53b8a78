# ncu --set full capture of the 3-allele tiled kernel inside the bench command
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"mcall_site_kernel" -s 6 -c 1 -f -o gpurun_out/prof_k3 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_k3.log 2>&1
tail -2 gpurun_out/ncu_k3.log
