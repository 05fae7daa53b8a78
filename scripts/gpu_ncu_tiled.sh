# ncu --set full captures of the tiled 3/4/5-allele kernels inside the bench command
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"mcall_site_kernel" -s 4 -c 4 -f -o gpurun_out/prof_tiled python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_tiled.log 2>&1
tail -3 gpurun_out/ncu_tiled.log
