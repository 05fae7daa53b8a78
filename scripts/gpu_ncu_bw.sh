# one ncu --set full capture of the two-allele warp kernel inside the bench command (after a plain run exits 0)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_short.json 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"mcall_biallelic" -s 2 -c 1 -f -o gpurun_out/prof_bw python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
