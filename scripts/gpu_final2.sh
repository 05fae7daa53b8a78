cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_short.json 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"mcall_site_kernel|mcall_biallelic" -s 5 -c 5 -f -o gpurun_out/prof_bench_v13 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
