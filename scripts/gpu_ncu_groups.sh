# ncu --set full capture of the grouped-calling kernels on C5 (5 groups)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"mcall_groups_kernel" -s 10 -c 5 -f -o gpurun_out/prof_groups python scripts/quick_bench.py --config C5 --sites 2048 --rep 2 --groups 5 --iters 2 > gpurun_out/ncu_groups.log 2>&1
tail -2 gpurun_out/ncu_groups.log | cut -c1-300
