# ncu --set full of the grouped-calling kernels on C5 (5 groups), source-level, after the plain run exited 0
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/quick_bench.py --config C5 --groups 5 --sites 2048 --rep 4 --iters 3 --classes 2>&1 | grep -v generated | tee gpurun_out/r2_qb_groups31.log || exit 1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"groups_kernel" -s 12 -c 6 -f -o gpurun_out/r2_prof_groups31 python scripts/quick_bench.py --config C5 --groups 5 --sites 2048 --rep 4 --iters 3 > gpurun_out/r2_ncu_groups31.log 2>&1
tail -2 gpurun_out/r2_ncu_groups31.log | cut -c1-200
