# ncu --set full capture of the class kernels of one C3 step (after the un-profiled command exited 0)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/quick_bench.py --config C3 --sites 8192 --rep 4 --iters 3 > gpurun_out/r2_qb_pre_ncu.log 2>&1 || exit 1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"mcall_multi_kernel|mcall_biallelic" -s 8 -c 4 -f -o gpurun_out/r2_prof_iter8 python scripts/quick_bench.py --config C3 --sites 8192 --rep 4 --iters 3 > gpurun_out/r2_ncu_iter8.log 2>&1
tail -2 gpurun_out/r2_ncu_iter8.log | cut -c1-300
