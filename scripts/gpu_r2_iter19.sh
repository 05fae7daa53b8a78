cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mcall_biallelic_groups" -s 3 -c 1 -f -o gpurun_out/r2_prof_bgroups python scripts/quick_bench.py --config C5 --groups 5 --sites 8192 --rep 1 --iters 3 > gpurun_out/r2_ncu_bgroups.log 2>&1
tail -2 gpurun_out/r2_ncu_bgroups.log | cut -c1-200
