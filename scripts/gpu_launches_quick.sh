# launch list (per-kernel durations) of one quick C3 run
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_quick.csv python scripts/quick_bench.py --config C3 --sites 16384 --iters 2 > gpurun_out/ncu_quick.log 2>&1
tail -2 gpurun_out/ncu_quick.log | cut -c1-200
