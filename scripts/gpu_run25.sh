set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err || exit 1
tail -1 gpurun_out/bench_v8.json
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
