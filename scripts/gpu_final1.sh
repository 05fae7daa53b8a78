cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v13.json 2> gpurun_out/bench_ref_v13.err; tail -1 gpurun_out/bench_ref_v13.json | cut -c1-600
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err || { tail -5 gpurun_out/bench_v13.err; exit 1; }
tail -1 gpurun_out/bench_v13.json | cut -c1-300
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v13.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
