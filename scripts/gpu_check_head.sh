# full GPU check of the current build: parity suite, smoke, both bench arms
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v16.json 2> gpurun_out/bench_v16.err || { tail -5 gpurun_out/bench_v16.err; exit 1; }
tail -1 gpurun_out/bench_v16.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d['e2e'])[:900])"
