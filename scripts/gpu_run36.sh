cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick36.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python scripts/quick_bench.py --iters 5 --config C3 --sites 16384 --rep 4 --classes 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['class_ms'])" | tee -a gpurun_out/quick36.log
