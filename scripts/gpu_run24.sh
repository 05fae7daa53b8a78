set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick24.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
python scripts/quick_bench.py --iters 5 --config C5 --sites 2048 --rep 2 --groups 5 2>&1 | tail -1 | tee -a gpurun_out/quick24.log
python scripts/quick_bench.py --iters 5 --config C5 --sites 2048 --rep 2 --groups 26 2>&1 | tail -1 | tee -a gpurun_out/quick24.log
