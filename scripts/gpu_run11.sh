set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick11.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
Q="python scripts/quick_bench.py"
$Q --config C4 --sites 96 --rep 4 2>&1 | tail -1 | tee -a gpurun_out/quick11.log
$Q --config C4 --sites 96 --rep 4 --block 256 2>&1 | tail -1 | tee -a gpurun_out/quick11.log
$Q --config C1 --sites 100000 --rep 4 2>&1 | tail -1 | tee -a gpurun_out/quick11.log
