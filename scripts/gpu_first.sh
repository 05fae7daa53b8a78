set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc; free -g | head -2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
for cfg in C2 C3; do
  timeout 600 python scripts/quick_bench.py --config $cfg --sites 2048 --rep 4 --check 2>&1 | tail -5 | tee -a gpurun_out/quick.log
done
