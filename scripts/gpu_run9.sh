set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/quick9.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
Q="python scripts/quick_bench.py --config C3 --sites 8192 --rep 8"
$Q --opt concurrent=0 2>&1 | tail -1 | tee -a gpurun_out/quick9.log
$Q --opt concurrent=1 2>&1 | tail -1 | tee -a gpurun_out/quick9.log
for rb in 16384 32768; do
$Q --opt concurrent=1 --opt ring_bytes_3=$rb --opt ring_bytes_4=$rb --opt ring_bytes_5=$rb 2>&1 | tail -1 | tee -a gpurun_out/quick9.log
$Q --opt concurrent=1 --opt ring_bytes_4=$rb --opt ring_bytes_5=$rb 2>&1 | tail -1 | tee -a gpurun_out/quick9.log
done
$Q --opt concurrent=1 --opt ring_bytes_4=16384 --opt ring_bytes_5=16384 --opt block_4=256 --opt block_5=256 2>&1 | tail -1 | tee -a gpurun_out/quick9.log
$Q --opt concurrent=1 --opt tile_bytes=16384 2>&1 | tail -1 | tee -a gpurun_out/quick9.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench9.json
