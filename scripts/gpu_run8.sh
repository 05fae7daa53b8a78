set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
Q="python scripts/quick_bench.py"
for b in 128 256; do
  $Q --config C2 --sites 4096 --rep 8 --block $b 2>&1 | tail -1 | tee -a gpurun_out/quick8.log
  $Q --config C3 --sites 4096 --rep 8 --block $b 2>&1 | tail -1 | tee -a gpurun_out/quick8.log
done
$Q --config C3 --sites 2048 --rep 4 --iters 2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mcall_site_kernel -s 16 -c 4 -o gpurun_out/prof_c3_v5 $Q --config C3 --sites 2048 --rep 4 --iters 2 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
