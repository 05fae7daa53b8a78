set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick19.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
Q="python scripts/quick_bench.py --config C3 --sites 16384 --rep 4 --classes --iters 5"
$Q 2>&1 | tail -1 | tee -a gpurun_out/quick19.log
python scripts/quick_bench.py --config C2 --sites 8192 --rep 8 2>&1 | tail -1 | tee -a gpurun_out/quick19.log
Q="python scripts/quick_bench.py"
$Q --config C3 --sites 2048 --rep 4 --iters 2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mcall_site_kernel -s 16 -c 4 -o gpurun_out/prof_c3_v6 $Q --config C3 --sites 2048 --rep 4 --iters 2 > gpurun_out/ncu3.log 2>&1
