set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "biallelic_warp" 2>&1 | tail -30 | tee gpurun_out/pytest_bw.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
python scripts/quick_bench.py --iters 5 --config C3 --sites 16384 --rep 4 --classes 2>&1 | tail -2 | tee gpurun_out/quick28.log
python scripts/quick_bench.py --iters 5 --config C3 --sites 16384 --rep 4 --classes --opt warp2=0 2>&1 | tail -2 | tee -a gpurun_out/quick28.log
