# development iteration on the box: parity suite, then device-resident timings with per-class breakdown
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python scripts/quick_bench.py --config C3 --sites 16384 --classes 2>&1 | tail -1 | tee gpurun_out/qb_c3.json
python scripts/quick_bench.py --config C2 --sites 16384 --classes 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/qb_c2.json
