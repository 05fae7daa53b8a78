# parity suite, then C5 pooled (mixed ploidy) and C3 device-resident timings
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python scripts/quick_bench.py --config C5 --sites 4096 --rep 2 --classes 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/qb_c5.json
python scripts/quick_bench.py --config C3 --sites 16384 --classes 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/qb_c3.json
