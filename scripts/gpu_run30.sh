set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "biallelic_warp or synthetic or adversarial or golden" 2>&1 | tail -8 | tee gpurun_out/pytest_bw.log
python scripts/quick_bench.py --iters 5 --config C3 --sites 16384 --rep 4 --classes 2>&1 | tail -1 | tee gpurun_out/quick31.log
ncu --set full --clock-control none --import-source on -k regex:biallelic -s 2 -c 1 -f -o gpurun_out/prof_bw_v3 python scripts/quick_bench.py --iters 3 --config C3 --sites 16384 --rep 4 > gpurun_out/ncu_bw.log 2>&1
