# GPU parity suite on the box: /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash scripts/gpu_tests.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/smoke.log
