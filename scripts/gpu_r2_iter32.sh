# round 2, iteration 32: new grouped tests, then the evidence capture of the bench command for the current kernel sources
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sample_groups or per_class_times" 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2_pytest_groups32.log
grep -q "failed\|error" gpurun_out/r2_pytest_groups32.log && exit 1
bash scripts/gpu_r2_profiles.sh iter32
