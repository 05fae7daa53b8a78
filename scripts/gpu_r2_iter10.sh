# round 2, iteration 10: the htslib-free `call -m` driver end to end on the device
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vcfcall.py -m gpu -q 2>&1 | tail -40 | cut -c1-600 | tee gpurun_out/r2_pytest_vcfcall.log
