# round 2, iteration 13: the driver's BCF paths on the device, e2e at two batch sizes (fill / drain share of the typed path), compute-sanitizer memcheck on a cross-section
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vcfcall.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_vcfcall13.log
{ echo "# e2e through mcb_call_host, BCF typed vectors both ways (int16 PL in; int8 GT, int8 GQ, int16 PL out), C3: 8,192 vs 32,768 sites per call";
  timeout 300 python scripts/e2e_typed_bench.py 8192 2>&1 | grep -v generated; timeout 600 python scripts/e2e_typed_bench.py 32768 2>&1 | grep -v generated;
  echo "# e2e with int32 buffers both ways, 8,192 vs 32,768 sites per call (slab 64 MB, ramp from 8 MB)";
  timeout 300 python scripts/e2e_bench.py 8192 2>&1 | grep '"min_mb": 8' | head -1; timeout 600 python scripts/e2e_bench.py 32768 2>&1 | grep '"min_mb": 8' | head -1; } | tee gpurun_out/r2_e2e_sizes.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vcfcall.py -m gpu -x -q -k "(biallelic_warp_kernel and 128-0) or (multi_allelic_kernel and 300-0) or (literal_phase1 and 7-5-0) or (sample_groups and 30-5) or more_than_five or mpileup.1 or cAls.7 or adjudicates" 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/r2_memcheck.log
