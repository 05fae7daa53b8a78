# round 2, first look at mcall_multi.cu on the box: its parity tests first (fail fast, everything under timeout), then timings
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_allelic" 2>&1 | tail -15 | tee gpurun_out/r2_multi_tests.log
timeout 900 python scripts/quick_bench.py --config C3 --sites 16384 --classes --iters 7 --sweep ";multi=0;mm_nst=1;mm_nst=2;mm_nst=3;mm_nst=4;bps_3=1,bps_4=1" 2>&1 | grep -v generated | cut -c1-800 | tee gpurun_out/r2_qb.log
for v in _v1; do
  if [ -f bcftools_b200/lib/libmcall_b200$v.so ]; then
      echo "variant $v" | tee -a gpurun_out/r2_qb.log
      MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 600 python scripts/quick_bench.py --config C3 --sites 16384 --classes --iters 7 --sweep ";mm_nst=2;mm_nst_5=1" 2>&1 | grep -v generated | cut -c1-800 | tee -a gpurun_out/r2_qb.log
  fi
done
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_pytest_gpu.log
