# round 2, iteration 14: grouped calling with one warp per big group (no block barriers in phase A); register-cap variants
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb14.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vcfcall.py -m gpu -x -q -k "sample_groups or baseline_configs or more_than_five or hwe or call-G or af-fixation or goldens" 2>&1 | tail -5 | tee gpurun_out/r2_pytest_groups14.log
grep -q "passed" gpurun_out/r2_pytest_groups14.log || exit 1
grep -q "failed" gpurun_out/r2_pytest_groups14.log && exit 1
for v in "" _v1 _v2; do
  for ng in 5 26; do
    echo "variant '$v' groups $ng" | tee -a gpurun_out/r2_qb14.log
    MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 600 python scripts/quick_bench.py --config C5 --groups $ng --sites 2048 --rep 4 --iters 5 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms %.3f calls/s %.3e' % (d['ms'], d['calls_per_s']))
" | tee -a gpurun_out/r2_qb14.log
  done
done
