# round 2, iteration 9: phase 1 of mcall_multi.cu as an FP64 tensor-path matrix product; register-cap variants
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb9.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_pytest_gpu9.log
grep -q "passed" gpurun_out/r2_pytest_gpu9.log || exit 1
grep -q "failed" gpurun_out/r2_pytest_gpu9.log && exit 1
for v in "" _v1 _v2; do
  echo "variant '$v'" | tee -a gpurun_out/r2_qb9.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C3 --sites 16384 --classes --iters 5 --sweep ";mm_nst=1;mm_block=256;mm_block=256,mm_nst=1;mm_block=64;mm_block=64,mm_nst=1" 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['opts'], 'ms %.3f' % d['ms'], d['class_ms'])
" | tee -a gpurun_out/r2_qb9.log
done
