# round 2, iteration 36: two-allele kernel, phase-1 tiles through cp.async into the unused tail of the packed-copy buffer ("" = on, _v1 = off)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb36.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_pytest36.log
grep -q "failed\|error" gpurun_out/r2_pytest36.log && exit 1
for rep in 1 2; do
for v in "" _v1; do
  echo "variant '$v'" | tee -a gpurun_out/r2_qb36.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C3 --sites 16384 --rep 4 --classes --iters 8 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['opts'], 'ms %.3f' % d['ms'], d['class_ms'])
" | tee -a gpurun_out/r2_qb36.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C2 --sites 8192 --rep 4 --classes --iters 8 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('C2', 'ms %.3f' % d['ms'], d['class_ms'])
" | tee -a gpurun_out/r2_qb36.log
done
done
