# round 2, iteration 6: class-3 variants with the coefficients in shared memory (fewer registers, more resident warps)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb6.log
for v in "" _v1 _v2; do
  echo "variant '$v'" | tee -a gpurun_out/r2_qb6.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C3 --sites 16384 --classes --iters 5 --sweep ";mm_block_3=64;mm_block_3=128,mm_nst_3=1;mm_block_3=256" 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['opts'], 'ms %.3f' % d['ms'], d['class_ms'])
" | tee -a gpurun_out/r2_qb6.log
done
