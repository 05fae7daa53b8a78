# round 2, iteration 23: two-allele kernel -- shared-memory base kept opaque / next site claimed early (A/B against the current build)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb23.log
for rep in 1 2; do
for v in "" _v1 _v2 _v3; do
  echo "variant '$v'" | tee -a gpurun_out/r2_qb23.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C3 --sites 65536 --classes --iters 8 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['opts'], 'ms %.3f' % d['ms'], d['class_ms'])
" | tee -a gpurun_out/r2_qb23.log
done
done
