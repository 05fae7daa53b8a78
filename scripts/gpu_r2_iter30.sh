# round 2, iteration 30: grouped calling with the phase-C screen: register caps again ("" = 6,4,2 CTAs/SM; _v1 = 5,4,2; _v2 = 4,4,2; _v3 = 5,3,2)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb30.log
for v in "" _v1 _v2 _v3; do
for cfg in "5 2048 4" "26 2048 4"; do
  set -- $cfg
  echo "variant '$v' groups $1 sites $2 x rep $3" | tee -a gpurun_out/r2_qb30.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C5 --groups $1 --sites $2 --rep $3 --iters 5 --classes 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms %.3f calls/s %.3e launches %d' % (d['ms'], d['calls_per_s'], d['launches']), d.get('class_ms'))
" | tee -a gpurun_out/r2_qb30.log
done
done
