# round 2, iteration 25: grouped calling: per-class durations (classes serialised) beside the concurrent step, and instruction counts
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb25.log
for v in "" _v1 _v2; do
for cfg in "5 2048 4" "26 2048 4"; do
  set -- $cfg
  echo "variant '$v' groups $1 sites $2 x rep $3" | tee -a gpurun_out/r2_qb25.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C5 --groups $1 --sites $2 --rep $3 --iters 5 --classes 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms %.3f calls/s %.3e launches %d' % (d['ms'], d['calls_per_s'], d['launches']), d.get('class_ms'))
" | tee -a gpurun_out/r2_qb25.log
done
done
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread --clock-control none -k regex:"groups" -s 16 -c 12 --csv --log-file gpurun_out/r2_launches_groups25.csv python scripts/quick_bench.py --config C5 --groups 5 --sites 2048 --rep 4 --iters 3 > gpurun_out/r2_ncu25.log 2>&1
tail -3 gpurun_out/r2_ncu25.log | cut -c1-200
