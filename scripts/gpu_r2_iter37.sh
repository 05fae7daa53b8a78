# round 2, iteration 37: grouped call, launch order of the class kernels (option gorder; default 54321)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb37.log
for g in 5 26; do
echo "groups $g" | tee -a gpurun_out/r2_qb37.log
timeout 900 python scripts/quick_bench.py --config C5 --groups $g --sites 2048 --rep 4 --iters 8 --sweep "gorder=54321;gorder=25431;gorder=23451;gorder=32451;gorder=34521;gorder=52431;gorder=35421;gorder=45321" 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['opts'], 'ms %.3f calls/s %.3e' % (d['ms'], d['calls_per_s']))
" | tee -a gpurun_out/r2_qb37.log
done
