# round 2, iteration 34: two-allele grouped kernel, more than 16 groups: the AD chunk staged once for both halves of the groups
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb34.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vcfcall.py -m gpu -x -q -k "two_allele_grouped or sample_groups or baseline_configs or call-G or goldens" 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_pytest_groups34.log
grep -q "failed" gpurun_out/r2_pytest_groups34.log && exit 1
for cfg in "5 2048 4" "26 2048 4"; do
  set -- $cfg
  echo "groups $1 sites $2 x rep $3" | tee -a gpurun_out/r2_qb34.log
  timeout 900 python scripts/quick_bench.py --config C5 --groups $1 --sites $2 --rep $3 --iters 5 --classes 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms %.3f calls/s %.3e launches %d' % (d['ms'], d['calls_per_s'], d['launches']), d.get('class_ms'))
" | tee -a gpurun_out/r2_qb34.log
done
