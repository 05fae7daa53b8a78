# round 2, iteration 24: grouped calling, 3-5 allele classes: phase C without the full PL row (normalisers kept from phase A);
# variants: "" = 128-thread CTAs, _v1 = 160 threads (five warps: five big groups in one round), _v2 = 256 threads
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb24.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vcfcall.py -m gpu -x -q -k "adjudicates or two_allele_grouped or sample_groups or baseline_configs or more_than_five or hwe or call-G or af-fixation or goldens" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_pytest_groups24.log
grep -q "failed" gpurun_out/r2_pytest_groups24.log && exit 1
for v in "" _v1 _v2; do
for cfg in "5 2048 4" "26 2048 4" "5 8192 2"; do
  set -- $cfg
  echo "variant '$v' groups $1 sites $2 x rep $3" | tee -a gpurun_out/r2_qb24.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C5 --groups $1 --sites $2 --rep $3 --iters 5 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms %.3f calls/s %.3e launches %d' % (d['ms'], d['calls_per_s'], d['launches']))
" | tee -a gpurun_out/r2_qb24.log
done
done
