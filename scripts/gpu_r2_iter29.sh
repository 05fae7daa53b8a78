# round 2, iteration 29: grouped calling, float32 screen in phase C of mcall_groups.cu ("" = on, _v1 = off)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb29.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vcfcall.py -m gpu -x -q -k "adjudicates or two_allele_grouped or sample_groups or baseline_configs or more_than_five or hwe or call-G or af-fixation or goldens" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_pytest_groups29.log
grep -q "failed" gpurun_out/r2_pytest_groups29.log && exit 1
for v in "" _v1; do
for cfg in "5 2048 4" "26 2048 4" "5 8192 2"; do
  set -- $cfg
  echo "variant '$v' groups $1 sites $2 x rep $3" | tee -a gpurun_out/r2_qb29.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C5 --groups $1 --sites $2 --rep $3 --iters 5 --classes 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms %.3f calls/s %.3e launches %d' % (d['ms'], d['calls_per_s'], d['launches']), d.get('class_ms'))
" | tee -a gpurun_out/r2_qb29.log
done
done
