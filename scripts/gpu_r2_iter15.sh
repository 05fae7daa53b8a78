# round 2, iteration 15: grouped calling, CTA of 8 warps (5 groups in one round) vs 4 warps; batch-size dependence
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb15.log
for v in "" _v1 _v2; do
  for cfg in "5 2048 4" "26 2048 4" "5 8192 2"; do
    set -- $cfg
    echo "variant '$v' groups $1 sites $2 x rep $3" | tee -a gpurun_out/r2_qb15.log
    MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C5 --groups $1 --sites $2 --rep $3 --iters 5 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms %.3f calls/s %.3e' % (d['ms'], d['calls_per_s']))
" | tee -a gpurun_out/r2_qb15.log
  done
done
MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200_v1.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sample_groups or (baseline_configs and C5)" 2>&1 | tail -3 | tee gpurun_out/r2_pytest_groups15.log
