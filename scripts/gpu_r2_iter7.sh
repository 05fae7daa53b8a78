# round 2, iteration 7: float32 screen of phase 2 (two-allele kernel + pair/triple paths of mcall_multi.cu)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/screen_selftest.py 2>&1 | tee gpurun_out/r2_screen_selftest.log | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_pytest_gpu7.log
timeout 600 python scripts/quick_bench.py --config C3 --sites 16384 --classes --iters 5 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['opts'], 'ms %.3f' % d['ms'], d['class_ms'])
" | tee gpurun_out/r2_qb7.log
