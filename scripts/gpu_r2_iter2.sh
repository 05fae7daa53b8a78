# round 2, iteration 2: mcall_multi.cu v2 (small CTAs, sums in L2 scratch): parity of the kernel, (CTA size x ring depth) sweep, full suite
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_allelic or compacted" 2>&1 | tail -15 | tee gpurun_out/r2_multi_tests.log
SW=";multi=0"
for b in 64 128 256; do for n in 1 2 3; do SW="$SW;mm_block=$b,mm_nst=$n"; done; done
timeout 1200 python scripts/quick_bench.py --config C3 --sites 16384 --classes --iters 5 --sweep "$SW" 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['opts'], 'ms %.3f' % d['ms'], d['class_ms'])
" | tee gpurun_out/r2_qb2.log
for v in _v1; do
  if [ -f bcftools_b200/lib/libmcall_b200$v.so ]; then
      echo "variant $v" | tee -a gpurun_out/r2_qb2.log
      MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 600 python scripts/quick_bench.py --config C3 --sites 16384 --classes --iters 5 --sweep "mm_block=64,mm_nst=2;mm_block=128,mm_nst=2;mm_block=128,mm_nst=1;mm_block=256,mm_nst=1" 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['opts'], 'ms %.3f' % d['ms'], d['class_ms'])
" | tee -a gpurun_out/r2_qb2.log
  fi
done
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_pytest_gpu.log
