# round 2, iteration 5: pipelined mcall_multi.cu (two sites per CTA): parity under timeout first, then the sweep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb5.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_allelic or compacted or synthetic" 2>&1 | tail -8 | tee gpurun_out/r2_multi_tests.log
grep -q "passed" gpurun_out/r2_multi_tests.log || exit 1
grep -q "failed" gpurun_out/r2_multi_tests.log && exit 1
timeout 900 python scripts/quick_bench.py --config C3 --sites 16384 --classes --iters 5 --sweep ";mm_block=128,mm_k0=0;mm_block=128,mm_k0=4;mm_block=128,mm_k0=8;mm_block=128,mm_k0=12;mm_block=256,mm_k0=10;mm_block=256,mm_k0=40;mm_block=64,mm_k0=3;mm_block=64,mm_k0=0" 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['opts'], 'ms %.3f' % d['ms'], d['class_ms'])
" | tee -a gpurun_out/r2_qb5.log
timeout 600 python scripts/quick_bench.py --config C3 --sites 8192 --rep 4 --iters 3 > gpurun_out/r2_qb_pre_ncu.log 2>&1 || exit 1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mcall_multi_kernel" -s 6 -c 3 -f -o gpurun_out/r2_prof_multi python scripts/quick_bench.py --config C3 --sites 8192 --rep 4 --iters 3 > gpurun_out/r2_ncu_multi.log 2>&1
tail -1 gpurun_out/r2_ncu_multi.log | cut -c1-300
