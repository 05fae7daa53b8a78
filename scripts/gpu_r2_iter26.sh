# round 2, iteration 26: grouped calling, CTA size and per-class register caps of mcall_groups.cu
# "" 128 threads (4,4,4 CTAs/SM for <=3 / 4 / 5 alleles); A 128 (6,4,4); B 128 (5,5,4); C 160 (5,4,3); D 160 (4,3,3); E 160 (3,3,3); F 256 (3,2,2)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/r2_qb26.log
for v in "" _vA _vB _vC _vD _vE _vF; do
for cfg in "5 2048 4" "26 2048 4"; do
  set -- $cfg
  echo "variant '$v' groups $1 sites $2 x rep $3" | tee -a gpurun_out/r2_qb26.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so timeout 900 python scripts/quick_bench.py --config C5 --groups $1 --sites $2 --rep $3 --iters 5 --classes 2>&1 | grep -v generated | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('ms %.3f calls/s %.3e launches %d' % (d['ms'], d['calls_per_s'], d['launches']), d.get('class_ms'))
" | tee -a gpurun_out/r2_qb26.log
done
done
