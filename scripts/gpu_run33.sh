cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err || { tail -5 gpurun_out/bench_v9.err; exit 1; }
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v9.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['all_kernels']['ms_per_class'])
print(d['e2e'])
PY
python scripts/e2e_bench.py 8192 2>&1 | tee gpurun_out/e2e33.log
