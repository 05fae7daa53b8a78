# A/B of runtime options: bash scripts/gpu_opts.sh "concurrent=0" "concurrent=2"
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/opts.log
for o in "$@"; do
  echo "opts $o" | tee -a gpurun_out/opts.log
  python scripts/quick_bench.py --iters 7 --config C3 --sites 16384 --rep 4 $(for kv in ${o//,/ }; do echo --opt $kv; done) 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['tmin_ms'])" | tee -a gpurun_out/opts.log
done
