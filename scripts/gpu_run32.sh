cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick33.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for cfg in ; do
  set -- $cfg
  for o in "warp2=-1" "warp2=0"; do
    echo "$1 $o" | tee -a gpurun_out/quick33.log
    python scripts/quick_bench.py --iters 5 --config $1 --sites $2 --rep $3 --opt $o 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['calls_per_s'])" | tee -a gpurun_out/quick33.log
  done
done
