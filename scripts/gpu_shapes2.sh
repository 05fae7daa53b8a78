cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/shapes2.log
run() { echo "$*" | tee -a gpurun_out/shapes2.log; python scripts/quick_bench.py --iters 5 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms'],4), '%.4g'%d['calls_per_s'], d.get('class_ms'))" | tee -a gpurun_out/shapes2.log; }
run --config C5 --sites 16384 --rep 4 --classes
run --config C5 --sites 16384 --rep 4 --classes --opt warp2=0
