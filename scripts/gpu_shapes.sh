# throughput of the other workload shapes (development aid)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/shapes.log
run() { echo "$*" | tee -a gpurun_out/shapes.log; python scripts/quick_bench.py --iters 5 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms'],4), '%.4g'%d['calls_per_s'])" | tee -a gpurun_out/shapes.log; }
run --config C3 --sites 16384 --rep 4
run --config C5 --sites 4096 --rep 2
run --config C5 --sites 4096 --rep 2 --opt warp2=0
run --config C5 --sites 2048 --rep 2 --groups 5
run --config C2 --sites 65536 --rep 2
run --config C1 --sites 100000 --rep 1
run --config C4 --sites 256 --rep 2
