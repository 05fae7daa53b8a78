set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_b200.json
