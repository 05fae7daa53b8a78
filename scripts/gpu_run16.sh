set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/e2e_bench.py 8192 2>&1 | tail -6 | tee gpurun_out/e2e16.log
