set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/quick_bench.py --iters 3 --config C3 --sites 16384 --rep 4 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:biallelic -s 2 -c 1 -f -o gpurun_out/prof_bw_v1 python scripts/quick_bench.py --iters 3 --config C3 --sites 16384 --rep 4 > gpurun_out/ncu_bw.log 2>&1
tail -2 gpurun_out/ncu_bw.log
