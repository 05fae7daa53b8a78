set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
Q="python scripts/quick_bench.py"
$Q --config C3 --sites 2048 --rep 4 --iters 2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mcall_site_kernel -s 16 -c 4 -o gpurun_out/prof_c3 $Q --config C3 --sites 2048 --rep 4 --iters 2 > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
ls -la gpurun_out
