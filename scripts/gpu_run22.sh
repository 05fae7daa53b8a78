set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick22.log
Q="python scripts/quick_bench.py --iters 5"
$Q --config C1 --sites 100000 --rep 4 2>&1 | tail -1 | tee -a gpurun_out/quick22.log
$Q --config C2 --sites 16384 --rep 8 2>&1 | tail -1 | tee -a gpurun_out/quick22.log
$Q --config C3 --sites 16384 --rep 4 --classes 2>&1 | tail -1 | tee -a gpurun_out/quick22.log
$Q --config C3 --sites 16384 --rep 4 --tags 0 2>&1 | tail -1 | tee -a gpurun_out/quick22.log
$Q --config C3 --sites 16384 --rep 4 --flag 2 2>&1 | tail -1 | tee -a gpurun_out/quick22.log
$Q --config C4 --sites 256 --rep 8 2>&1 | tail -1 | tee -a gpurun_out/quick22.log
$Q --config C5 --sites 4096 --rep 2 2>&1 | tail -1 | tee -a gpurun_out/quick22.log
$Q --config C5 --sites 2048 --rep 2 --groups 5 2>&1 | tail -1 | tee -a gpurun_out/quick22.log
