set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
Q="python scripts/quick_bench.py"
# bigger batch numbers
$Q --config C2 --sites 4096 --rep 8 2>&1 | tail -1 | tee -a gpurun_out/quick2.log
$Q --config C3 --sites 4096 --rep 8 2>&1 | tail -1 | tee -a gpurun_out/quick2.log
for bps in 1 2 3; do $Q --config C2 --sites 4096 --rep 8 --opt blocks_per_sm=$bps 2>&1 | tail -1 | tee -a gpurun_out/quick2.log; done
$Q --config C2 --sites 4096 --rep 8 --opt tile_bytes=4096 --opt ring_bytes=8192 2>&1 | tail -1 | tee -a gpurun_out/quick2.log
$Q --config C2 --sites 4096 --rep 8 --tags 0 2>&1 | tail -1 | tee -a gpurun_out/quick2.log
# launch list
$Q --config C3 --sites 2048 --rep 4 --iters 2 > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c3.csv $Q --config C3 --sites 2048 --rep 4 --iters 2 > gpurun_out/ncu1.log 2>&1
# full capture of the biallelic kernel
$Q --config C2 --sites 4096 --rep 4 --iters 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mcall_site_kernelILi2 -s 3 -c 1 -o gpurun_out/prof_c2 $Q --config C2 --sites 4096 --rep 4 --iters 2 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
