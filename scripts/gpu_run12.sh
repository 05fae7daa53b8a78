set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick12.log
Q="python scripts/quick_bench.py --config C2 --sites 8192 --rep 8"
$Q 2>&1 | tail -1 | tee -a gpurun_out/quick12.log
for tb in 4096 8192; do for rb in 8192 16384 24576; do
$Q --opt tile_bytes=$tb --opt ring_bytes_2=$rb 2>&1 | tail -1 | tee -a gpurun_out/quick12.log
done; done
Q="python scripts/quick_bench.py --config C3 --sites 8192 --rep 8"
$Q --opt tile_bytes=8192 --opt ring_bytes_2=16384 --opt ring_bytes_3=16384 --opt ring_bytes_4=16384 --opt ring_bytes_5=16384 2>&1 | tail -1 | tee -a gpurun_out/quick12.log
