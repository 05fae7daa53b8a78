set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick14.log
Q="python scripts/quick_bench.py --config C3 --sites 16384 --rep 4 --classes --iters 5"
$Q 2>&1 | tail -1 | tee -a gpurun_out/quick14.log
for rb in 8192 16384 32768; do for tb in 8192 16384; do
$Q --opt tile_bytes=$tb --opt ring_bytes=$rb 2>&1 | tail -1 | tee -a gpurun_out/quick14.log
done; done
$Q --block 256 2>&1 | tail -1 | tee -a gpurun_out/quick14.log
$Q --block 256 --opt tile_bytes=16384 --opt ring_bytes=16384 2>&1 | tail -1 | tee -a gpurun_out/quick14.log
