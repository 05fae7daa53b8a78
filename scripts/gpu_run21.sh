set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick21.log
Q="python scripts/quick_bench.py --config C3 --sites 16384 --rep 4 --classes --iters 5"
for b in 32 64; do
 python scripts/quick_bench.py --config C3 --sites 1024 --rep 1 --iters 2 --block $b --check 2>&1 | tail -1 | tee -a gpurun_out/quick21.log
 for tb in 4096 8192; do for rb in 8192 16384 32768; do
  $Q --block $b --opt tile_bytes=$tb --opt ring_bytes=$rb 2>&1 | tail -1 | tee -a gpurun_out/quick21.log
 done; done
done
for b in 32 64 128; do
 python scripts/quick_bench.py --config C2 --sites 16384 --rep 4 --iters 5 --block $b --opt tile_bytes=4096 --opt ring_bytes=16384 2>&1 | tail -1 | tee -a gpurun_out/quick21.log
done
python scripts/quick_bench.py --config C1 --sites 100000 --rep 4 --iters 5 --block 32 --opt tile_bytes=4096 --opt ring_bytes=4096 2>&1 | tail -1 | tee -a gpurun_out/quick21.log
