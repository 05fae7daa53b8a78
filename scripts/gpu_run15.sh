set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick15.log
for lib in "" _A _B _C; do
 export MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$lib.so
 echo "lib=$lib" | tee -a gpurun_out/quick15.log
 python scripts/quick_bench.py --config C3 --sites 16384 --rep 4 --classes --iters 5 --opt tile_bytes=16384 --opt ring_bytes=16384 2>&1 | tail -1 | tee -a gpurun_out/quick15.log
done
