set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick13.log
for lib in "" _u22_m4 _u12_m3 _u22_m3 _u22_m2; do
 export MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$lib.so
 echo "lib=$lib" | tee -a gpurun_out/quick13.log
 python scripts/quick_bench.py --config C2 --sites 8192 --rep 8 --check 2>&1 | tail -2 | tee -a gpurun_out/quick13.log
done
