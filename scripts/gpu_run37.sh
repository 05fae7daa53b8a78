cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick37.log
for v in "" _nogp; do
  echo "variant $v" | tee -a gpurun_out/quick37.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so python scripts/quick_bench.py --iters 5 --config C3 --sites 16384 --rep 4 --classes 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['class_ms'])" | tee -a gpurun_out/quick37.log
done
