# A/B of build variants (bcftools_b200/build.py build_variant): bash scripts/gpu_variants.sh "" _w13 _w14
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
for v in "$@"; do
  echo "variant $v" | tee -a gpurun_out/variants.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so python scripts/quick_bench.py --iters 7 --config C3 --sites 16384 --rep 4 --classes 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['class_ms'])" | tee -a gpurun_out/variants.log
done
