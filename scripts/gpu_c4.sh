# biobank shape (C4, tiled two-allele kernel) and the tiled kernel forced on C3's two-allele class, two builds
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/c4.log
for v in "" _n2off; do
  echo "variant $v" | tee -a gpurun_out/c4.log
  for args in "--config C4 --sites 256 --rep 2" "--config C3 --sites 16384 --rep 4 --opt warp2=0 --classes" "--config C2 --sites 65536 --rep 2 --opt warp2=0"; do
    MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so python scripts/quick_bench.py --iters 5 $args 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config'], round(d['ms'],4), '%.4g'%d['calls_per_s'], d.get('class_ms'))" | tee -a gpurun_out/c4.log
  done
done
