# grouped calling (C5, 5 and 26 groups): two builds
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/groups_ab.log
for v in "" _bg1024; do
  echo "variant $v" | tee -a gpurun_out/groups_ab.log
  for args in "--config C5 --sites 2048 --rep 2 --groups 5 --classes" "--config C5 --sites 2048 --rep 2 --groups 26"; do
    MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so python scripts/quick_bench.py --iters 5 $args 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config'], round(d['ms'],4), '%.4g'%d['calls_per_s'], d.get('class_ms'))" | tee -a gpurun_out/groups_ab.log
  done
done
