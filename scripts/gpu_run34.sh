cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out; rm -f gpurun_out/quick34.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "biallelic_warp or synthetic or adversarial or golden or compact" 2>&1 | tail -4 | tee gpurun_out/pytest_bw.log
for v in "" _grp; do
  echo "variant $v" | tee -a gpurun_out/quick34.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so python scripts/quick_bench.py --iters 5 --config C3 --sites 16384 --rep 4 --classes 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['class_ms'])" | tee -a gpurun_out/quick34.log
  MCALL_B200_LIB=$GRAFT_REPO_ROOT/bcftools_b200/lib/libmcall_b200$v.so python scripts/quick_bench.py --iters 5 --config C2 --sites 65536 --rep 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['calls_per_s'])" | tee -a gpurun_out/quick34.log
done
ncu --set full --clock-control none --import-source on -k regex:biallelic -s 2 -c 1 -f -o gpurun_out/prof_bw_v5 python scripts/quick_bench.py --iters 3 --config C3 --sites 16384 --rep 4 > gpurun_out/ncu_bw.log 2>&1
