# round 2: 8 GPUs of one box -- the aggregate host<->device copy ceiling, then the bench (weak scaling of the C3 step, e2e, the one-job C4 leg)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
for n in 1 2 4 $N; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 scripts/pcie_probe_all.py 2>/dev/null | tail -1 | tee -a gpurun_out/r2_pcie_all.log
done
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/r2_lscpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err || { tail -20 gpurun_out/r2_bench_n$N.err; exit 1; }
tail -1 gpurun_out/r2_bench_n$N.json | cut -c1-300
