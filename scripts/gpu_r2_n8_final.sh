# round 2, final: 8 GPUs of one box -- bench (weak scaling of the C3 step, e2e, the one-job C4 leg)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_n${N}_final.json 2> gpurun_out/r2_bench_n${N}_final.err || { tail -20 gpurun_out/r2_bench_n${N}_final.err; exit 1; }
tail -1 gpurun_out/r2_bench_n${N}_final.json | cut -c1-300
