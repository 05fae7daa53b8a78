cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || { tail -20 gpurun_out/bench_n$N.err; exit 1; }
tail -1 gpurun_out/bench_n$N.json | cut -c1-400
