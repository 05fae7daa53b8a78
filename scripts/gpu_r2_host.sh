# round 2: host-side pieces on the box -- async batcher, one job over several contexts, the whole GPU suite, a short bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "async or ploidy_vector_once or job_over" 2>&1 | tail -8 | tee gpurun_out/r2_host_tests.log
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 3000 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
