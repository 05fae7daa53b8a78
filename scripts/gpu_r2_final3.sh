# round 2, final: the whole GPU suite, smoke(), then the bench lines (both arms) with the traffic JSON of the current kernel sources
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2_pytest_gpu_final3.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2_smoke_final3.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final3.json 2> gpurun_out/r2_bench_ref_final3.err; tail -1 gpurun_out/r2_bench_ref_final3.json | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final3.json 2> gpurun_out/r2_bench_final3.err || { tail -5 gpurun_out/r2_bench_final3.err; exit 1; }
tail -1 gpurun_out/r2_bench_final3.json | cut -c1-300
