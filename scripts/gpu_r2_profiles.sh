# round 2: the committed evidence of the bench command -- bench line (b200 + reference arm), ncu launch list, one ncu --set full
# capture of the class kernels (traffic JSON keyed by the kernel source hash), widened parity tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-final}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "baseline_configs or typed_end_to_end" 2>&1 | tail -5 | tee gpurun_out/r2_pytest_wide_$T.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_$T.json 2> gpurun_out/r2_bench_$T.err || { tail -5 gpurun_out/r2_bench_$T.err; exit 1; }
tail -1 gpurun_out/r2_bench_$T.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_$T.json 2> gpurun_out/r2_bench_ref_$T.err; tail -1 gpurun_out/r2_bench_ref_$T.json | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --sustain 0 > gpurun_out/r2_bench_short_$T.json 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$T.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --sustain 0 > gpurun_out/r2_ncu_launch_$T.log 2>&1
tail -1 gpurun_out/r2_ncu_launch_$T.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"mcall_multi_kernel|mcall_biallelic" -s 8 -c 4 -f -o gpurun_out/r2_prof_bench_$T python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --sustain 0 > gpurun_out/r2_ncu_full_$T.log 2>&1
tail -2 gpurun_out/r2_ncu_full_$T.log | cut -c1-200
