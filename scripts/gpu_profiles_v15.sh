# profiles of the bench command for profiles/: reference arm, launch list, one full capture of the dominant kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v15.json 2> gpurun_out/bench_ref_v15.err; tail -1 gpurun_out/bench_ref_v15.json | cut -c1-400
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v15.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"mcall_site_kernel|mcall_biallelic" -s 5 -c 5 -f -o gpurun_out/prof_bench_v15 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
