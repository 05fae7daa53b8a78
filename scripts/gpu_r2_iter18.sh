# launch list of a grouped C5 step: per-kernel durations (ncu serialises the launches)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/quick_bench.py --config C5 --groups 5 --sites 8192 --rep 1 --iters 3 > gpurun_out/r2_qb18_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 40 -c 9 --csv --log-file gpurun_out/r2_launches_groups.csv python scripts/quick_bench.py --config C5 --groups 5 --sites 8192 --rep 1 --iters 3 > gpurun_out/r2_ncu_launch_groups.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_groups.csv")) if len(r) > 10 and r[0] != "ID"]
agg = {}
for r in rows:
    agg.setdefault((r[0], r[4][:60]), {})[r[-3] if False else r[12]] = r[-1]
for k, v in agg.items(): print(k, v)
PY
