"""Summarise an .ncu-rep (read on the CPU box): per-kernel headline metrics + top SASS blocks by executed instructions."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct',
        'smsp__cycles_active.avg', 'sm__cycles_active.avg']
print("kernel", [r[idx['Kernel Name']][:40] for r in data])
for w in want:
    if w in idx:
        print("%-62s %-14s" % (w, units[idx[w]]), [r[idx[w]][:12] for r in data])
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for h in stall:
    vals = [float(r[idx[h]] or 0) for r in data]
    if max(vals) > 0.3:
        print("%-62s" % h.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', ''), ["%.2f" % v for v in vals])
