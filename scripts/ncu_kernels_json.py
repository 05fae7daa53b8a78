"""profiles/r02_ncu_kernels.json from an `ncu --set full` capture of the bench command: DRAM bytes per launch of every class kernel.
usage: python scripts/ncu_kernels_json.py <capture.ncu-rep> <sites_per_step> <kernel_source_hash> <out.json>
bench.py quotes roofline.traffic from this file only while the hash matches the kernel sources it runs."""
import csv, io, json, re, subprocess, sys
rep, sites, src_hash, out = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v, u = float(r[idx[name]].replace(",", "")), units[idx[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(u, 1)
classes = {}
for r in data:
    name = r[idx["Kernel Name"]]
    m = re.search(r"mcall_multi_kernel<\(?(?:int\))?(\d)", name)
    k = 2 if "biallelic" in name else (int(m.group(1)) if m else None)
    if k is None or k in classes:
        continue
    classes[k] = dict(kernel=name[:60], dram_bytes_read=val(r, "dram__bytes_read.sum"), dram_bytes_write=val(r, "dram__bytes_write.sum"),
                      duration_s_under_ncu=val(r, "gpu__time_duration.sum"), registers=int(float(r[idx["launch__registers_per_thread"]])),
                      issue_active_pct=float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]))
json.dump(dict(sites_per_step=sites, kernel_source_hash=src_hash, source="ncu --set full --clock-control none of `python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e` (%s)" % rep.split("/")[-1],
               classes={str(k): v for k, v in sorted(classes.items())}), open(out, "w"), indent=1)
print(open(out).read())
