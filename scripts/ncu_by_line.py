"""Stall samples and executed instructions of one kernel of an .ncu-rep, aggregated by CUDA source line.
The SASS page of the report has no line column in CSV form, so the lines come from `nvdisasm -g` of the same object
(instruction order is the same).  usage: python scripts/ncu_by_line.py <rep> <kernel substring> <object.o> [top]"""
import csv, io, re, subprocess, sys, collections, tempfile, os, glob
rep, kpat, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
sections, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = dict(name=r[1], rows=[]); sections.append(cur); continue
    if r and r[0] == 'Address':
        cur['hdr'] = r; continue
    if cur is not None and 'hdr' in cur and len(r) >= len(cur['hdr']) - 2:
        cur['rows'].append(r)
s = [x for x in sections if re.sub(r'\((?:int|bool)\)|\s', '', kpat) in re.sub(r'\((?:int|bool)\)|\s', '', x['name']) and x.get('rows')][0]
idx = {h: i for i, h in enumerate(s['hdr'])}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
dis = "\n".join(subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout for c in sorted(glob.glob(os.path.join(tmp, "*.cubin"))))
# the kernel's section in the disassembly: mangled names differ from the demangled one; match by template digits
want = re.sub(r"[^0-9A-Za-z]", "", kpat)
lines, cur_line, active = [], 0, False
for l in dis.split("\n"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        norm = lambda t: re.sub(r"\((?:int|bool)\)|\s", "", t).replace("<false>", "<0>").replace("<true>", "<1>")
        active = norm(kpat) in norm(name)
        continue
    if not active:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur_line)
n = min(len(lines), len(s['rows']))
if abs(len(lines) - len(s['rows'])) > 0:
    sys.stderr.write("warning: %d disassembled instructions vs %d profiled rows\n" % (len(lines), len(s['rows'])))
agg = collections.defaultdict(lambda: collections.Counter())
cols = ['Warp Stall Sampling (All Samples)', 'Instructions Executed', 'stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_lg', 'stall_mio', 'stall_no_inst', 'stall_branch_resolving', 'stall_math']
for i in range(n):
    for c in cols:
        if c in idx:
            agg[lines[i]][c] += int(s['rows'][i][idx[c]] or 0)
tot = sum(v[cols[0]] for v in agg.values()); toti = sum(v[cols[1]] for v in agg.values())
print(s['name'][:90], "stall samples", tot, "instructions", toti)
srcs = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][cols[0]])[:top]:
    if f not in srcs:
        p = [q for q in glob.glob("/root/repo/bcftools_b200/csrc/**/" + f, recursive=True)]
        srcs[f] = open(p[0]).read().split("\n") if p else []
    text = srcs[f][ln-1].strip()[:90] if srcs[f] and ln-1 < len(srcs[f]) else ""
    det = " ".join("%s=%d" % (c.replace("stall_", ""), v[c]) for c in cols[2:] if v[c] > 0.05*max(v[cols[0]], 1))
    print("%5.1f%% stall %5.1f%% instr  %s:%d  [%s]  %s" % (100*v[cols[0]]/max(tot, 1), 100*v[cols[1]]/max(toti, 1), f, ln, det, text))
