"""Dump per-instruction executed counts of one kernel from an .ncu-rep and summarise basic blocks."""
import csv, subprocess, sys, io
rep, kidx = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '0')
out = sys.argv[3] if len(sys.argv) > 3 else "/tmp/sass.txt"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
sections, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = dict(name=r[1], rows=[]); sections.append(cur); continue
    if r and r[0] == 'Address':
        cur['hdr'] = r; continue
    if cur is not None and 'hdr' in cur and len(r) >= len(cur['hdr']) - 2:
        cur['rows'].append(r)
sections = [x for x in sections if 'hdr' in x and x['rows']]
s = sections[int(kidx)] if kidx.isdigit() else [x for x in sections if kidx in x['name']][0]
idx = {h: i for i, h in enumerate(s['hdr'])}
lines = []
with open(out, "w") as fh:
    for k, r in enumerate(s['rows']):
        n, st = int(r[idx['Instructions Executed']]), int(r[idx['Warp Stall Sampling (All Samples)']])
        lines.append((k, n, st, r[idx['Source']].strip()))
        fh.write('%5d %10d %6d  %s\n' % lines[-1])
print(s['name'], 'total instr', sum(l[1] for l in lines), 'stall samples', sum(l[2] for l in lines))
blocks = []
for k, n, st, _ in lines:
    if blocks and blocks[-1][2] == n:
        blocks[-1][1] = k; blocks[-1][3] += st
    else:
        blocks.append([k, k, n, st])
tot = sum((b[1] - b[0] + 1) * b[2] for b in blocks)
big = sorted(blocks, key=lambda b: -(b[1] - b[0] + 1) * b[2])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]
for b in sorted(big):
    print('rows %5d-%5d  len %4d  count %9d  instr %11d (%4.1f%%) stall %d' % (b[0], b[1], b[1] - b[0] + 1, b[2], (b[1] - b[0] + 1) * b[2], 100 * (b[1] - b[0] + 1) * b[2] / tot, b[3]))
