"""Stall / instruction breakdown of one kernel of an .ncu-rep by source-line ranges of a .cu file (needs -lineinfo)."""
import csv, subprocess, io, sys, re, collections
rep, kidx = sys.argv[1], int(sys.argv[2])
marks = [m.split(':') for m in sys.argv[3:]]        # name:first_line ... (sorted), ranges are [line_i, line_{i+1})
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
s = sections[kidx]
idx = {h: i for i, h in enumerate(s['hdr'])}
cols = ['Instructions Executed', 'Warp Stall Sampling (All Samples)', 'stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_no_inst', 'stall_barrier', 'stall_branch_resolving', 'stall_not_selected', 'stall_math']
tot = collections.Counter()
for r in s['rows']:
    for c in cols:
        if c in idx: tot[c] += int(r[idx[c]] or 0)
print(s['name'][:80])
print(' '.join('%s=%d' % (c.replace('Warp Stall Sampling (All Samples)', 'stall_all').replace('Instructions Executed', 'instr'), tot[c]) for c in cols))
