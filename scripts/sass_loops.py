"""Static view of one kernel's SASS: instruction count, opcode mix and the loops (backward branches) with their lengths.
usage: python scripts/sass_loops.py <lib.so|obj.o> <kernel-name substring> [out.sass]"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
hits = [f for f in funcs[1:] if pat in f.split("\n")[0]]
for f in hits:
    name = f.split("\n")[0]
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(f"== {name[:110]}: {len(ins)} instructions")
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", t)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_idx:
                j = addr_idx[tgt]
                body = [x[1] for x in ins[j:i + 1]]
                ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", b).split()[0].split(".")[0] for b in body)
                top = ", ".join(f"{k}:{v}" for k, v in ops.most_common(12))
                print(f"   loop @{tgt:#06x}-{a:#06x}  len {i - j + 1:5d}   {top}")
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(f)
