"""Sweep CTA size / tile / ring of the tiled kernel per allele-count class on the C3 workload (development aid).
Prints the serialised device ms of the class for every combination that fits."""
import sys, os, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bcftools_b200 import abi, synth, mcall, device
classes = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "345")]
params, hb, tab = synth.make_batch("C3", 16384, with_groups=0)
db = device.DeviceBatch(hb, replicate=4)
dr = device.DeviceResult(db)
b, r = db.c_struct(), dr.c_struct()
stream = torch.cuda.current_stream().cuda_stream
for c in classes:
    best = None
    blocks = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (64, 128, 256)
    for block, tile_kb, ring_kb in itertools.product(blocks, (4, 8, 16, 32), (8, 16, 32, 64)):
        if ring_kb < tile_kb:
            continue
        opts = {"time_kernels": 1, "block_%d" % c: block, "tile_bytes_%d" % c: tile_kb << 10, "ring_bytes_%d" % c: ring_kb << 10}
        try:
            mc = mcall.MCaller(params, ploidy_tab=tab, options=opts)
            for _ in range(2):
                mc.call_device(b, r, stream)
            kt = []
            for _ in range(4):
                mc.call_device(b, r, stream)
                kt.append(mc.kernel_times_ms()[c])
            torch.cuda.synchronize()
            ms = float(np.median(kt))
            mc.close()
        except Exception as e:      # configuration does not fit
            print(json.dumps(dict(cls=c, block=block, tile_kb=tile_kb, ring_kb=ring_kb, error=str(e)[:60])), flush=True)
            continue
        print(json.dumps(dict(cls=c, block=block, tile_kb=tile_kb, ring_kb=ring_kb, ms=round(ms, 4))), flush=True)
        if best is None or ms < best[0]:
            best = (ms, block, tile_kb, ring_kb)
    print("BEST class %d: %.4f ms block=%d tile=%dK ring=%dK" % ((c,) + best), flush=True)
