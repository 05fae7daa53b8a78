"""End-to-end timing of mcb_call_host with BCF typed vectors both ways (int16 PL in, int8/int8/int16 out) for different
slab sizes (development aid)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bcftools_b200 import abi, synth, mcall
sites = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
params, hb, tab = synth.make_batch("C3", sites, with_groups=0)
sub = mcall.pin_batch(hb.subset(range(sites)).to_int16())
extra = {k: int(v) for k, v in (o.split("=") for o in sys.argv[2:])}
for slab_mb, min_mb in ((64, 8), (32, 4)):
    mc = mcall.MCaller(params, ploidy_tab=tab, options=dict({"slab_bytes": slab_mb << 20, "slab_min": min_mb << 20}, **extra))
    res = mcall.pin_result(abi.HostResult(sub, compact=True, typed=True))
    for _ in range(2):
        mc.call_host(sub, res)
    t0 = time.perf_counter()
    n = 8
    for _ in range(n):
        mc.call_host(sub, res)
    dt = (time.perf_counter() - t0) / n
    print(json.dumps(dict(slab_mb=slab_mb, min_mb=min_mb, ms=round(dt * 1e3, 3), calls_per_s=sites * params.nsmpl / dt)), flush=True)
    mc.close()
