"""Device-resident C3 step under different per-class residency / launch-order options (development aid).
usage: mix_sweep.py '<json list of option dicts>' [sites] — prints the median step ms per option set."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bcftools_b200 import abi, synth, mcall, device
sets = json.loads(sys.argv[1])
nsites = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
params, hb, tab = synth.make_batch("C3", nsites, with_groups=0)
db = device.DeviceBatch(hb, replicate=4)
dr = device.DeviceResult(db)
b, r = db.c_struct(), dr.c_struct()
stream = torch.cuda.current_stream().cuda_stream
calls = db.nsites * params.nsmpl
for opts in sets:
    try:
        mc = mcall.MCaller(params, ploidy_tab=tab, options=opts)
        for _ in range(3):
            mc.call_device(b, r, stream)
        torch.cuda.synchronize()
        ts = []
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mc.call_device(b, r, stream); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        mc.close()
        ms = float(np.median(ts))
        print(json.dumps(dict(opts=opts, ms=round(ms, 4), min_ms=round(min(ts), 4), calls_per_s=calls / ms * 1e3)), flush=True)
    except Exception as e:
        print(json.dumps(dict(opts=opts, error=str(e)[:80])), flush=True)
