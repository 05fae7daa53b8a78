"""One C4-shaped job (100,000 samples) over 1 / 2 / 4 / 8 GPUs of the box through mcb_job_call_host (include/mcall_job.h):
ONE process, one host thread per device, pinned host buffers, contiguous site ranges, results in input order.  Strong
scaling of the end-to-end path; prints one JSON line.  usage: python scripts/job_bench.py [--sites 1536] [--typed]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bcftools_b200 import abi, mcall, synth

ap = argparse.ArgumentParser()
ap.add_argument("--sites", type=int, default=1536)
ap.add_argument("--config", default="C4")
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
ngpu = torch.cuda.device_count()
t0 = time.time()
params, hb, tab = synth.make_batch(args.config, args.sites)
hb = mcall.pin_batch(hb)
res = mcall.pin_result(abi.HostResult(hb, compact=True))
out = dict(workload=args.config, nsmpl=params.nsmpl, sites=hb.nsites, h2d_bytes=int(hb.pl.nbytes), gen_s=round(time.time() - t0, 1),
           scaling="strong: one job, contiguous site ranges balanced by PL volume, ordered concatenation", unit="calls/s")
ref = None
for ndev in [n for n in (1, 2, 4, 8) if n <= ngpu]:
    with mcall.MJob(params, list(range(ndev))) as jb:
        jb.call_host(hb, res)
        ts = []
        for _ in range(args.iters):
            t = time.perf_counter(); r = jb.call_host(hb, res); ts.append(time.perf_counter() - t)
        dt = float(np.median(ts))
        sig = (int(np.asarray(r.ret).sum()), int(np.asarray(r.gt).sum(dtype=np.int64)), int(np.asarray(r.gq).sum(dtype=np.int64)), np.asarray(r.qual).tobytes())
    if ref is None: ref = sig
    out["devices_%d" % ndev] = dict(value=hb.nsites * params.nsmpl / dt, ms_per_call=round(1e3 * dt, 2), first_site=[int(x) for x in r.first_site],
                                    same_result_as_1=bool(sig == ref))
out["speedup"] = {k: round(v["value"] / out["devices_1"]["value"], 3) for k, v in out.items() if k.startswith("devices_")}
print(json.dumps(out))
