"""Quick device-resident timing of the fused kernel on one config (development aid, not bench.py)."""
import argparse, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bcftools_b200 import abi, synth, mcall, device

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--sites", type=int, default=2048)
ap.add_argument("--rep", type=int, default=4)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--tags", type=int, default=abi.CALL_FMT_GQ)
ap.add_argument("--flag", type=int, default=0)
ap.add_argument("--groups", type=int, default=0, help="-G groups (0 = pooled)")
ap.add_argument("--check", action="store_true")
ap.add_argument("--classes", action="store_true", help="also print the device ms per allele-count class (serialised)")
ap.add_argument("--block", type=int, default=0)
ap.add_argument("--sweep", default="", help="option sets separated by ';' (each 'k=v,k=v'; empty = defaults): one timing line per set on the same data")
args = ap.parse_args()

t0 = time.time()
params, hb, tab = synth.make_batch(args.config, args.sites, flag=args.flag, output_tags=args.tags, with_groups=args.groups)
print("generated", args.config, hb.nsites, "sites in %.1fs" % (time.time() - t0), flush=True)
base_opts = {k: int(v) for k, v in (o.split("=") for o in args.opt)}
if args.block: base_opts["block"] = args.block
db = device.DeviceBatch(hb, replicate=args.rep)
dr = device.DeviceResult(db)
b, r = db.c_struct(), dr.c_struct()
stream = torch.cuda.current_stream().cuda_stream
sets = [x for x in args.sweep.split(";")] if args.sweep else [""]
for sweep_set in sets:
    opts = dict(base_opts)
    opts.update({k: int(v) for k, v in (o.split("=") for o in sweep_set.split(",") if o)})
    mc = mcall.MCaller(params, ploidy_tab=tab, options=opts)
    for _ in range(3):
        mc.call_device(b, r, stream)
    torch.cuda.synchronize()
    times = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mc.call_device(b, r, stream); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    res = dr.to_host()
    rd, wr = synth.algorithmic_bytes(hb, res, params.output_tags)
    calls = db.nsites * params.nsmpl
    t = float(np.median(times))
    out = dict(config=args.config, sites=db.nsites, nsmpl=params.nsmpl, opts=opts, ms=t * 1e3, calls_per_s=calls / t,
               alg_GBs=(rd + wr) * args.rep / t / 1e9, bytes_per_call=(rd + wr) / (hb.nsites * params.nsmpl),
               frac_of_6551=(rd + wr) * args.rep / t / 1e9 / 6551.4, tmin_ms=min(times) * 1e3, launches=int(mc.stats()[0]))
    if args.classes:
        mc.set_option("time_kernels", 1)
        kt = []
        for _ in range(5):
            mc.call_device(b, r, stream)
            kt.append(mc.kernel_times_ms())
        kt = np.median(np.array(kt), axis=0)
        cnt = np.bincount(hb.nals, minlength=6) * args.rep
        out["class_ms"] = {str(k): round(float(kt[k]), 4) for k in range(1, 6)}
        out["class_calls_per_s"] = {str(k): (float(cnt[k] * params.nsmpl / (kt[k] * 1e-3)) if kt[k] > 0 and cnt[k] else 0) for k in range(1, 6)}
        mc.set_option("time_kernels", 0)
    print(json.dumps(out), flush=True)
    if sweep_set != sets[-1]:
        mc.close()
if args.check:
    from oracle import pyoracle
    from tests import parity
    n = min(hb.nsites, 256)
    sub = hb.subset(range(n))
    exp, secs = pyoracle.call("port", params, sub, tab)
    got = mc.call_host(sub)
    print("parity", parity.compare(got, exp, params), "oracle calls/s %.3g" % (n * params.nsmpl / secs))
