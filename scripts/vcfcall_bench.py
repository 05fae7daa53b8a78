"""The htslib-free `call -m` driver end to end (include/b200_vcfcall.h): a synthetic C3-shaped VCF (2,504 samples, FORMAT/PL,
INFO/QS) goes file -> reader -> batcher -> CUDA -> finaliser -> writer -> file, as text and as BCF.  Reports records/s,
sample-genotype calls/s and input MB/s per container; the same records through mcb_call_host alone (arrays already unpacked)
are timed beside it, so the share of parsing / formatting is visible.  usage: python scripts/vcfcall_bench.py [--sites 2048]"""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bcftools_b200 import abi, mcall, synth, vcfcall

ap = argparse.ArgumentParser()
ap.add_argument("--sites", type=int, default=2048)
ap.add_argument("--config", default="C3")
args = ap.parse_args()
params, hb, tab = synth.make_batch(args.config, args.sites, with_groups=0)
S, R = params.nsmpl, hb.nsites
ALTS = ["C", "G", "T", "AC"]
t0 = time.time()
lines = ["##fileformat=VCFv4.2", "##contig=<ID=1,length=249250621>",
         '##INFO=<ID=QS,Number=R,Type=Float,Description="Auxiliary tag used for calling">',
         '##FORMAT=<ID=PL,Number=G,Type=Integer,Description="List of Phred-scaled genotype likelihoods">',
         "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join("s%d" % i for i in range(S))]
for i in range(R):
    n = int(hb.nals[i]); G = n * (n + 1) // 2
    pl = hb.site_pl(i).reshape(S, G)
    smp = [",".join(map(str, row)) for row in pl.tolist()]
    qs = ",".join("%g" % x for x in hb.qs[i, :n])
    lines.append("1\t%d\t.\tA\t%s\t.\t.\tQS=%s\tPL\t%s" % (1000 + i, ",".join(ALTS[:n - 1]), qs, "\t".join(smp)))
text = ("\n".join(lines) + "\n").encode()
gen_s = time.time() - t0
tmp = tempfile.mkdtemp()
vcf_in = os.path.join(tmp, "in.vcf"); open(vcf_in, "wb").write(text)
bcf_in = os.path.join(tmp, "in.bcf"); open(bcf_in, "wb").write(vcfcall.vcf_to_bcf(text, 1))
out = dict(workload=args.config, nsmpl=S, sites=R, vcf_bytes=len(text), bcf_bytes=os.path.getsize(bcf_in), gen_s=round(gen_s, 1), unit="calls/s")

def timed(argv, src, dst):
    vcfcall.run(argv, src, dst)                         # warm-up: context, pinned slabs
    t = time.perf_counter(); vcfcall.run(argv, src, dst); dt = time.perf_counter() - t
    return dict(s=round(dt, 3), records_per_s=round(R / dt), value=R * S / dt, in_MB_per_s=round(os.path.getsize(src) / dt / 1e6, 1), out_bytes=os.path.getsize(dst))

out["vcf_to_vcf"] = timed(["-m", "-a", "GQ"], vcf_in, os.path.join(tmp, "o.vcf"))
out["bcf_to_bcf_u"] = timed(["-m", "-a", "GQ", "-O", "u"], bcf_in, os.path.join(tmp, "o.u.bcf"))
out["bcf_to_bcf_b"] = timed(["-m", "-a", "GQ", "-O", "b"], bcf_in, os.path.join(tmp, "o.b.bcf"))
# same records, arrays already unpacked: the C-ABI host entry alone
sub = mcall.pin_batch(hb)
res = mcall.pin_result(abi.HostResult(sub, compact=True))
with mcall.MCaller(params, ploidy_tab=tab) as mc:
    mc.call_host(sub, res)
    t = time.perf_counter(); mc.call_host(sub, res); dt = time.perf_counter() - t
out["mcb_call_host_only"] = dict(s=round(dt, 4), value=R * S / dt)
# the two outputs agree record for record
a = open(os.path.join(tmp, "o.vcf"), "rb").read()
b = vcfcall.bcf_to_vcf(open(os.path.join(tmp, "o.u.bcf"), "rb").read())
out["text_equals_bcf_output"] = bool(a == b)
print(json.dumps(out))
