"""Bundles the text fixtures of the reference's `call -m` tests (test/test.pl:276-308) into tests/golden/vcf_text_cases.json.gz:
input VCFs, expected outputs, and the option files they name (-S samples / PED, --ploidy-file, -G groups, -T targets).
Also extracts (I16, PV4) pairs from the consensus-caller cases (test/mpileup.c*.vcf -> *.out), which pin test16().
Run in the build container (needs /root/reference); the GPU box only sees the bundle.
    python tests/golden/make_vcf_fixtures.py"""
import gzip
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import vcf_cases

REF = "/root/reference/test"
files = {}
for c in vcf_cases.CASES:
    for name in [c["in"] + ".vcf", c["out"]] + c.get("files", []):
        files[name] = open(os.path.join(REF, name), "rb").read().decode("latin-1")
pairs = []
for inp, out in (("mpileup.c.vcf", "mpileup.c.1.out"), ("mpileup.c.X.vcf", "mpileup.c.X.out")):
    i16 = {}
    for line in open(os.path.join(REF, inp)):
        if line.startswith("#"):
            continue
        f = line.split("\t")
        for kv in f[7].split(";"):
            if kv.startswith("I16="):
                i16[(f[0], f[1], f[3], f[4].replace(",<*>", "").replace("<*>", "."))] = kv[4:]
    for line in open(os.path.join(REF, out)):
        if line.startswith("#"):
            continue
        f = line.split("\t")
        pv4 = [kv[4:] for kv in f[7].split(";") if kv.startswith("PV4=")]
        key = (f[0], f[1], f[3], f[4])
        cands = [v for k, v in i16.items() if k[:3] == key[:3]]
        if pv4 and len(cands) == 1:
            pairs.append([cands[0], pv4[0]])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vcf_text_cases.json.gz")
with gzip.GzipFile(out, "wb", mtime=0) as fh:
    fh.write(json.dumps({"files": files, "pv4_pairs": pairs}, sort_keys=True).encode())
print(out, len(files), "files", len(pairs), "PV4 pairs", os.path.getsize(out), "bytes")
