"""The reference's `call -m` test cases (test/test.pl:276-308) as data: options, input, expected output.  The texts come
from tests/golden/vcf_text_cases.json.gz (tests/golden/make_vcf_fixtures.py)."""
import gzip
import json
import os
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
BUNDLE = os.path.join(HERE, "golden", "vcf_text_cases.json.gz")


def _c(id_, in_, out, args, files=()):
    return dict(id=id_, **{"in": in_}, out=out, args=args, files=list(files))


def _cals(id_, in_, out, tab, extra=""):
    return _c(id_, in_, out, ("-mA -C alleles -T {PATH}/%s.tab %s" % (tab, extra)).split(), [tab + ".tab"])


CASES = [
    _c("mpileup.1", "mpileup", "mpileup.1.out", ["-mv"]),
    _c("mpileup.2-gvcf", "mpileup", "mpileup.2.out", ["-mg0"]),
    _c("mpileup.3", "mpileup", "mpileup.3.out", ["-mv", "-S", "{PATH}/mpileup.3.samples"], ["mpileup.3.samples"]),
    _c("mpileup.4", "mpileup", "mpileup.4.out", ["-mv", "-S", "{PATH}/mpileup.4.samples"], ["mpileup.4.samples"]),
    _c("mpileup.5", "mpileup", "mpileup.5.out", ["-mv", "-S", "{PATH}/mpileup.5.samples"], ["mpileup.5.samples"]),
    _c("mpileup.X-samples", "mpileup.X", "mpileup.X.out", ["-mv", "--ploidy-file", "{PATH}/mpileup.ploidy", "-S", "{PATH}/mpileup.samples"], ["mpileup.ploidy", "mpileup.samples"]),
    _c("mpileup.X-ped", "mpileup.X", "mpileup.X.out", ["-mv", "--ploidy-file", "{PATH}/mpileup.ploidy", "-S", "{PATH}/mpileup.ped"], ["mpileup.ploidy", "mpileup.ped"]),
    _c("mpileup.X.2", "mpileup.X", "mpileup.X.2.out", ["-mv", "--ploidy-file", "{PATH}/mpileup.ploidy", "-S", "{PATH}/mpileup.2.samples"], ["mpileup.ploidy", "mpileup.2.samples"]),
    _c("hwe.1", "mpileup.NA19213.NA19129", "mpileup.hwe.1.out", ["-mv"]),
    _c("hwe.1b", "mpileup.NA19213.NA19129", "mpileup.hwe.1b.out", ["-mv", "-G", "-", "--group-samples-tag", "AD"]),
    _c("hwe.2", "mpileup.hwe", "mpileup.hwe.2.out", ["-mv"]),
    _c("hwe.3", "mpileup.hwe", "mpileup.hwe.3.out", ["-mv", "-G", "-", "--group-samples-tag", "AD"]),
    _c("hwe.4", "mpileup.hwe", "mpileup.hwe.4.out", ["-mv", "-G", "{PATH}/mpileup.hwe.samples", "--group-samples-tag", "AD"], ["mpileup.hwe.samples"]),
    _cals("cAls.1", "mpileup", "mpileup.cAls.out", "mpileup"),
    _cals("cAls.2", "mpileup.2", "mpileup.cAls.2.out", "mpileup.2"),
    _cals("cAls.3", "mpileup.3", "mpileup.cAls.3.out", "mpileup.3", "-i"),
    _cals("cAls.4", "mpileup.3", "mpileup.cAls.4.out", "mpileup.4", "-i"),
    _cals("cAls.5", "mpileup.3", "mpileup.cAls.5.out", "mpileup.5", "-i"),
    _cals("cAls.6", "mpileup.4", "mpileup.cAls.6.out", "mpileup.6", "-i"),
    _cals("cAls.7", "mpileup.5", "mpileup.cAls.7.out", "mpileup.7", "-i"),
    _cals("cals.8", "mpileup.cals.1", "mpileup.cals.8.out", "mpileup.cals.1"),
    _cals("cals.9", "mpileup.cals.2", "mpileup.cals.9.out", "mpileup.cals.2"),
    _c("call-G.1", "call-G", "call-G.1.out", ["-mv"]),
    _c("call-G.2", "call-G", "call-G.2.out", ["-mv", "-G", "-", "--group-samples-tag", "AD"]),
    _c("call-G.2.1", "call-G.2", "call-G.2.1.out", ["-mv", "-F", "AN_POP,AC_POP"]),
    _c("af-fixation.1", "call.af-fixation", "call.af-fixation.1.out", ["-m"]),
    _c("af-fixation.2", "call.af-fixation", "call.af-fixation.2.out", ["-m", "-G", "{PATH}/call.af-fixation.txt"], ["call.af-fixation.txt"]),
    _c("af-fixation.3", "call.af-fixation", "call.af-fixation.3.out", ["-m", "-G", "{PATH}/call.af-fixation.txt", "-a", "GP,GQ"], ["call.af-fixation.txt"]),
]

_bundle = None
_tmp = None


def bundle():
    global _bundle
    if _bundle is None:
        _bundle = json.loads(gzip.open(BUNDLE).read())
    return _bundle


def load(case):
    """-> (input VCF bytes, expected output bytes, option list with {PATH} pointing at the materialised option files)"""
    global _tmp
    b = bundle()["files"]
    if _tmp is None:
        _tmp = tempfile.mkdtemp(prefix="b200_vcf_cases_")
    for name in case["files"]:
        path = os.path.join(_tmp, name)
        if not os.path.exists(path):
            with open(path, "wb") as fh:
                fh.write(b[name].encode("latin-1"))
    args = [a.replace("{PATH}", _tmp) for a in case["args"]]
    return b[case["in"] + ".vcf"].encode("latin-1"), b[case["out"]].encode("latin-1"), args


def write_input(case):
    inp, _, _ = load(case)
    path = os.path.join(_tmp, case["in"] + ".vcf")
    if not os.path.exists(path):
        with open(path, "wb") as fh:
            fh.write(inp)
    return path


def pv4_pairs():
    return [([float(x) for x in a.split(",")], b) for a, b in bundle()["pv4_pairs"]]
