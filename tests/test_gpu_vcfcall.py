"""`bcftools call -m` end to end on the device: text VCF in, b200_vcfcall_run (reader -> filters -> batcher -> CUDA -> finaliser
-> gVCF / constrained alleles -> writer), text VCF out, compared byte for byte with the reference's expected outputs
(test/test.pl:276-308).  The CPU replay of the same cases is tests/test_vcfcall_host.py."""
import os

import pytest

from bcftools_b200 import vcfcall
from tests import vcf_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", vcf_cases.CASES, ids=[c["id"] for c in vcf_cases.CASES])
def test_call_m_on_the_device_reproduces_the_reference_output_bytes(case, tmp_path):
    _, exp, args = vcf_cases.load(case)
    out = str(tmp_path / "out.vcf")
    vcfcall.run(args, vcf_cases.write_input(case), out)
    got = open(out, "rb").read()
    if got != exp:
        g, e = got.split(b"\n"), exp.split(b"\n")
        for k, (a, b) in enumerate(zip(g, e)):
            assert a == b, (case["id"], k, a[:300], b[:300])
        assert len(g) == len(e), (case["id"], len(g), len(e))


@pytest.mark.parametrize("case_id,otype", [("mpileup.1", "b"), ("mpileup.2-gvcf", "u"), ("cAls.7", "b"), ("af-fixation.3", "z")])
def test_bcf_in_and_out(case_id, otype, tmp_path):
    """what test.pl's second command checks (`call -Ob ... | bcftools view`): BCF (or BGZF) input -> device -> -O b / u / z output,
    decoded back to text, equals the expected output"""
    import gzip
    case = [c for c in vcf_cases.CASES if c["id"] == case_id][0]
    inp, exp, args = vcf_cases.load(case)
    src = str(tmp_path / "in.bcf")
    with open(src, "wb") as fh:
        fh.write(vcfcall.vcf_to_bcf(inp))
    out = str(tmp_path / "out.bin")
    vcfcall.run(args + ["-O", otype], src, out)
    data = open(out, "rb").read()
    got = gzip.decompress(data) if otype == "z" else vcfcall.bcf_to_vcf(data)
    assert got == exp
