"""The `bcftools call -m` driver without htslib (include/b200_vcfcall.h), replayed on the CPU against the reference's own
expected outputs: every `call -m` case of test/test.pl:276-308 -- incl. -g (gVCF blocks), -C alleles -T [-i] (constrained
alleles), -S / PED / --ploidy-file, -G, -F, -a GQ,GP -- must come out byte for byte like test/<name>.out.
The host halves (everything in front of / behind the likelihood code) are the product's; the per-record results they
are fed here come from the CPU oracle, because this suite runs without a GPU (tests/test_gpu_vcfcall.py runs the same
cases through b200_vcfcall_run on the device)."""
import os

import numpy as np
import pytest

from bcftools_b200 import abi, vcfcall
from oracle import pyoracle
from tests import vcf_cases


def oracle_kind():
    return "reference" if pyoracle.have_ref() else "port"


def run_case_on_cpu(args, vcf_text):
    vc = vcfcall.VcfCall(args, vcf_text)
    try:
        S = vc.nsmpl
        groups = vc.groups()
        init_ploidy = vc.ploidy()
        while True:
            r = vc.next()
            if r is None:
                break
            A = r["n_allele"]
            M = max(5, A)
            G = A * (A + 1) // 2
            pl = r["pl"]
            assert pl.shape[1] == G, "haploid-shaped PL rows are not part of the batcher interface"
            qs = nqs = None
            if r["qs"] is not None:
                qs = np.zeros((1, M), np.float32)
                n = min(len(r["qs"]), M)
                qs[0, :n] = r["qs"][:n]
                nqs = [n]
            prior_an = prior_ac = None
            use_prior = bool(vc.call.use_prior)
            if use_prior:
                prior_an = [r["prior_an"]]
                prior_ac = np.full((1, M), abi.INT32_VECTOR_END, np.int32)
                if r["prior_ac"] is not None:
                    prior_ac[0, :len(r["prior_ac"])] = r["prior_ac"]
            batch = abi.HostBatch(S, M, [A], pl_blocks=[pl], unseen=[r["unseen"]], ploidy_id=[0], qs=qs, nqs=nqs,
                                  ad_blocks=None if r["ad"] is None else [r["ad"]], prior_an=prior_an, prior_ac=prior_ac)
            params = abi.CallParams(S, M, theta=vc.call.theta, init_ploidy=init_ploidy, flag=vc.call.flag,
                                    output_tags=vc.call.output_tags, groups=groups, use_prior=use_prior)
            res, _ = pyoracle.call(oracle_kind(), params, batch, r["ploidy"].reshape(1, -1),
                                   want_gp=bool(vc.call.output_tags & abi.CALL_FMT_GP))
            vc.finish(r["handle"], res, 0)
        vc.flush()
        return vc.output()
    finally:
        vc.close()


@pytest.mark.parametrize("case", vcf_cases.CASES, ids=[c["id"] for c in vcf_cases.CASES])
def test_call_m_case_reproduces_the_reference_output_bytes(case):
    inp, exp, args = vcf_cases.load(case)
    got = run_case_on_cpu(args, inp)
    if got != exp:
        g, e = got.split(b"\n"), exp.split(b"\n")
        for k, (a, b) in enumerate(zip(g, e)):
            assert a == b, (case["id"], k, a[:300], b[:300])
        assert len(g) == len(e), (case["id"], len(g), len(e))


def test_pv4_matches_the_values_in_the_reference_outputs():
    """test16 (ccall.c:115-138): PV4 of the consensus-caller outputs vs the I16 of the same records in their input"""
    pairs = vcf_cases.pv4_pairs()
    assert len(pairs) >= 10
    for i16, pv4_text in pairs:
        tested, p = vcfcall.pv4(i16)
        assert tested
        assert ",".join(vcfcall.format_float(x) for x in p) == pv4_text, (i16, p, pv4_text)
