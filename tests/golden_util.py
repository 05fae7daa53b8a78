"""Loads tests/golden/*.json (made by tests/golden/make_golden.py from the reference's test/*.vcf -> *.out pairs)
and checks a result object against the expected VCF records."""
import glob
import json
import os

import numpy as np

from bcftools_b200 import abi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json")))


def load_case(name):
    case = json.load(open(os.path.join(GOLDEN_DIR, name + ".json")))
    S = case["nsmpl"]
    sites = case["sites"]
    max_nals = max(5, max(len(s["alleles"]) for s in sites))
    nals = np.array([len(s["alleles"]) for s in sites], np.uint8)
    qs = np.zeros((len(sites), max_nals), np.float32)
    nqs = np.zeros(len(sites), np.uint8)
    for i, s in enumerate(sites):
        v = s.get("qs", [])
        nqs[i] = len(v)
        qs[i, :len(v)] = [np.float32(float(x)) for x in v]
    groups = case["groups"]
    ad_blocks = [np.array(s["ad"], np.int32) for s in sites] if groups else None
    prior_an = prior_ac = None
    if case["use_prior"]:
        prior_an = np.array([s.get("prior_an", abi.INT32_MISSING) for s in sites], np.int32)
        prior_ac = np.full((len(sites), max_nals), abi.INT32_VECTOR_END, np.int32)
        for i, s in enumerate(sites):
            v = s.get("prior_ac", [])
            prior_ac[i, :len(v)] = v
    batch = abi.HostBatch(S, max_nals, nals, pl_blocks=[np.array(s["pl"], np.int32) for s in sites],
                          unseen=[s["unseen"] for s in sites], ploidy_id=[s["ploidy_id"] for s in sites],
                          qs=qs, nqs=nqs, ad_blocks=ad_blocks, prior_an=prior_an, prior_ac=prior_ac)
    params = abi.CallParams(S, max_nals, theta=case["theta"], flag=case["flag"], output_tags=case["output_tags"],
                            groups=groups, use_prior=case["use_prior"])
    tab = np.array(case["ploidy_vectors"], np.uint8)
    return params, batch, tab, case


def _gt_str(g):
    out = []
    for v in g:
        if v == abi.INT32_VECTOR_END:
            break
        out.append("." if (v >> 1) == 0 else str((v >> 1) - 1))
    return "/".join(out)


def _vec_str(row):
    out = []
    for v in row:
        if v == abi.INT32_VECTOR_END:
            break
        out.append("." if v == abi.INT32_MISSING else str(int(v)))
    return ",".join(out) if out else "."


def _fvec_str(row):
    out = []
    for v in row:
        bits = int(np.float32(v).view(np.uint32))
        if bits == abi.FLOAT_VECTOR_END_BITS:
            break
        out.append("." if bits == abi.FLOAT_MISSING_BITS else "%g" % float(v))
    return ",".join(out) if out else "."


def check_against_expect(case, params, batch, res):
    """Every record the reference wrote must be reproduced: ALT set, QUAL (6 significant digits, the precision
    of the golden text), AC, AN, GT, trimmed PL, GQ and GP."""
    emitted = [i for i in range(batch.nsites) if res.ret[i] > 0]
    exp = case["expect"]
    assert len(emitted) == len(exp), (case["name"], len(emitted), len(exp))
    for i, e in zip(emitted, exp):
        s = case["sites"][i]
        tag = (case["name"], e["chrom"], e["pos"])
        assert s["pos"] == e["pos"], tag
        amap = res.als_map[i]
        als = [a for _, a in sorted((amap[k], s["alleles"][k]) for k in range(len(s["alleles"])) if amap[k] >= 0)]
        als = als[:int(res.ret[i])]
        assert als == e["alleles"], (tag, als, e["alleles"])
        q = res.qual[i]
        if e["qual"] == ".":
            assert int(q.view(np.uint32)) == abi.FLOAT_MISSING_BITS, (tag, q)
        else:
            assert "%g" % float(q) == e["qual"] or abs(float(q) - float(e["qual"])) <= 1.5e-6 * abs(float(e["qual"])), (tag, q, e["qual"])
        n = int(res.ret[i])
        assert list(res.ac[i][1:n]) == e["ac"], (tag, res.ac[i], e["ac"])
        assert int(res.an[i]) == e["an"], (tag, res.an[i], e["an"])
        gts = [_gt_str(g) for g in res.gt[i]]
        assert gts == e["gt"], (tag, gts, e["gt"])
        if "pl" in e:
            assert not (res.site_flags[i] & abi.SITE_PL_DROPPED), tag
            pls = [_vec_str(r) for r in res.site_pl(i)]
            assert pls == e["pl"], (tag, pls, e["pl"])
        else:
            assert res.site_flags[i] & abi.SITE_PL_DROPPED, tag
        if "gq" in e:
            assert [str(int(x)) for x in res.gq[i]] == e["gq"], (tag, list(res.gq[i]), e["gq"])
        if "gp" in e and res.gp is not None:
            gps = [_fvec_str(r) for r in res.site_gp(i)]
            assert gps == e["gp"], (tag, gps, e["gp"])
    return len(exp)
