"""Shared helpers of the parity tests: randomized inputs and the result comparison rules.

Comparison rules (BASELINE.json north_star): ret / ALT selection (als_new, als_map) / GT / AC / AN /
trimmed PL bit-exact; GQ exact; QUAL within 1e-6 relative; sites whose two best allele sets are closer
than tie_eps are LISTED (returned) and exempt from the discrete checks only if they actually diverge.
"""
import numpy as np

from bcftools_b200 import abi

QUAL_RTOL = 1e-6


def mask_after_end(x):
    """Entries behind the first vector_end of a PL row are don't-care (the reference leaves stale values there)."""
    x = x.copy()
    seen = np.cumsum(x == abi.INT32_VECTOR_END, axis=1) > 0
    x[seen] = abi.INT32_VECTOR_END
    return x


def random_batch(rng, R, S, maxA, zq=True, miss=True, minA=1, pl_max=256):
    """Adversarial random sites: uniform PLs, zero QS entries, missing/partial-missing PLs, unseen alleles."""
    nals = rng.integers(minA, maxA + 1, R)
    blocks, ads = [], []
    qs = np.zeros((R, maxA), np.float32)
    unseen = np.zeros(R, np.uint8)
    for i in range(R):
        A = int(nals[i])
        G = A * (A + 1) // 2
        pl = rng.integers(0, pl_max, (S, G)).astype(np.int32)
        pl[np.arange(S), rng.integers(0, G, S)] = 0
        r = rng.random(S)
        pl[r < 0.05] = 0
        if miss:
            m = (r >= 0.05) & (r < 0.08)
            pl[m] = abi.INT32_VECTOR_END
            pl[m, 0] = abi.INT32_MISSING
            if G > 1:
                pm = np.where((r >= 0.08) & (r < 0.12))[0]
                pl[pm, rng.integers(1, G, len(pm))] = abi.INT32_MISSING
        blocks.append(pl)
        q = (rng.random(A) * 10).astype(np.float32)
        q[rng.random(A) < (0.2 if zq else 0)] = 0
        qs[i, :A] = q
        if A > 2 and rng.random() < 0.5:
            unseen[i] = A - 1
        ads.append(rng.integers(0 if zq else 1, 20, (S, A)).astype(np.int32))
    return abi.HostBatch(S, maxA, nals, pl_blocks=blocks, unseen=unseen, qs=qs, ad_blocks=ads)


def compare(got, exp, params, exact_qual=False, check_flags=True):
    """Compare two HostResult objects.  Returns dict(compared=, near_ties=[...], qual_max_rel=)."""
    near = []
    compared = 0
    qmax = 0.0
    S = params.nsmpl
    for i in range(len(exp.ret)):
        tie = bool(got.site_flags[i] & abi.SITE_NEAR_TIE)
        if got.ret[i] != exp.ret[i] or (exp.ret[i] > 0 and got.als_new[i] != exp.als_new[i]):
            if tie:
                near.append(dict(site=i, kind="allele-set", got=int(got.als_new[i]), exp=int(exp.als_new[i])))
                continue
            raise AssertionError(f"site {i}: ret/als_new differ: got {got.ret[i]}/{got.als_new[i]:b} exp {exp.ret[i]}/{exp.als_new[i]:b}")
        if exp.ret[i] <= 0:
            if check_flags:
                fm = abi.SITE_TOO_MANY_ALS | abi.SITE_NO_QS
                assert (int(got.site_flags[i]) & fm) == (int(exp.site_flags[i]) & fm), (i, got.site_flags[i], exp.site_flags[i])
            continue
        compared += 1
        assert (got.als_map[i] == exp.als_map[i]).all(), (i, "als_map", got.als_map[i], exp.als_map[i])
        fm = abi.SITE_PL_DROPPED | abi.SITE_UNSEEN_SEL
        assert (int(got.site_flags[i]) & fm) == (int(exp.site_flags[i]) & fm), (i, "flags", got.site_flags[i], exp.site_flags[i])
        if exp.site_flags[i] & abi.SITE_UNSEEN_SEL:
            continue        # reference behaviour undefined (writes past its arrays), SURVEY.md §8 quirks
        a, b = got.qual[i], exp.qual[i]
        if exact_qual:
            assert a.view(np.uint32) == b.view(np.uint32), (i, "qual", a, b)
        elif np.isnan(a) or np.isnan(b):
            assert a.view(np.uint32) == b.view(np.uint32), (i, "qual(missing)", a, b)
        else:
            rel = abs(float(a) - float(b)) / max(abs(float(a)), abs(float(b)), 1e-30) if a != b else 0.0
            qmax = max(qmax, rel)
            assert rel <= QUAL_RTOL, (i, "qual", a, b, rel)
        assert (got.ac[i] == exp.ac[i]).all(), (i, "ac", got.ac[i], exp.ac[i])
        assert got.an[i] == exp.an[i], (i, "an")
        assert (got.gt[i] == exp.gt[i]).all(), (i, "gt", np.where((got.gt[i] != exp.gt[i]).any(1))[0][:5])
        if not (exp.site_flags[i] & abi.SITE_PL_DROPPED):
            ga, ea = mask_after_end(got.site_pl(i)), mask_after_end(exp.site_pl(i))
            assert (ga == ea).all(), (i, "pl", np.where((ga != ea).any(1))[0][:5])
        if (params.output_tags & (abi.CALL_FMT_GQ | abi.CALL_FMT_GP)) and not (got.site_flags[i] & abi.SITE_REF_GT):
            d = np.where(got.gq[i] != exp.gq[i])[0]
            assert d.size == 0, (i, "gq", d[:5], got.gq[i][d[:5]], exp.gq[i][d[:5]])
            if exp.gp is not None and got.gp is not None:
                assert (got.site_gp(i).view(np.uint32) == exp.site_gp(i).view(np.uint32)).all(), (i, "gp")
    return dict(compared=compared, near_ties=near, qual_max_rel=qmax)
