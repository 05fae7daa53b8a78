"""Host-side logic that needs no GPU: batch packing, synthetic generator, byte accounting, oracle cross-checks."""
import numpy as np
import pytest

from bcftools_b200 import abi, synth
from tests import parity


def test_hostbatch_layout_is_16_byte_aligned_and_padded():
    rng = np.random.default_rng(0)
    b = parity.random_batch(rng, 50, 7, 5)
    assert (b.pl_off % 4 == 0).all()
    for i in range(b.nsites):
        g = int(b.ngt[i])
        assert b.site_pl(i).shape == (7, g)
        end = b.pl_off[i] + 7 * g
        nxt = b.pl_off[i + 1] if i + 1 < b.nsites else b.pl.size
        assert end <= nxt and (b.pl[end:nxt] == abi.INT32_VECTOR_END).all()


def test_subset_roundtrip():
    params, b, tab = synth.make_batch("C3", 40)
    sub = b.subset([3, 7, 11])
    for k, i in enumerate([3, 7, 11]):
        assert (sub.site_pl(k) == b.site_pl(i)).all() and sub.nals[k] == b.nals[i]
        assert (sub.qs[k] == b.qs[i]).all()


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C5"])
def test_synthetic_generator_is_seeded_and_mpileup_shaped(cfg):
    p1, b1, _ = synth.make_batch(cfg, 30)
    p2, b2, _ = synth.make_batch(cfg, 30)
    assert (b1.pl == b2.pl).all() and (b1.qs == b2.qs).all()
    for i in range(b1.nsites):
        pl = b1.site_pl(i)
        ok = pl[(pl != abi.INT32_MISSING) & (pl != abi.INT32_VECTOR_END)]
        assert ok.min() >= 0 and ok.max() <= 255           # bam2bcf.c:645-646
    assert b1.nsmpl == synth.CONFIGS[cfg]["nsmpl"]


def test_restatement_equals_compiled_reference_on_random_inputs(oracle_built):
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    for (R, S, A, flag, tags, groups) in [(150, 7, 5, 0, abi.CALL_FMT_GQ, None), (120, 33, 4, abi.CALL_KEEPALT, abi.CALL_FMT_GQ, None),
                                          (120, 40, 5, abi.CALL_VARONLY, abi.CALL_FMT_GQ | abi.CALL_FMT_GP, None),
                                          (100, 30, 5, 0, abi.CALL_FMT_GQ, 3), (60, 9, 4, abi.CALL_VARONLY, abi.CALL_FMT_GQ, "single")]:
        b = parity.random_batch(rng, R, S, A, zq=not (tags & abi.CALL_FMT_GP))
        g = None
        if groups == 3:
            g = [list(range(k, S, 3)) for k in range(3)]
        elif groups == "single":
            g = [[s] for s in range(S)]
        params = abi.CallParams(S, A, flag=flag, output_tags=tags, groups=g)
        tab = np.full((2, S), 2, np.uint8)
        tab[1, ::3] = 1
        tab[1, 1::7] = 0
        b.ploidy_id = rng.integers(0, 2, R).astype(np.uint16)
        want_gp = bool(tags & abi.CALL_FMT_GP)
        a, _ = oracle_built.call("port", params, b, tab, want_gp=want_gp)
        r, _ = oracle_built.call("reference", params, b, tab, want_gp=want_gp)
        st = parity.compare(a, r, params, exact_qual=True)
        assert st["compared"] > 0 and not st["near_ties"]


def test_algorithmic_bytes_accounting(oracle_built):
    params, b, tab = synth.make_batch("C2", 20)
    res, _ = oracle_built.call("port", params, b, tab)
    rd, wr = synth.algorithmic_bytes(b, res, params.output_tags)
    S = params.nsmpl
    assert rd >= 20 * S * 3 * 4
    var = (res.ret == 2).sum()
    ref = (res.ret == 1).sum()
    assert wr == var * S * (8 + 4 + 12) + ref * S * 8       # 36 B per biallelic call with GQ and PL kept
