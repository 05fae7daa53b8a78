"""GPU parity tests proper: the CUDA path (through the C-ABI) versus the CPU oracle on identical inputs."""
import numpy as np
import pytest

from bcftools_b200 import abi, synth
from tests import parity

pytestmark = pytest.mark.gpu


BLOCKS = (0, 256, 128, 64, 32)      # 0 = the automatic per-class choice


def ORACLE(pyoracle):
    """The checker: the reference's own mcall.c compiled in place (oracle/_ref, ships to the GPU box); the plain-C port
    only where that library is missing."""
    return "reference" if pyoracle.have_ref() else "port"


def _run(params, batch, tab, oracle_built, options=None, kind=None):
    """Runs the CUDA path with every CTA size (0 = automatic: the class kernels of mcall_biallelic.cu / mcall_multi.cu where
    they apply; an explicit size = the general tiled kernel) and compares each against the oracle."""
    from bcftools_b200 import mcall
    want_gp = bool(params.output_tags & abi.CALL_FMT_GP)
    exp, _ = oracle_built.call(kind or ORACLE(oracle_built), params, batch, tab, want_gp=want_gp)
    st = None
    for block in BLOCKS:
        opts = dict(options or {})
        opts.setdefault("block", block)
        with mcall.MCaller(params, ploidy_tab=tab, options=opts) as mc:
            got = mc.call_host(batch, want_gp=want_gp)
        st = parity.compare(got, exp, params)
    return st


@pytest.mark.parametrize("cfg,nsites", [("C1", 10000), ("C2", 300), ("C3", 300), ("C5", 120)])
def test_synthetic_configs(cfg, nsites, oracle_built):
    params, batch, tab = synth.make_batch(cfg, nsites)
    st = _run(params, batch, tab, oracle_built)
    assert st["compared"] > 0 and not st["near_ties"], st


@pytest.mark.parametrize("flag", [0, abi.CALL_VARONLY, abi.CALL_KEEPALT])
@pytest.mark.parametrize("S,maxA", [(3, 5), (7, 5), (40, 4), (300, 5), (1000, 3)])
def test_random_adversarial(S, maxA, flag, oracle_built):
    rng = np.random.default_rng([S, maxA, flag])
    batch = parity.random_batch(rng, 200 if S < 100 else 60, S, maxA)
    params = abi.CallParams(S, maxA, flag=flag, output_tags=abi.CALL_FMT_GQ)
    st = _run(params, batch, None, oracle_built)
    assert st["compared"] > 0, st


@pytest.mark.parametrize("S,maxA", [(9, 5), (130, 4)])
def test_gp_output(S, maxA, oracle_built):
    """FORMAT/GP (-a GP): float32 posteriors bit-identical to the reference's (mcall.c:859-884)."""
    rng = np.random.default_rng([S, maxA, 7])
    batch = parity.random_batch(rng, 150, S, maxA, zq=False)
    tab = np.full((2, S), 2, np.uint8)
    tab[1, ::3] = 1
    tab[1, 1::7] = 0
    batch.ploidy_id = rng.integers(0, 2, batch.nsites).astype(np.uint16)
    params = abi.CallParams(S, maxA, output_tags=abi.CALL_FMT_GQ | abi.CALL_FMT_GP)
    st = _run(params, batch, tab, oracle_built, kind="port")    # the compiled reference aborts on these inputs (assert(max), mcall.c:881)
    assert st["compared"] > 0, st


@pytest.mark.parametrize("S", [5, 64, 513])
def test_mixed_ploidy(S, oracle_built):
    rng = np.random.default_rng(S)
    batch = parity.random_batch(rng, 120, S, 5)
    tab = np.full((3, S), 2, np.uint8)
    tab[1, ::2] = 1
    tab[2, ::3] = 1
    tab[2, 1::5] = 0
    batch.ploidy_id = rng.integers(0, 3, batch.nsites).astype(np.uint16)
    params = abi.CallParams(S, 5, output_tags=abi.CALL_FMT_GQ)
    st = _run(params, batch, tab, oracle_built)
    assert st["compared"] > 0, st


@pytest.mark.parametrize("S", [2, 64, 1000, 2504])
@pytest.mark.parametrize("flag", [0, abi.CALL_KEEPALT])
def test_two_allele_kernel_under_ploidy_vectors(S, flag, oracle_built):
    """Two-allele sites with per-sample ploidy 0 / 1 / 2 (even S: the straight-line pair path of the warp kernel with the
    ploidy folded into selects; missing values and zero QS send sites / samples through the general code)."""
    from bcftools_b200 import mcall
    rng = np.random.default_rng(100 + S + flag)
    batch = parity.random_batch(rng, 60 if S > 500 else 160, S, 2, minA=2)
    tab = np.full((3, S), 2, np.uint8)
    tab[1, ::2] = 1
    tab[2, ::3] = 1
    tab[2, 1::5] = 0
    batch.ploidy_id = rng.integers(0, 3, batch.nsites).astype(np.uint16)
    params = abi.CallParams(S, 2, flag=flag, output_tags=abi.CALL_FMT_GQ)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    with mcall.MCaller(params, ploidy_tab=tab) as mc:
        got = mc.call_host(batch, compact=True)
    assert parity.compare(got, exp, params)["compared"] > 0
    if S == 2504:       # the C5 shape, pooled: every 2nd sample haploid
        p5, b5, t5 = synth.make_batch("C5", 200, with_groups=0)
        e5, _ = oracle_built.call(ORACLE(oracle_built), p5, b5, t5)
        with mcall.MCaller(p5, ploidy_tab=t5) as mc:
            g5 = mc.call_host(b5)
        st = parity.compare(g5, e5, p5)
        assert st["compared"] > 0 and not st["near_ties"], st


@pytest.mark.parametrize("S,maxA,mode,flag", [(30, 5, 3, 0), (9, 4, "single", abi.CALL_VARONLY), (200, 5, 7, abi.CALL_KEEPALT), (64, 3, 2, 0),
                                              (300, 5, "mixed", 0), (700, 4, 3, abi.CALL_VARONLY), (1000, 5, 5, 0), (1300, 5, 13, 0)])
def test_sample_groups(S, maxA, mode, flag, oracle_built):
    """-G: per-group quality sums from FORMAT/AD (float32, group order), per-group allele sets, union of the sets,
    QUAL of the best group, per-sample genotypes with the sample's own group (mcall.c:1466-1504, 1546-1561, 1608-1614)."""
    rng = np.random.default_rng([S, maxA, 99])
    batch = parity.random_batch(rng, 80, S, maxA)
    if mode == "single":
        groups = [[s] for s in range(S)]
    elif mode == "mixed":       # big groups (whole-CTA path) next to small ones (warp path), interleaved sample indices
        perm = rng.permutation(S)
        cuts = [0, 150, 250, 290, S]
        groups = [sorted(perm[cuts[k]:cuts[k + 1]].tolist()) for k in range(4)]
    else:
        groups = [list(range(k, S, mode)) for k in range(mode)]
    tab = np.full((2, S), 2, np.uint8)
    tab[1, ::3] = 1
    tab[1, 1::7] = 0
    batch.ploidy_id = rng.integers(0, 2, batch.nsites).astype(np.uint16)
    params = abi.CallParams(S, maxA, flag=flag, output_tags=abi.CALL_FMT_GQ, groups=groups)
    st = _run(params, batch, tab, oracle_built)
    assert st["compared"] > 0, st


def test_grouped_launch_order_option_takes_permutations_only():
    from bcftools_b200 import mcall
    params = abi.CallParams(64, 5, output_tags=abi.CALL_FMT_GQ, groups=[list(range(0, 64, 2)), list(range(1, 64, 2))])
    with mcall.MCaller(params) as mc:
        mc.set_option("gorder", 12345)
        mc.set_option("gorder", 35421)
        for bad in (11111, 5432, 123456, 12340, -1):
            with pytest.raises(Exception):
                mc.set_option("gorder", bad)


def test_grouped_call_reports_per_class_times():
    """time_kernels=1 in a grouped call: the classes run one after the other and every class with sites reports a duration
    (what scripts/quick_bench.py --classes and the sweeps under profiles/r02_groups_* read)."""
    from bcftools_b200 import mcall
    import torch
    rng = np.random.default_rng(77)
    S = 320
    batch = parity.random_batch(rng, 64, S, 5)
    params = abi.CallParams(S, 5, output_tags=abi.CALL_FMT_GQ, groups=[list(range(k, S, 4)) for k in range(4)])
    from bcftools_b200 import device
    db = device.DeviceBatch(batch)
    dr = device.DeviceResult(db)
    with mcall.MCaller(params, options={"time_kernels": 1}) as mc:
        mc.call_device(db.c_struct(), dr.c_struct(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ms = mc.kernel_times_ms()
    cnt = np.bincount(batch.nals, minlength=6)
    assert ms[0] > 0 and all(ms[k] > 0 for k in range(1, 6) if cnt[k])


@pytest.mark.parametrize("S,ng,flag,prior", [(64, 2, 0, False), (333, 5, abi.CALL_VARONLY, False), (1000, 16, 0, True), (1000, 17, abi.CALL_KEEPALT, False),
                                             (2504, 5, 0, False), (2504, 26, 0, True), (640, 32, 0, False)])
def test_two_allele_grouped_kernel(S, ng, flag, prior, oracle_built):
    """-G on two-allele sites: the warp-per-site kernel of mcall_biallelic_groups.cu (sequential float32 AD sums per group, one
    pass per group over the packed copy, per-sample group records in phase 2) against the reference; ploidy 0 / 1 / 2, samples
    without data (PL = 0,0,0 and first-value-missing vectors), -F priors, sites whose ALT is the unseen allele (fallback)."""
    from bcftools_b200 import mcall
    rng = np.random.default_rng([S, ng, flag, 5])
    R = 96
    batch = parity.random_batch(rng, R, S, 2, zq=False, miss=False, minA=2)
    for i in range(R):                      # whole-vector missing samples: first value missing, the rest vector_end
        blk = batch.site_pl(i)
        m = rng.random(S) < 0.03
        blk[m] = abi.INT32_VECTOR_END
        blk[m, 0] = abi.INT32_MISSING
        if i % 11 == 3:
            blk[rng.integers(0, S)] = [255, 255, 255]       # a genuine all-255 vector: the general kernel's
    batch.unseen = np.where(np.arange(R) % 13 == 5, 1, 0).astype(np.uint8)
    groups = [list(range(k, S, ng)) for k in range(ng)]
    tab = np.full((3, S), 2, np.uint8)
    tab[1, ::2] = 1
    tab[2, ::3] = 1
    tab[2, 1::5] = 0
    batch.ploidy_id = rng.integers(0, 3, R).astype(np.uint16)
    if prior:
        batch.prior_an = rng.integers(10, 400, R).astype(np.int32)
        batch.prior_ac = np.full((R, 2), abi.INT32_VECTOR_END, np.int32)
        batch.prior_ac[:, 0] = (batch.prior_an * rng.random(R) * 0.5).astype(np.int32)
    params = abi.CallParams(S, 2, flag=flag, output_tags=abi.CALL_FMT_GQ, groups=groups, use_prior=prior)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    with mcall.MCaller(params, ploidy_tab=tab, options={"bgroups": 1}) as mc:
        got = mc.call_host(batch)
        n_new = int(mc.stats()[0])
    with mcall.MCaller(params, ploidy_tab=tab, options={"bgroups": 0}) as mc:
        old = mc.call_host(batch)
        n_old = int(mc.stats()[0])
    assert n_new == n_old + 1                # the extra launch is the warp-per-site kernel
    st = parity.compare(got, exp, params)
    assert st["compared"] > 0, st
    assert parity.compare(old, exp, params)["compared"] == st["compared"]


@pytest.mark.parametrize("S,maxA,groups,flag,tags", [(12, 8, None, 0, abi.CALL_FMT_GQ), (40, 7, None, abi.CALL_VARONLY, abi.CALL_FMT_GQ | abi.CALL_FMT_GP),
                                                     (25, 9, 3, abi.CALL_KEEPALT, abi.CALL_FMT_GQ), (6, 32, None, 0, abi.CALL_FMT_GQ),
                                                     (150, 6, None, 0, abi.CALL_FMT_GQ)])
def test_more_than_five_alleles(S, maxA, groups, flag, tags, oracle_built):
    """6..32 alleles (the reference accepts up to 32, mcall.c:1539-1543): the generic kernel of mcall_generic.cu,
    mixed in one batch with sites of 1..5 alleles that go through the templated kernels."""
    rng = np.random.default_rng([S, maxA, 3])
    batch = parity.random_batch(rng, 40 if maxA > 16 else 90, S, maxA, zq=not (tags & abi.CALL_FMT_GP))
    g = None if groups is None else [list(range(k, S, groups)) for k in range(groups)]
    tab = np.full((2, S), 2, np.uint8)
    tab[1, ::3] = 1
    tab[1, 1::7] = 0
    batch.ploidy_id = rng.integers(0, 2, batch.nsites).astype(np.uint16)
    params = abi.CallParams(S, maxA, flag=flag, output_tags=tags, groups=g)
    st = _run(params, batch, tab, oracle_built, kind="port" if tags & abi.CALL_FMT_GP else None)   # GP: assert(max), mcall.c:881
    assert st["compared"] > 0 and (batch.nals > 5).sum() > 0, st


def test_streaming_ring_matches_resident(oracle_built):
    """Small tiles force the streaming (two TMA passes) path; results must not depend on the tiling."""
    params, batch, tab = synth.make_batch("C3", 64)
    for opts in ({"tile_bytes": 4096, "ring_bytes": 8192}, {"tile_bytes": 16384, "ring_bytes": 200000}):
        st = _run(params, batch, tab, oracle_built, options=opts)
        assert st["compared"] > 0 and not st["near_ties"], (opts, st)


def test_biobank_shape_100k_samples(oracle_built):
    """BASELINE config 4 shape: 100,000 samples per site (1.2 MB of PL per site, streamed through the ring twice)."""
    params, batch, tab = synth.make_batch("C4", 6)
    st = _run(params, batch, tab, oracle_built)
    assert st["compared"] > 0 and not st["near_ties"], st


@pytest.mark.parametrize("flag", [0, abi.CALL_KEEPALT])
def test_tiled_two_allele_pair_path_above_8192_samples(flag, oracle_built):
    """S > 8,192: the two-allele class runs the tiled kernel, whose variant sites take the straight-line pair path
    (two adjacent samples per lane); adversarial PLs incl. missing values and zero QS send warps back to the general path."""
    from bcftools_b200 import mcall
    S = 8194
    rng = np.random.default_rng(8194 + flag)
    batch = parity.random_batch(rng, 24, S, 2, minA=2)
    params = abi.CallParams(S, 2, flag=flag, output_tags=abi.CALL_FMT_GQ)
    tab = np.full((1, S), 2, np.uint8)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    for opts in ({}, {"tile_bytes": 4096, "ring_bytes": 8192}):
        with mcall.MCaller(params, ploidy_tab=tab, options=opts) as mc:
            got = mc.call_host(batch, compact=True)
        assert parity.compare(got, exp, params)["compared"] > 0, opts


def test_full_size_properties_without_oracle():
    """Size-independent properties at a batch the CPU oracle would need minutes for (C3 shape, 4096 sites = 10 M calls):
    AN = sum(AC) = number of called alleles in GT; every GT allele index < ret; trimmed PL rows keep a zero for
    samples whose input row had its zero on a kept genotype; idempotence (a second run is bit-identical)."""
    from bcftools_b200 import mcall
    params, batch, tab = synth.make_batch("C3", 4096)
    with mcall.MCaller(params, ploidy_tab=tab) as mc:
        r1 = mc.call_host(batch)
        r2 = mc.call_host(batch)
    for name in ("ret", "als_new", "als_map", "ac", "an", "gt", "gq", "site_flags"):
        assert (getattr(r1, name) == getattr(r2, name)).all(), name
    assert (r1.qual.view(np.uint32) == r2.qual.view(np.uint32)).all()
    called = r1.ret > 0
    assert called.all()                                     # no -v: every site is emitted
    gt = r1.gt[called]
    alle = (gt >> 1) - 1                                    # -1 = missing
    n_called = (alle >= 0).sum(axis=(1, 2))
    assert (r1.an[called] == n_called).all()
    assert (r1.ac[called].sum(1) == r1.an[called]).all()
    assert (alle.max(axis=(1, 2)) < r1.ret[called]).all()
    for j in range(5):
        assert (r1.ac[called][:, j] == (alle == j).sum(axis=(1, 2))).all()
    for i in np.where(called & ((r1.site_flags & abi.SITE_PL_DROPPED) == 0))[0][:200]:
        tp = r1.site_pl(i)
        assert tp.shape[1] == r1.ret[i] * (r1.ret[i] + 1) // 2
        src = batch.site_pl(i)
        kept = [k for k in range(src.shape[1])]
        assert ((tp >= 0) | (tp == abi.INT32_MISSING) | (tp == abi.INT32_VECTOR_END)).all()


@pytest.mark.parametrize("cfg,nsites,flag", [("C3", 200, 0), ("C2", 100, abi.CALL_VARONLY)])
def test_int16_pl_transport(cfg, nsites, flag, oracle_built):
    """mcb_batch.pl_type = 2: the BCF on-disk int16 typed vectors are shipped as they are and widened on the device
    (SURVEY.md 8f N1); results must equal the int32 path bit for bit, including missing values and the PL fill."""
    from bcftools_b200 import mcall
    params, batch, tab = synth.make_batch(cfg, nsites, flag=flag)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    b16 = batch.to_int16()
    with mcall.MCaller(params, ploidy_tab=tab) as mc:
        got = mc.call_host(b16, compact=True)
    st = parity.compare(got, exp, params)
    assert st["compared"] > 0 and not st["near_ties"], st
    rng = np.random.default_rng(5)
    rb = parity.random_batch(rng, 120, 40, 5)
    tabr = np.full((2, 40), 2, np.uint8)
    tabr[1, ::3] = 1
    tabr[1, 1::7] = 0
    rb.ploidy_id = rng.integers(0, 2, rb.nsites).astype(np.uint16)
    pr = abi.CallParams(40, 5, output_tags=abi.CALL_FMT_GQ)
    expr, _ = oracle_built.call(ORACLE(oracle_built), pr, rb, tabr)
    with mcall.MCaller(pr, ploidy_tab=tabr) as mc:
        gotr = mc.call_host(rb.to_int16())
    assert parity.compare(gotr, expr, pr)["compared"] > 0


@pytest.mark.parametrize("int16", [False, True])
def test_idle_warps_do_not_index_the_table_with_sentinels(int16, oracle_built):
    """Regression: with 40 samples and 128-thread CTAs, warps 2-3 have no valid lane and re-read the last row; when that
    row is a missing-value row its sentinel must not reach the pl2p table lookup (it faulted with int16 PLs and was
    silently masked by 32-bit wrap-around with int32 PLs)."""
    from bcftools_b200 import mcall
    rng = np.random.default_rng(100)
    rb = parity.random_batch(rng, 60, 40, 1, minA=1, miss=True)
    for i in range(rb.nsites):                   # make the LAST row of every site a missing row
        rb.site_pl(i)[-1, :] = abi.INT32_MISSING
    rb.max_nals = 5
    qs = np.zeros((rb.nsites, 5), np.float32)
    qs[:, :rb.qs.shape[1]] = rb.qs
    rb.qs = qs
    pr = abi.CallParams(40, 5, output_tags=abi.CALL_FMT_GQ)
    exp, _ = oracle_built.call(ORACLE(oracle_built), pr, rb, None)
    with mcall.MCaller(pr) as mc:
        got = mc.call_host(rb.to_int16() if int16 else rb)
    assert parity.compare(got, exp, pr)["compared"] > 0


@pytest.mark.parametrize("flag", [0, abi.CALL_VARONLY, abi.CALL_KEEPALT])
@pytest.mark.parametrize("S", [1, 3, 4, 37, 128, 131, 1001, 2504, 4000])
def test_biallelic_warp_kernel(S, flag, oracle_built):
    """The warp-per-site kernel of the two-allele class (mcall_biallelic.cu): sample counts around its 4-sample lane
    groups and 128-sample iterations (odd totals take the unaligned store path), missing / vector_end / PL >= 256
    values (general-path iterations), an unseen ALT (`<*>` selected: one PL per sample is kept), zero QS entries,
    compacted and in-place PL output.  Forced on with warp2=<warps per CTA>; compared with the oracle and with the
    tiled kernel."""
    from bcftools_b200 import mcall
    rng = np.random.default_rng([S, flag, 11])
    R = 96 if S < 1000 else 40
    batch = parity.random_batch(rng, R, S, 2, minA=2, pl_max=300 if S in (37, 131) else 256)
    batch.unseen[rng.random(R) < 0.25] = 1
    for i in range(0, R, 7):            # clean sites: every iteration on the packed fast path
        blk = batch.site_pl(i)
        blk[...] = rng.integers(0, 256, blk.shape)
        blk[np.arange(S), rng.integers(0, 3, S)] = 0
    params = abi.CallParams(S, 2, flag=flag, output_tags=abi.CALL_FMT_GQ)
    # PL >= 256 inside a partially missing row indexes pl2p[] out of bounds in the reference (mcall.c:522): the port defines it
    BW_ORACLE = "port" if S in (37, 131) else ORACLE(oracle_built)
    exp, _ = oracle_built.call(BW_ORACLE, params, batch, None)
    for opts, compact in (({"warp2": 14}, False), ({"warp2": 3}, True), ({"warp2": 0}, False)):
        with mcall.MCaller(params, options=opts) as mc:
            got = mc.call_host(batch, compact=compact)
        st = parity.compare(got, exp, params)
        assert st["compared"] > 0, (opts, st)
    # the same sites under mixed ploidy vectors (haploid and ploidy-0 samples): the PLOIDY instance of the kernel
    tab = np.full((3, S), 2, np.uint8)
    tab[1, ::2] = 1
    tab[2, ::3] = 1
    tab[2, 1::5] = 0
    batch.ploidy_id = rng.integers(0, 3, R).astype(np.uint16)
    exp, _ = oracle_built.call(BW_ORACLE, params, batch, tab)
    for opts, compact in (({"warp2": 14}, True), ({"warp2": 0}, False)):
        with mcall.MCaller(params, ploidy_tab=tab, options=opts) as mc:
            got = mc.call_host(batch, compact=compact)
        st = parity.compare(got, exp, params)
        assert st["compared"] > 0, (opts, st)


def test_compacted_pl_output(oracle_built):
    """mcb_result.pl_off_out: trimmed PL blocks packed at the front of the output buffer (what leaves the device in
    the host path); content must be identical to the in-place layout, several slabs per call."""
    from bcftools_b200 import mcall
    params, batch, tab = synth.make_batch("C3", 200, flag=abi.CALL_VARONLY)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    with mcall.MCaller(params, ploidy_tab=tab, options={"slab_bytes": 4 << 20}) as mc:
        got = mc.call_host(batch, compact=True)
    st = parity.compare(got, exp, params)
    assert st["compared"] > 0
    used = sorted((int(got.pl_off_out[i]), int(got.ret[i])) for i in range(batch.nsites) if got.pl_off_out[i] >= 0)
    end = 0
    for off, n in used:                      # blocks are disjoint, 16-byte aligned and gap-free
        assert off == end and off % 4 == 0
        end = off + abi.pad4(params.nsmpl * n * (n + 1) // 2)
    dropped = (got.site_flags & abi.SITE_PL_DROPPED) != 0
    assert ((got.pl_off_out < 0) == (dropped | (got.ret <= 0))).all()


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("int16_in", [False, True])
def test_bcf_typed_outputs(compact, int16_in, oracle_built):
    """mcb_result.gt8 / gq8 / pl16: GT / GQ / trimmed PL narrowed on the device to the BCF types must
    widen back to exactly the int32 results (sentinels included: haploid vector_end, ploidy-0 missing), several slabs."""
    from bcftools_b200 import mcall
    params, batch, tab = synth.make_batch("C3", 160, flag=0)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    bin_ = batch.to_int16() if int16_in else batch
    with mcall.MCaller(params, ploidy_tab=tab, options={"slab_bytes": 4 << 20}) as mc:
        got = mc.call_host(bin_, compact=compact, typed=True)
    assert got.gt8.dtype == np.int8 and got.pl16.dtype == np.int16
    st = parity.compare(got.widen(), exp, params)
    assert st["compared"] > 0, st
    rng = np.random.default_rng(11)
    rb = parity.random_batch(rng, 150, 37, 5)
    tabr = np.full((2, 37), 2, np.uint8)
    tabr[1, ::3] = 1
    tabr[1, 1::7] = 0
    rb.ploidy_id = rng.integers(0, 2, rb.nsites).astype(np.uint16)
    pr = abi.CallParams(37, 5, output_tags=abi.CALL_FMT_GQ)
    expr, _ = oracle_built.call(ORACLE(oracle_built), pr, rb, tabr)
    with mcall.MCaller(pr, ploidy_tab=tabr) as mc:
        gotr = mc.call_host(rb, compact=compact, typed=True)
    assert parity.compare(gotr.widen(), expr, pr)["compared"] > 0


@pytest.mark.parametrize("flag", [0, abi.CALL_VARONLY])
@pytest.mark.parametrize("S", [128, 130, 300, 1000, 1280, 1282, 2504, 2560, 4000, 6000])
def test_multi_allelic_kernel(S, flag, oracle_built):
    """The CTA-per-site kernel of the 3-5 allele classes (mcall_multi.cu): sample counts around its 64-sample warp tiles
    and its three CTA sizes (64 / 128 / 256 threads), adversarial PLs -- missing and partially missing rows (the list
    warp 0 evaluates; more than 32 of them hand the site back to the general kernel through the fallback list), zero QS
    entries (dead allele sets), unseen alleles and PL >= 256 (fallback list again) -- compacted and in-place PL output,
    every ring depth; compared with the oracle and with the general tiled kernel."""
    from bcftools_b200 import mcall
    rng = np.random.default_rng([S, flag, 23])
    R = 60 if S > 1500 else 120
    batch = parity.random_batch(rng, R, S, 5, minA=3, pl_max=256)
    for i in range(0, R, 5):            # PL >= 256 on some sites
        blk = batch.site_pl(i)
        blk[rng.integers(0, S, 3), rng.integers(0, blk.shape[1], 3)] = 300
    for i in range(3, R, 4):            # a handful of special rows per site (the bench workload's shape): the list stays below its 32 entries
        blk = batch.site_pl(i)
        sp = (blk < 0).any(1)
        keep = rng.choice(np.where(sp)[0], size=min(int(sp.sum()), 7), replace=False) if sp.any() else []
        fix = sp.copy(); fix[keep] = False
        blk[fix] = rng.integers(0, 256, (int(fix.sum()), blk.shape[1]))
        blk[fix, 0] = 0
    for i in range(1, R, 4):            # clean sites: every pair on the fast path, every allele live
        blk = batch.site_pl(i)
        blk[...] = rng.integers(0, 256, blk.shape)
        blk[np.arange(S), rng.integers(0, blk.shape[1], S)] = 0
        batch.qs[i, :batch.nals[i]] = (1 + rng.random(int(batch.nals[i])) * 10).astype(np.float32)
        batch.unseen[i] = 0
    for i in range(2, R, 9):            # mpileup-shaped sites: most samples REF/REF, so that pairs and triples WITH the REF allele win
        A = int(batch.nals[i]); blk = batch.site_pl(i)
        truth = rng.choice(A, size=(S, 2), p=np.array([0.8] + [0.2 / (A - 1)] * (A - 1)))
        k = 0
        for x in range(A):
            for y in range(x + 1):
                m = ((truth[:, 0] == x) & (truth[:, 1] == y)) | ((truth[:, 0] == y) & (truth[:, 1] == x))
                one = (truth == x).any(1) | (truth == y).any(1)
                blk[:, k] = np.where(m, 0, np.where(one, 20 + rng.integers(0, 9, S), 200 + rng.integers(0, 50, S)))
                k += 1
        batch.qs[i, :A] = np.array([((truth == x).sum()) for x in range(A)], np.float32)
        batch.unseen[i] = 0
    params = abi.CallParams(S, 5, flag=flag, output_tags=abi.CALL_FMT_GQ)
    exp, _ = oracle_built.call("port", params, batch, None)     # PL >= 256 next to missing values: undefined in the reference (mcall.c:522)
    for opts, compact in (({}, False), ({"mm_nst": 1, "mm_block": 64}, True), ({"mm_nst": 3, "mm_block": 256}, False), ({"mm_nst": 4, "mm_block": 128}, True),
                          ({"multi": 0}, True)):
        with mcall.MCaller(params, options=opts) as mc:
            got = mc.call_host(batch, compact=compact)
        st = parity.compare(got, exp, params)
        assert st["compared"] > 0, (opts, st)


@pytest.mark.parametrize("cfg,nsites", [("C3", 2048)])
def test_multi_allelic_kernel_on_the_bench_workload(cfg, nsites, oracle_built):
    """2,048 sites of the C3 mix against the compiled reference; the multi-allelic classes must run mcall_multi.cu
    (launch count) and hand only a small share of their sites back to the general kernel."""
    from bcftools_b200 import mcall
    params, batch, tab = synth.make_batch(cfg, nsites)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    with mcall.MCaller(params, ploidy_tab=tab) as mc:
        got = mc.call_host(batch, compact=True)
        n_multi = int(mc.stats()[0])
    with mcall.MCaller(params, ploidy_tab=tab, options={"multi": 0}) as mc:
        mc.call_host(batch, compact=True)
        n_old = int(mc.stats()[0])
    assert n_multi > n_old                   # three more launches per slab: the multi-allelic kernels
    st = parity.compare(got, exp, params)
    assert st["compared"] > 0 and not st["near_ties"], st


@pytest.mark.parametrize("cfg,nsites,groups", [("C2", 2048, None), ("C5", 2048, 0), ("C5", 2048, 5), ("C5", 2048, 26), ("C4", 64, None)])
def test_baseline_configs_against_the_compiled_reference(cfg, nsites, groups, oracle_built):
    """>= 2,048 sites of every BASELINE.json config (C3: the test above; C4, 100,000 samples: 64 sites) with the default
    kernels against the reference's own mcall.c; C5 pooled, with 5 and with 26 -G groups."""
    from bcftools_b200 import mcall
    params, batch, tab = synth.make_batch(cfg, nsites, with_groups=groups)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    with mcall.MCaller(params, ploidy_tab=tab) as mc:
        got = mc.call_host(batch, compact=True)
    st = parity.compare(got, exp, params)
    assert st["compared"] > 0 and not st["near_ties"], st


def test_typed_end_to_end_path_on_the_bench_workload(oracle_built):
    """The BCF typed transport both ways (int16 PL in; int8 GT, int8 GQ, int16 PL out, compacted) on 2,048 C3 sites
    against the compiled reference."""
    from bcftools_b200 import mcall
    params, batch, tab = synth.make_batch("C3", 2048)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    with mcall.MCaller(params, ploidy_tab=tab) as mc:
        got = mc.call_host(batch.to_int16(), compact=True, typed=True)
    st = parity.compare(got.widen(), exp, params)
    assert st["compared"] > 0 and not st["near_ties"], st


@pytest.mark.parametrize("S,maxA,flag", [(7, 5, 0), (40, 4, abi.CALL_VARONLY), (300, 5, 0), (64, 5, abi.CALL_KEEPALT)])
def test_literal_phase1_for_near_tie_adjudication(S, maxA, flag, oracle_built):
    """mcb_set_option("exact_phase1", 1): the literal sample-sequential sums of logs of mcall_find_best_alleles
    (mcall.c:591-710) on the device must give the reference's calls on adversarial inputs, mixed ploidy included."""
    from bcftools_b200 import mcall
    rng = np.random.default_rng([S, maxA, flag, 99])
    batch = parity.random_batch(rng, 80, S, maxA)
    tab = np.full((3, S), 2, np.uint8)
    tab[1, ::2] = 1
    tab[2, ::3] = 1
    tab[2, 1::5] = 0
    batch.ploidy_id = rng.integers(0, 3, batch.nsites).astype(np.uint16)
    params = abi.CallParams(S, maxA, flag=flag, output_tags=abi.CALL_FMT_GQ)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    with mcall.MCaller(params, ploidy_tab=tab, options={"exact_phase1": 1}) as mc:
        got = mc.call_host(batch)
    st = parity.compare(got, exp, params)
    assert st["compared"] > 0 and not st["near_ties"], st


def test_host_batcher_adjudicates_near_ties(oracle_built):
    """With tie_eps so large that every record counts as a near tie, the batcher re-submits the whole batch through the
    literal phase 1: results must still be the reference's and carry the ADJUDICATED flag."""
    from bcftools_b200 import host_call
    rng = np.random.default_rng(4242)
    batch = parity.random_batch(rng, 150, 37, 5)
    params = abi.CallParams(37, 5, output_tags=abi.CALL_FMT_GQ)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, None)
    for async_flush in (False, True):
        got = host_call.replay(params, batch, None, max_records=64, async_flush=async_flush, tie_eps=1e30)
        called = got.ret > 0
        tie, adj = (got.site_flags[called] & abi.SITE_NEAR_TIE) != 0, (got.site_flags[called] & (1 << 9)) != 0
        assert (tie == adj).all() and adj.mean() > 0.5       # a site with a single candidate set has no runner-up: not a tie
        st = parity.compare(got, exp, params)
        assert st["compared"] > 0, st
    plain = host_call.replay(params, batch, None, max_records=64)
    assert not (plain.site_flags & (1 << 9)).any()


def test_empty_batch():
    from bcftools_b200 import mcall
    params = abi.CallParams(8, 5)
    batch = abi.HostBatch(8, 5, np.zeros(0, np.uint8), pl_blocks=[], qs=np.zeros((0, 5), np.float32))
    with mcall.MCaller(params) as mc:
        res = mc.call_host(batch)
    assert res.ret.size == 0


from tests import golden_util  # noqa: E402


@pytest.mark.parametrize("name", golden_util.case_names())
def test_cuda_reproduces_reference_goldens(name):
    """The CUDA path against the reference's own golden records (tests/golden, from test/test.pl:276-308)."""
    from bcftools_b200 import mcall
    params, batch, tab, case = golden_util.load_case(name)
    want_gp = bool(params.output_tags & abi.CALL_FMT_GP)
    with mcall.MCaller(params, ploidy_tab=tab) as mc:
        res = mc.call_host(batch, want_gp=want_gp)
    assert golden_util.check_against_expect(case, params, batch, res) == len(case["expect"])


@pytest.mark.parametrize("name", golden_util.case_names())
def test_host_batcher_replays_reference_goldens(name):
    """The C host layer (include/b200_call.h): records pushed one at a time through b200_mcall(), the way
    vcfcall.c:1089-1148 drives mcall(), small batches so that several flushes happen."""
    from bcftools_b200 import host_call
    params, batch, tab, case = golden_util.load_case(name)
    if any(s.get("qs") is None for s in case["sites"]) and params.ngroups <= 1:
        pytest.skip("no QS")
    res = host_call.replay(params, batch, tab, max_records=16)
    assert golden_util.check_against_expect(case, params, batch, res) == len(case["expect"])


@pytest.mark.parametrize("name", golden_util.case_names())
def test_host_batcher_bcf_typed_replays_reference_goldens(name):
    """b200_call_t.bcf_typed: the same replay with FORMAT/PL handed over as the int8/int16 typed vector a BCF record
    holds and GT / GQ / PL coming back as int8 / int8 / int16 vectors."""
    from bcftools_b200 import host_call
    params, batch, tab, case = golden_util.load_case(name)
    if params.ngroups > 1:
        pytest.skip("typed transport is for pooled calling")
    if any(s.get("qs") is None for s in case["sites"]):
        pytest.skip("no QS")
    if batch.pl.max() > 32767:
        pytest.skip("PL beyond int16")
    res = host_call.replay(params, batch, tab, max_records=16, typed=True)
    assert golden_util.check_against_expect(case, params, batch, res) == len(case["expect"])


@pytest.mark.parametrize("name", golden_util.case_names())
def test_async_host_batcher_replays_reference_goldens(name):
    """b200_call_t.async_flush: two slab sets, the batch in flight runs on the batcher's worker thread while records are
    queued into the other set; results arrive one batch late and must still come back complete and in input order."""
    from bcftools_b200 import host_call
    params, batch, tab, case = golden_util.load_case(name)
    if any(s.get("qs") is None for s in case["sites"]) and params.ngroups <= 1:
        pytest.skip("no QS")
    res = host_call.replay(params, batch, tab, max_records=5, async_flush=True)
    assert golden_util.check_against_expect(case, params, batch, res) == len(case["expect"])


@pytest.mark.parametrize("async_flush,groups", [(False, 0), (True, 0), (True, 3)])
def test_host_batcher_flushes_when_the_slab_is_full(async_flush, groups, oracle_built, monkeypatch):
    """The pinned PL / AD slabs hold a volume, not max_records worst-case records: a batch is handed over as soon as less than
    one worst-case record fits.  With the slab shrunk to 40,000 elements (B200_PL_SLAB_ELEMS) 300 records of 200 samples go
    out in many partial batches; every record must come back, in order, equal to the reference."""
    from bcftools_b200 import host_call
    monkeypatch.setenv("B200_PL_SLAB_ELEMS", "40000")
    S = 200
    rng = np.random.default_rng([5, int(async_flush), groups])
    batch = parity.random_batch(rng, 300, S, 5)
    params = abi.CallParams(S, 5, output_tags=abi.CALL_FMT_GQ, groups=[list(range(k, S, groups)) for k in range(groups)] if groups else None)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, None)
    res = host_call.replay(params, batch, None, max_records=256, async_flush=async_flush)      # 256 x 200 x 15 = 768,000 elements worst case
    assert parity.compare(res, exp, params)["compared"] > 0


@pytest.mark.parametrize("async_flush", [False, True])
def test_host_batcher_registers_each_distinct_ploidy_vector_once(async_flush, oracle_built):
    """chrX-style alternation (PAR / non-PAR): 1,000 records switch between two ploidy vectors; the batcher must end with
    two registered vectors (ids are looked up by content), and the calls must equal the oracle's."""
    from bcftools_b200 import host_call
    S = 24
    rng = np.random.default_rng(77)
    batch = parity.random_batch(rng, 1000, S, 3, minA=2, miss=False)
    tab = np.full((2, S), 2, np.uint8)
    tab[1, ::2] = 1
    batch.ploidy_id = (np.arange(1000) % 2).astype(np.uint16)
    params = abi.CallParams(S, 3, output_tags=abi.CALL_FMT_GQ)
    exp, _ = oracle_built.call(ORACLE(oracle_built), params, batch, tab)
    st = {}
    res = host_call.replay(params, batch, tab, max_records=37, async_flush=async_flush, stats=st)
    assert st["n_ploidy"] == 2, st
    assert parity.compare(res, exp, params)["compared"] > 0
