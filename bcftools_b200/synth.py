"""Synthetic FORMAT/PL workloads of the BASELINE.json shapes, numpy, seeded.

The data imitate `bcftools mpileup` output (bam2bcf.c:577-646 in the reference): alleles are REF then ALTs,
PL is capped at 255 with min 0 per sample, zero-coverage samples are PL=0,..,0, INFO/QS is the float32 sum over
samples of the per-sample allele fractions.  The same bytes feed the GPU path and the CPU oracle.

configs (BASELINE.json `configs`):
  C1  3 samples,       biallelic, diploid
  C2  1000 samples,    biallelic, diploid
  C3  2504 samples,    2-5 alleles (70/20/7/3 %), diploid, trimming on
  C4  100000 samples,  biallelic
  C5  2504 samples,    2-5 alleles, every 2nd sample haploid, 5 -G groups, FORMAT/AD
"""
import numpy as np

from . import abi

CONFIGS = {
    "C1": dict(nsmpl=3, seed=0xB2000001, amix={2: 1.0}, full_sites=10_000),
    "C2": dict(nsmpl=1000, seed=0xB2000002, amix={2: 1.0}, full_sites=1_000_000),
    "C3": dict(nsmpl=2504, seed=0xB2000003, amix={2: 0.70, 3: 0.20, 4: 0.07, 5: 0.03}, full_sites=1_000_000),
    "C4": dict(nsmpl=100_000, seed=0xB2000004, amix={2: 1.0}, full_sites=200_000),
    "C5": dict(nsmpl=2504, seed=0xB2000005, amix={2: 0.70, 3: 0.20, 4: 0.07, 5: 0.03}, full_sites=500_000,
               haploid_every=2, ngroups=5),
}
GENERATOR_VERSION = 1


def _gt_pairs(A):
    return [(a, b) for a in range(A) for b in range(a + 1)]     # index a(a+1)/2+b, a>=b  (bcf_alleles2gt)


def _class_blocks(rng, R, S, A, ploidy, p_nocov=0.02, p_missing=0.002):
    """PL [R,S,G] int32, QS [R,A] float32, AD [R,S,A] int32 for R sites with A alleles."""
    G = A * (A + 1) // 2
    # allele frequencies: f_1 = 0.5*10^(-3u); extra ALTs are artefacts (f=0) half of the time
    f = np.zeros((R, A))
    f[:, 1] = 0.5 * 10.0 ** (-3 * rng.random(R))
    for k in range(2, A):
        present = rng.random(R) < 0.5
        f[:, k] = np.where(present, 0.15 * 10.0 ** (-3 * rng.random(R)), 0.0)
    alt = f[:, 1:].sum(1)
    scale = np.where(alt > 0.9, 0.9 / np.maximum(alt, 1e-300), 1.0)
    f[:, 1:] *= scale[:, None]
    f[:, 0] = 1.0 - f[:, 1:].sum(1)
    cum = np.cumsum(f, 1)[:, None, :-1]                       # [R,1,A-1]
    t1 = (rng.random((R, S, 1)) > cum).sum(-1).astype(np.int8)
    t2 = (rng.random((R, S, 1)) > cum).sum(-1).astype(np.int8)
    hap = (ploidy[None, :] == 1)
    t2 = np.where(hap, t1, t2)
    d = (1 + rng.integers(0, 12, (R, S))).astype(np.int32)
    pl = np.empty((R, S, G), np.int32)
    for j, (x, y) in enumerate(_gt_pairs(A)):
        # multiset intersection size between (x,y) and the true genotype (t1,t2)
        m = ((t1 == x) & (t2 == y)) | ((t1 == y) & (t2 == x))
        one = (t1 == x) | (t1 == y) | (t2 == x) | (t2 == y)
        u = rng.integers(0, 7, (R, S)).astype(np.int32)
        v = np.where(one, 3 * d + u, 30 * d + u)
        pl[:, :, j] = np.where(m, 0, np.minimum(v, 255))
    r = rng.random((R, S))
    nocov = r < p_nocov
    missing = (r >= p_nocov) & (r < p_nocov + p_missing)
    pl[nocov] = 0
    if G > 1:
        pl[missing] = abi.INT32_VECTOR_END
    pl[missing, 0] = abi.INT32_MISSING
    covered = ~(nocov | missing)
    # INFO/QS: float32 running sum over samples of copies/ploidy (bam2bcf.c:566-575)
    qs = np.zeros((R, A), np.float32)
    ad = np.zeros((R, S, A), np.int32)
    for a in range(A):
        copies = ((t1 == a).astype(np.float32) + (t2 == a).astype(np.float32)) * 0.5
        copies = np.where(covered, copies, np.float32(0))
        qs[:, a] = np.cumsum(copies, axis=1, dtype=np.float32)[:, -1]
        ad[:, :, a] = np.where(covered, (copies * d).astype(np.int32), 0)
    # artefact alleles pick up a little error signal half of the time
    for k in range(2, A):
        noise = (rng.random(R) < 0.5) & (f[:, k] == 0)
        qs[:, k] += np.where(noise, rng.random(R) * 0.5, 0).astype(np.float32)
    return pl, qs, ad


def make_batch(config, nsites, seed_offset=0, flag=0, output_tags=abi.CALL_FMT_GQ, with_groups=None, chunk=256):
    """Returns (CallParams, HostBatch, ploidy_tab or None) for `nsites` sites of a named config."""
    cfg = CONFIGS[config]
    S = cfg["nsmpl"]
    rng = np.random.default_rng([cfg["seed"], seed_offset])
    alleles = sorted(cfg["amix"])
    probs = np.array([cfg["amix"][a] for a in alleles])
    nals = rng.choice(alleles, size=nsites, p=probs / probs.sum()).astype(np.uint8)
    max_nals = 5
    ploidy = np.full(S, 2, np.uint8)
    ploidy_tab = None
    if cfg.get("haploid_every"):
        ploidy[::cfg["haploid_every"]] = 1
        ploidy_tab = np.stack([np.full(S, 2, np.uint8), ploidy])
    ngt = nals.astype(np.int64) * (nals.astype(np.int64) + 1) // 2
    sizes = (S * ngt + 3) & ~3
    pl_off = np.zeros(nsites, np.int64)
    pl_off[1:] = np.cumsum(sizes)[:-1]
    pl = np.full(int(sizes.sum()), abi.INT32_VECTOR_END, np.int32)
    qs = np.zeros((nsites, max_nals), np.float32)
    groups = None
    use_groups = cfg.get("ngroups") if with_groups is None else with_groups
    ad_blocks = [None] * nsites if use_groups else None
    for A in alleles:
        idx = np.where(nals == A)[0]
        G = A * (A + 1) // 2
        for c0 in range(0, len(idx), chunk):
            ids = idx[c0:c0 + chunk]
            bpl, bqs, bad = _class_blocks(rng, len(ids), S, A, ploidy)
            qs[ids, :A] = bqs
            for k, i in enumerate(ids):
                pl[pl_off[i]:pl_off[i] + S * G] = bpl[k].reshape(-1)
                if ad_blocks is not None:
                    ad_blocks[i] = bad[k]
    if use_groups:
        ng = int(use_groups)
        groups = [list(range(g, S, ng)) for g in range(ng)]
    params = abi.CallParams(S, max_nals, flag=flag, output_tags=output_tags, groups=groups)
    batch = abi.HostBatch(S, max_nals, nals, pl=pl, pl_off=pl_off, qs=qs, ad_blocks=ad_blocks,
                          ploidy_id=None if ploidy_tab is None else np.ones(nsites, np.uint16))
    return params, batch, ploidy_tab


def algorithmic_bytes(batch, result, output_tags):
    """bytes per call = 4G (PL read) + 8 (GT) + 4 [GQ] + 4G' (trimmed PL, 0 when dropped),
    summed over the batch; plus the small per-site records."""
    S = batch.nsmpl
    G = batch.ngt.astype(np.int64)
    ret = result.ret.astype(np.int64)
    Gn = ret * (ret + 1) // 2
    dropped = (result.site_flags & abi.SITE_PL_DROPPED) != 0
    refgt = (result.site_flags & abi.SITE_REF_GT) != 0
    called = ret > 0
    rd = 4 * G * S
    wr = np.where(called, 8 * S + np.where(dropped, 0, 4 * Gn * S), 0)
    if output_tags & (abi.CALL_FMT_GQ | abi.CALL_FMT_GP):
        wr = wr + np.where(called & ~refgt, 4 * S, 0)
    per_site = 4 * batch.nals.astype(np.int64) + 4 + 4 * ret + 4 + 8
    return int(rd.sum() + per_site.sum()), int(wr.sum())
