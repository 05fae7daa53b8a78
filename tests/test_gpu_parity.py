"""GPU parity tests proper: the CUDA path (through the C-ABI) versus the CPU oracle on identical inputs."""
import numpy as np
import pytest

from bcftools_b200 import abi, synth
from tests import parity

pytestmark = pytest.mark.gpu


BLOCKS = (256, 128)


def _run(params, batch, tab, oracle_built, options=None, kind="port"):
    """Runs the CUDA path with both CTA sizes and compares each against the oracle."""
    from bcftools_b200 import mcall
    want_gp = bool(params.output_tags & abi.CALL_FMT_GP)
    exp, _ = oracle_built.call(kind, params, batch, tab, want_gp=want_gp)
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
    st = _run(params, batch, tab, oracle_built)
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


@pytest.mark.parametrize("S,maxA,mode,flag", [(30, 5, 3, 0), (9, 4, "single", abi.CALL_VARONLY), (200, 5, 7, abi.CALL_KEEPALT), (64, 3, 2, 0)])
def test_sample_groups(S, maxA, mode, flag, oracle_built):
    """-G: per-group quality sums from FORMAT/AD (float32, group order), per-group allele sets, union of the sets,
    QUAL of the best group, per-sample genotypes with the sample's own group (mcall.c:1466-1504, 1546-1561, 1608-1614)."""
    rng = np.random.default_rng([S, maxA, 99])
    batch = parity.random_batch(rng, 80, S, maxA)
    groups = [[s] for s in range(S)] if mode == "single" else [list(range(k, S, mode)) for k in range(mode)]
    tab = np.full((2, S), 2, np.uint8)
    tab[1, ::3] = 1
    tab[1, 1::7] = 0
    batch.ploidy_id = rng.integers(0, 2, batch.nsites).astype(np.uint16)
    params = abi.CallParams(S, maxA, flag=flag, output_tags=abi.CALL_FMT_GQ, groups=groups)
    st = _run(params, batch, tab, oracle_built)
    assert st["compared"] > 0, st


def test_streaming_ring_matches_resident(oracle_built):
    """Small tiles force the streaming (two TMA passes) path; results must not depend on the tiling."""
    params, batch, tab = synth.make_batch("C3", 64)
    for opts in ({"tile_bytes": 4096, "ring_bytes": 8192}, {"tile_bytes": 16384, "ring_bytes": 200000}):
        st = _run(params, batch, tab, oracle_built, options=opts)
        assert st["compared"] > 0 and not st["near_ties"], (opts, st)


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
