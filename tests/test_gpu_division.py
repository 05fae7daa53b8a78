"""The per-sample normalisation pdg = p/sum (mcall.c:539) must be an IEEE division for GT to be bit-exact.
Phase 2 shares one refined reciprocal between the numerators of a sample; this test proves on the device that
the shared form is bit-identical to `a/b` over the whole biallelic table domain and on random multi-allelic sums."""
import pytest

from bcftools_b200 import abi

pytestmark = pytest.mark.gpu


def test_shared_reciprocal_division_is_ieee_exact():
    from bcftools_b200 import mcall
    with mcall.MCaller(abi.CallParams(4)) as mc:
        assert mc.selftest_div(0) == 0                      # 256^3 PL triples x 3 numerators
        for g in (6, 10, 15):
            assert mc.selftest_div(g, n=200_000_000, seed=g) == 0
