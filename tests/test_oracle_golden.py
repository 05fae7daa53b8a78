"""Pins the CPU oracle: the restatement (and, where built, the compiled reference) must reproduce every golden
record of the reference's own `call -m` tests (test/test.pl:276-308 -> tests/golden/*.json)."""
import pytest

from bcftools_b200 import abi
from tests import golden_util


@pytest.mark.parametrize("name", golden_util.case_names())
def test_restatement_reproduces_reference_goldens(name, oracle_built):
    params, batch, tab, case = golden_util.load_case(name)
    res, _ = oracle_built.call("port", params, batch, tab, want_gp=bool(params.output_tags & abi.CALL_FMT_GP))
    assert golden_util.check_against_expect(case, params, batch, res) == len(case["expect"])


@pytest.mark.parametrize("name", golden_util.case_names())
def test_compiled_reference_reproduces_goldens(name, oracle_built):
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    params, batch, tab, case = golden_util.load_case(name)
    res, _ = oracle_built.call("reference", params, batch, tab, want_gp=bool(params.output_tags & abi.CALL_FMT_GP))
    assert golden_util.check_against_expect(case, params, batch, res) == len(case["expect"])
