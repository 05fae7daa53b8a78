"""The C-ABI library must load on a CPU-only box and export every symbol include/mcall_b200.h declares
(no compute calls here), and it must refuse to run without a GPU instead of falling back to a CPU path."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mcall_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported():
    from bcftools_b200 import mcall
    L = mcall.lib()
    syms = declared_symbols()
    assert len(syms) >= 14, syms
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(mcall.EXPORTS) == syms, (sorted(mcall.EXPORTS), syms)


def test_host_layer_symbols_are_exported():
    """include/b200_call.h: the C mirror of mcall_init / mcall / mcall_destroy."""
    from bcftools_b200 import host_call, mcall
    hdr = open(os.path.join(ROOT, "include", "b200_call.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    syms = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))
    L = mcall.lib()
    assert syms == sorted(host_call.HOST_EXPORTS), syms
    assert all(hasattr(L, s) for s in syms)


def test_struct_layouts_match_header():
    """ctypes mirrors must have the field order of the C structs (a mismatch would silently scramble pointers)."""
    from bcftools_b200 import abi
    hdr = open(os.path.join(ROOT, "include", "mcall_b200.h")).read()
    for cname, cls in (("mcb_params", abi.McbParams), ("mcb_batch", abi.McbBatch), ("mcb_result", abi.McbResult)):
        body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (cname, cname), hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = [re.findall(r"\*?\s*([a-z_0-9]+)$", d.strip())[0] for d in body.split(";") if d.strip()]
        assert fields == [f[0] for f in cls._fields_], (cname, fields)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bcftools_b200 import abi, mcall
    with pytest.raises(mcall.McallError, match="no CUDA device"):
        mcall.MCaller(abi.CallParams(4))


def test_error_strings():
    from bcftools_b200 import mcall
    L = mcall.lib()
    assert b"CPU fallback" in L.mcb_strerror(-4)
    assert L.mcb_strerror(0) == b"ok"
