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


def test_job_layer_symbols_are_exported():
    """include/mcall_job.h: one job over N devices."""
    from bcftools_b200 import mcall
    hdr = open(os.path.join(ROOT, "include", "mcall_job.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    syms = sorted(set(re.findall(r"\b(mcb_job_[a-z0-9_]+)\s*\(", hdr)))
    L = mcall.lib()
    assert syms == sorted(mcall.JOB_EXPORTS), syms
    assert all(hasattr(L, s) for s in syms)


def test_vcf_layer_symbols_are_exported():
    """include/b200_vcf.h (text VCF model) and include/b200_vcfcall.h (the `call -m` driver)."""
    from bcftools_b200 import vcfcall, mcall
    L = mcall.lib()
    for header, exports in (("b200_vcf.h", vcfcall.VCF_EXPORTS), ("b200_vcfcall.h", vcfcall.VCFCALL_EXPORTS), ("b200_bcfio.h", vcfcall.BCFIO_EXPORTS)):
        hdr = open(os.path.join(ROOT, "include", header)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        syms = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))
        assert syms == sorted(exports), (header, sorted(set(syms) ^ set(exports)))
        assert all(hasattr(L, s) for s in syms), [s for s in syms if not hasattr(L, s)]


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


def test_headers_are_plain_c99(tmp_path):
    """The boundary is a C ABI: every header under include/ must compile as C99 on its own (no C++-isms, no torch types)
    and a driver written against all of them must type-check."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    inc = os.path.join(ROOT, "include")
    hdrs = sorted(h for h in os.listdir(inc) if h.endswith(".h"))
    assert {"mcall_b200.h", "b200_call.h", "b200_bcf.h", "b200_driver.h"} <= set(hdrs)
    for h in hdrs:
        src = tmp_path / ("only_" + h.replace(".h", ".c"))
        src.write_text('#include "%s"\nint main(void) { return 0; }\n' % h)
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src)], check=True)
    prog = tmp_path / "driver.c"
    prog.write_text('''
#include <string.h>
#include "mcall_b200.h"
#include "b200_call.h"
#include "b200_bcf.h"
#include "b200_driver.h"
int run(const char *const *alleles, int n_allele, const uint8_t *indiv, size_t len, int nsmpl)
{
    b200_ploidy_t *pl = b200_ploidy_init_alias("GRCh38");
    int prev[8], s2s[4] = {0,1,0,1}; uint8_t ploidy[4];
    b200_call_t call; memset(&call, 0, sizeof call);
    call.nsmpl = nsmpl; call.bcf_typed = 1; call.ploidy = ploidy;
    b200_set_ploidy(pl, "chrX", 5000000, s2s, 4, prev, ploidy);
    call.unseen = (uint8_t) b200_unseen_allele(alleles, n_allele);
    b200_bcf_fmt_t fmt[4];
    if ( b200_bcf_unpack_fmt(indiv, len, 1, nsmpl, fmt) ) return -1;
    b200_rec_t rec; memset(&rec, 0, sizeof rec);
    rec.n_allele = n_allele; rec.PL_typed = fmt[0].p; rec.PL_bt = fmt[0].type; rec.nPLs = fmt[0].n*nsmpl;
    b200_mcall_init(&call);
    int ready = b200_mcall(&call, &rec);
    b200_out_t out;
    if ( ready ) b200_mcall_result(&call, 0, &out);
    b200_mcall_destroy(&call);
    b200_ploidy_destroy(pl);
    return mcb_version();
}
''')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(prog)], check=True)
