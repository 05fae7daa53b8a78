"""include/b200_vcf.h: the text VCF reader / writer without htslib.  Every VCF of the reference's `call -m` tests -- inputs
and expected outputs, all written by htslib -- must parse and re-format to the same bytes, and every float in them must
survive float32 -> b200_str_putd (htslib's kputd) unchanged."""
import ctypes as C
import re
import struct

import numpy as np

from bcftools_b200 import mcall, vcfcall
from tests import vcf_cases


def _lib():
    L = mcall.lib()
    vcfcall._lib()
    L.b200_vhdr_parse.restype = C.c_void_p
    L.b200_vhdr_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.b200_vhdr_destroy.argtypes = [C.c_void_p]
    L.b200_vhdr_format.argtypes = [C.c_void_p, C.POINTER(vcfcall.B200Str)]
    L.b200_vrec_parse.restype = C.c_void_p
    L.b200_vrec_parse.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.b200_vrec_format.argtypes = [C.c_void_p, C.POINTER(vcfcall.B200Str)]
    L.b200_vrec_destroy.argtypes = [C.c_void_p]
    L.b200_vrec_fmt_ints.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int)]
    L.b200_vhdr_subset.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    return L


def test_round_trip_of_every_reference_vcf():
    L = _lib()
    files = vcf_cases.bundle()["files"]
    nrec = 0
    for name, text in sorted(files.items()):
        if not (name.endswith(".vcf") or name.endswith(".out")):
            continue
        data = text.encode("latin-1")
        used = C.c_size_t(0)
        h = L.b200_vhdr_parse(data, len(data), C.byref(used))
        assert h, name
        out = vcfcall.B200Str(None, 0, 0)
        L.b200_vhdr_format(h, C.byref(out))
        for line in data[used.value:].split(b"\n"):
            if not line:
                continue
            r = L.b200_vrec_parse(h, line, len(line))
            assert r, (name, line[:80])
            L.b200_vrec_format(r, C.byref(out))
            L.b200_vrec_destroy(r)
            nrec += 1
        assert C.string_at(out.s, out.l) == data, name
        L.b200_vhdr_destroy(h)
    assert nrec > 9000


def test_float_formatting_is_a_fixed_point_on_htslib_output():
    files = vcf_cases.bundle()["files"]
    tok = set()
    for name, text in files.items():
        if name.endswith(".vcf") or name.endswith(".out"):
            tok.update(re.findall(r"(?<![\w.])-?\d+\.\d+(?:e[-+]?\d+)?|(?<![\w.])-?\d+e[-+]?\d+", text))
    assert len(tok) > 1000
    for t in sorted(tok):
        v = struct.unpack("f", struct.pack("f", float(t)))[0]
        assert vcfcall.format_float(v) == t, (t, vcfcall.format_float(v))
    for v, s in ((0.0, "0"), (-0.0, "-0"), (1e-5, "1e-05"), (1234567.0, "1.23457e+06"), (999999.0, "999999"), (0.5, "0.5"),
                 (100000.0, "100000"), (0.000123456789, "0.000123457"), (59.5765, "59.5765")):
        assert vcfcall.format_float(v) == s, (v, vcfcall.format_float(v), s)


def test_format_vectors_follow_the_longest_input_sample_and_pad_with_vector_end():
    """what bcf_get_format_int32 returns: '.' = missing then vector_end; a subset keeps the width of the full record"""
    L = _lib()
    text = (b"##fileformat=VCFv4.2\n##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"x\">\n"
            b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\tc\n")
    line = b"1\t5\t.\tA\tC\t.\t.\t.\tPL:DP\t0,3,30:4\t.:1\t7,8"
    used = C.c_size_t(0)
    h = L.b200_vhdr_parse(text, len(text), C.byref(used))
    r = L.b200_vrec_parse(h, line, len(line))
    dst, m = C.POINTER(C.c_int32)(), C.c_int(0)
    n = L.b200_vrec_fmt_ints(r, b"PL", C.byref(dst), C.byref(m))
    MISS, END = -2**31, -2**31 + 1
    assert n == 9 and [dst[i] for i in range(9)] == [0, 3, 30, MISS, END, END, 7, 8, END]
    n = L.b200_vrec_fmt_ints(r, b"DP", C.byref(dst), C.byref(m))
    assert n == 3 and [dst[i] for i in range(3)] == [4, 1, MISS]         # dropped trailing field = missing
    L.b200_vrec_destroy(r)
    sel = (C.c_int * 2)(2, 1)
    assert L.b200_vhdr_subset(h, 2, sel) == 0
    r = L.b200_vrec_parse(h, line, len(line))
    n = L.b200_vrec_fmt_ints(r, b"PL", C.byref(dst), C.byref(m))
    assert n == 6 and [dst[i] for i in range(6)] == [7, 8, END, MISS, END, END]
    out = vcfcall.B200Str(None, 0, 0)
    L.b200_vrec_format(r, C.byref(out))
    assert C.string_at(out.s, out.l) == b"1\t5\t.\tA\tC\t.\t.\t.\tPL:DP\t7,8:.\t.:1\n"
    L.b200_vrec_destroy(r)
    L.b200_vhdr_destroy(h)
