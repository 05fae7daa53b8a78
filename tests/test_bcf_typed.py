"""include/b200_bcf.h: BCF2 typed FORMAT vectors without htslib -- known-answer bytes from the BCF2
specification's layout, htslib's type-selection thresholds, round trips, and slabs built from encoded records."""
import os
import re

import numpy as np
import pytest

from bcftools_b200 import abi, bcf_typed as bt, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I32_MISS, I32_EOV = np.iinfo(np.int32).min, np.iinfo(np.int32).min + 1


def test_symbols_are_exported():
    from bcftools_b200 import mcall
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "b200_bcf.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(b200_bcf_[a-z0-9_]+)\s*\(", hdr)))
    assert syms == sorted(bt.BCF_EXPORTS)
    assert all(hasattr(mcall.lib(), s) for s in syms)


def test_known_answer_bytes():
    # GT of 3 diploid samples 0/0, 0/1, ./. under key 1: key = typed int8 (0x11 0x01), 2 x int8 per sample (0x21)
    gt = np.array([[2, 2], [2, 4], [0, 0]], np.int32)
    assert bt.enc_int(1, gt, 3) == bytes([0x11, 0x01, 0x21, 2, 2, 2, 4, 0, 0])
    # haploid second value: int32 vector_end -> int8 vector_end 0x81; missing -> 0x80
    gt = np.array([[2, I32_EOV], [I32_MISS, I32_EOV]], np.int32)
    assert bt.enc_int(1, gt, 2) == bytes([0x11, 0x01, 0x21, 2, 0x81, 0x80, 0x81])
    # PL 0,30,255 needs int16 (255 > 127): descriptor 0x32, little endian
    pl = np.array([[0, 30, 255]], np.int32)
    assert bt.enc_int(5, pl, 1) == bytes([0x11, 0x05, 0x32, 0, 0, 30, 0, 255, 0])
    # 15 values per sample (5 alleles): length 15 overflows the nibble -> 0xF1 followed by the typed int 15 (0x11 0x0F)
    pl = np.arange(15, dtype=np.int32).reshape(1, 15)
    assert bt.enc_int(5, pl, 1) == bytes([0x11, 0x05, 0xF1, 0x11, 0x0F]) + bytes(range(15))
    # a key above 127 is a typed int16
    assert bt.enc_int(300, np.array([[1]], np.int32), 1)[:4] == bytes([0x12, 0x2C, 0x01, 0x11])


@pytest.mark.parametrize("lo,hi,want", [(-120, 127, bt.BT_INT8), (-121, 0, bt.BT_INT16), (0, 128, bt.BT_INT16),
                                        (-32760, 32767, bt.BT_INT16), (0, 32768, bt.BT_INT32), (-32761, 0, bt.BT_INT32)])
def test_type_selection_thresholds(lo, hi, want):
    """bcf_enc_vint: int8 for -120..127, int16 for -32760..32767, else int32; sentinels do not count."""
    v = np.array([[lo, hi, I32_MISS, I32_EOV]], np.int32)
    enc = bt.enc_int(2, v, 1)
    (f,), keep = bt.unpack_fmt(enc, 1, 1)
    assert (f.key, f.type, f.n) == (2, want, 4)
    assert (bt.get_int(f, 1, np.int32) == v).all()


@pytest.mark.parametrize("src", [np.int8, np.int16, np.int32])
def test_round_trip_of_device_typed_outputs(src):
    """gt8 / gq8 / pl16 as they leave mcb_call_host encode to the same bytes as their int32 widening."""
    rng = np.random.default_rng(3)
    info = np.iinfo(src)
    S, n = 37, 6
    v = rng.integers(0, 100, (S, n)).astype(src)
    v[rng.random((S, n)) < 0.1] = info.min
    v[:, -1][rng.random(S) < 0.3] = info.min + 1
    wide = v.astype(np.int32)
    wide[v == info.min] = I32_MISS
    wide[v == info.min + 1] = I32_EOV
    assert bt.enc_int(7, v, S) == bt.enc_int(7, wide, S)
    (f,), keep = bt.unpack_fmt(bt.enc_int(7, v, S), 1, S)
    assert (bt.get_int(f, S, np.int32) == wide).all()


def test_slab_from_encoded_records_equals_direct_slab():
    """A batcher that reads raw records: indiv = GT placeholder + PL + AD fields; the PL vector of every record,
    decoded straight to int16, must be the int16 slab HostBatch.to_int16() ships (and to int32, the original slab)."""
    params, hb, tab = synth.make_batch("C3", 24, with_groups=0)
    b16 = hb.to_int16()
    S = hb.nsmpl
    for i in range(hb.nsites):
        pl = hb.site_pl(i)
        G = pl.shape[1]
        other = np.full((S, 2), 2, np.int32)
        indiv = bt.enc_int(0, other, S) + bt.enc_int(9, pl, S) + bt.enc_int(11, np.zeros((S, int(hb.nals[i])), np.int32), S)
        fmt, keep = bt.unpack_fmt(indiv, 3, S)
        f = [x for x in fmt if x.key == 9][0]
        assert f.n == G and f.type in (bt.BT_INT8, bt.BT_INT16)
        assert (bt.get_int(f, S, np.int32) == pl).all()
        o = int(b16.pl_off[i])
        assert (bt.get_int(f, S, np.int16).ravel() == b16.pl[o:o + S * G]).all()


def test_malformed_input_is_reported():
    good = bt.enc_int(1, np.array([[1, 2, 3]], np.int32), 1)
    with pytest.raises(bt.BcfError):
        bt.unpack_fmt(good[:-1], 1, 1)                  # truncated data
    with pytest.raises(bt.BcfError):
        bt.unpack_fmt(bytes([0x15, 0, 0, 0, 0]) + good, 1, 1)   # key is not an integer
    (f,), keep = bt.unpack_fmt(bt.enc_int(1, np.array([[70000]], np.int32), 1), 1, 1)
    with pytest.raises(bt.BcfError):
        bt.get_int(f, 1, np.int16)                      # does not fit int16
