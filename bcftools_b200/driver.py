"""ctypes mirror of include/b200_driver.h: ploidy definitions, -G group files, the unseen allele (host code of libmcall_b200.so)."""
import ctypes as C

import numpy as np

from . import mcall

DRIVER_EXPORTS = ["b200_ploidy_init_string", "b200_ploidy_init_alias", "b200_ploidy_destroy", "b200_ploidy_add_sex", "b200_ploidy_nsex", "b200_ploidy_sex2id",
                  "b200_ploidy_id2sex", "b200_ploidy_min", "b200_ploidy_max", "b200_ploidy_query", "b200_set_ploidy",
                  "b200_groups_parse", "b200_unseen_allele", "b200_samples_parse", "b200_samples_default",
                  "b200_trim_numberR", "b200_i16_to_dp4_mq"]


class DriverError(ValueError):
    pass


def _lib():
    L = mcall.lib()
    if not getattr(L, "_drv_ready", False):
        L.b200_ploidy_init_string.argtypes = [C.c_char_p, C.c_int]
        L.b200_ploidy_init_string.restype = C.c_void_p
        L.b200_ploidy_init_alias.argtypes = [C.c_char_p]
        L.b200_ploidy_init_alias.restype = C.c_void_p
        L.b200_ploidy_destroy.argtypes = [C.c_void_p]
        L.b200_ploidy_destroy.restype = None
        L.b200_ploidy_add_sex.argtypes = [C.c_void_p, C.c_char_p]
        L.b200_ploidy_sex2id.argtypes = [C.c_void_p, C.c_char_p]
        for f in ("nsex", "min", "max"):
            getattr(L, "b200_ploidy_" + f).argtypes = [C.c_void_p]
        L.b200_ploidy_id2sex.argtypes = [C.c_void_p, C.c_int]
        L.b200_ploidy_id2sex.restype = C.c_char_p
        L.b200_ploidy_query.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.b200_set_ploidy.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b200_groups_parse.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                        C.c_char_p, C.c_size_t]
        L.b200_unseen_allele.argtypes = [C.POINTER(C.c_char_p), C.c_int]
        L.b200_samples_parse.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
        L.b200_samples_default.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b200_samples_default.restype = None
        L.b200_trim_numberR.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.b200_i16_to_dp4_mq.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        L.b200_i16_to_dp4_mq.restype = None
        L._drv_ready = True
    return L


class Ploidy:
    """ploidy_t of ploidy.c: definitions "CHROM FROM TO SEX PLOIDY" and per-position queries."""

    def __init__(self, text=None, dflt=2, alias=None):
        if alias is not None:
            self._p = _lib().b200_ploidy_init_alias(alias.encode())
            if not self._p:
                raise DriverError("no such ploidy alias: %s" % alias)
            return
        self._p = _lib().b200_ploidy_init_string(text.encode(), dflt)
        if not self._p:
            raise DriverError("could not parse the ploidy definition")

    def close(self):
        if self._p:
            _lib().b200_ploidy_destroy(self._p)
            self._p = None

    def __del__(self):
        self.close()

    def add_sex(self, sex):
        return _lib().b200_ploidy_add_sex(self._p, sex.encode())

    @property
    def sexes(self):
        L = _lib()
        return [L.b200_ploidy_id2sex(self._p, i).decode() for i in range(L.b200_ploidy_nsex(self._p))]

    def sex2id(self, sex):
        return _lib().b200_ploidy_sex2id(self._p, sex.encode())

    def min(self):
        return _lib().b200_ploidy_min(self._p)

    def max(self):
        return _lib().b200_ploidy_max(self._p)

    def query(self, seq, pos):
        """(hit, {sex: ploidy}, min, max) at 0-based pos."""
        L = _lib()
        s2p = np.zeros(max(1, L.b200_ploidy_nsex(self._p)), np.int32)
        mn, mx = C.c_int(), C.c_int()
        hit = L.b200_ploidy_query(self._p, seq.encode(), pos, s2p.ctypes.data, C.byref(mn), C.byref(mx))
        return hit, dict(zip(self.sexes, s2p.tolist())), mn.value, mx.value

    def samples_parse(self, text, hdr_samples):
        """-S file content -> (samples_map, sample2sex, warnings); new sexes are added to this definition (vcfcall.c:270-344)."""
        n = len(hdr_samples)
        arr = (C.c_char_p * max(1, n))(*[s.encode() for s in hdr_samples])
        smap, s2s = np.zeros(max(1, n), np.int32), np.zeros(max(1, n), np.int32)
        nsel, nwarn = C.c_int(), C.c_int()
        err = C.create_string_buffer(512)
        rc = _lib().b200_samples_parse(text.encode(), arr, n, self._p, smap.ctypes.data, s2s.ctypes.data, C.byref(nsel), C.byref(nwarn), err, len(err))
        if rc:
            raise DriverError("%d: %s" % (rc, err.value.decode()))
        return smap[:nsel.value].copy(), s2s[:nsel.value].copy(), nwarn.value

    def samples_default(self, nhdr):
        smap, s2s = np.zeros(nhdr, np.int32), np.zeros(nhdr, np.int32)
        _lib().b200_samples_default(self._p, nhdr, smap.ctypes.data, s2s.ctypes.data)
        return smap, s2s

    def set_ploidy(self, seq, pos, sample2sex, prev, ploidy):
        """vcfcall.c:807-825 on numpy state arrays (int32 sample2sex, int32 prev[nsex], uint8 ploidy[nsmpl])."""
        return _lib().b200_set_ploidy(self._p, seq.encode(), pos, sample2sex.ctypes.data, len(sample2sex), prev.ctypes.data, ploidy.ctypes.data)


def groups_parse(text, samples):
    """(grp_off, grp_smpl) as mcb_params wants them, from a -G file's content or "-"."""
    n = len(samples)
    arr = (C.c_char_p * max(1, n))(*[s.encode() for s in samples])
    off, smp = np.zeros(n + 1, np.uint32), np.zeros(max(1, n), np.uint32)
    ng = C.c_int()
    err = C.create_string_buffer(512)
    rc = _lib().b200_groups_parse(text.encode(), arr, n, off.ctypes.data, smp.ctypes.data, C.byref(ng), err, len(err))
    if rc:
        raise DriverError("%d: %s" % (rc, err.value.decode()))
    return off[:ng.value + 1].copy(), smp[:n].copy()


def unseen_allele(alleles):
    arr = (C.c_char_p * len(alleles))(*[a.encode() for a in alleles])
    return _lib().b200_unseen_allele(arr, len(alleles))


def trim_numberR(vals, als_map, nals_new):
    """vals[nvec][nals_ori] (int32 or float32) -> [nvec][nals_new] under als_map (mcall.c:1196-1265)."""
    a = np.ascontiguousarray(vals)
    assert a.dtype.itemsize == 4
    a = a.reshape(-1, a.shape[-1])
    out = np.zeros((a.shape[0], nals_new), a.dtype)
    m = np.ascontiguousarray(als_map, np.int8)
    _lib().b200_trim_numberR(a.ctypes.data, out.ctypes.data, a.shape[0], a.shape[1], nals_new, m.ctypes.data)
    return out


def i16_to_dp4_mq(i16):
    a = np.ascontiguousarray(i16, np.float32)
    dp4, mq = np.zeros(4, np.int32), C.c_int32()
    _lib().b200_i16_to_dp4_mq(a.ctypes.data, dp4.ctypes.data, C.byref(mq))
    return dp4.tolist(), mq.value
