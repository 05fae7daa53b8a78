"""ctypes mirror of include/b200_bcf.h: BCF2 typed FORMAT vectors without htslib (host code of libmcall_b200.so)."""
import ctypes as C

import numpy as np

from . import mcall

BT_INT8, BT_INT16, BT_INT32, BT_FLOAT, BT_CHAR = 1, 2, 3, 5, 7
BCF_EXPORTS = ["b200_bcf_unpack_fmt", "b200_bcf_get_int", "b200_bcf_enc_int"]


class BcfFmt(C.Structure):
    _fields_ = [("key", C.c_int32), ("type", C.c_int32), ("n", C.c_int32), ("size", C.c_int32), ("p", C.c_void_p)]


class BcfError(ValueError):
    pass


def _lib():
    L = mcall.lib()
    if not getattr(L, "_bcf_ready", False):
        L.b200_bcf_unpack_fmt.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(BcfFmt)]
        L.b200_bcf_unpack_fmt.restype = C.c_int
        L.b200_bcf_get_int.argtypes = [C.POINTER(BcfFmt), C.c_int, C.c_int, C.c_void_p]
        L.b200_bcf_get_int.restype = C.c_int
        L.b200_bcf_enc_int.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.b200_bcf_enc_int.restype = C.c_int64
        L._bcf_ready = True
    return L


def enc_int(key, vals, n_sample):
    """Encode one integer FORMAT field from vals[n_sample][n] (int8/int16/int32) the way bcf_update_format_int32 would."""
    a = np.ascontiguousarray(vals).reshape(n_sample, -1)
    n = a.shape[1]
    out = np.zeros(16 + a.size * 4, np.uint8)
    k = _lib().b200_bcf_enc_int(out.ctypes.data, out.size, int(key), a.ctypes.data, a.dtype.itemsize, n, n_sample)
    if k < 0:
        raise BcfError("b200_bcf_enc_int: %d" % k)
    return out[:k].tobytes()


def unpack_fmt(indiv, n_fmt, n_sample):
    """FORMAT field headers of an indiv block: list of BcfFmt (p points into `indiv`, keep it alive)."""
    buf = np.frombuffer(indiv, np.uint8)
    fmt = (BcfFmt * n_fmt)()
    rc = _lib().b200_bcf_unpack_fmt(buf.ctypes.data, buf.size, n_fmt, n_sample, fmt)
    if rc:
        raise BcfError("b200_bcf_unpack_fmt: %d" % rc)
    return list(fmt), buf


def get_int(f, n_sample, dtype):
    """Integer vector of one field as int16 (the device's pl_type=2 layout) or int32 (bcf_get_format_int32)."""
    dt = np.dtype(dtype)
    out = np.zeros((n_sample, f.n), dt)
    rc = _lib().b200_bcf_get_int(C.byref(f), n_sample, dt.itemsize, out.ctypes.data)
    if rc < 0:
        raise BcfError("b200_bcf_get_int: %d" % rc)
    return out
