"""ctypes mirror of include/b200_call.h -- the C host batcher with the reference's hook names
(mcall_init / mcall / mcall_destroy, call.h:131-147).  Used by the tests that replay records one by one."""
import ctypes as C

import numpy as np

from . import mcall


class B200Call(C.Structure):
    _fields_ = [("nsmpl", C.c_int), ("flag", C.c_uint32), ("output_tags", C.c_uint32), ("theta", C.c_double),
                ("ploidy", C.c_void_p), ("unseen", C.c_uint8), ("nsmpl_grp", C.c_int),
                ("grp_off", C.c_void_p), ("grp_smpl", C.c_void_p), ("use_prior", C.c_int),
                ("max_records", C.c_int), ("max_nals", C.c_int), ("device", C.c_int), ("bcf_typed", C.c_int),
                ("async_flush", C.c_int), ("tie_eps", C.c_double), ("batcher", C.c_void_p)]


class B200Rec(C.Structure):
    _fields_ = [("n_allele", C.c_int), ("PLs", C.c_void_p), ("nPLs", C.c_int), ("QS", C.c_void_p), ("nQS", C.c_int),
                ("ADs", C.c_void_p), ("nADs", C.c_int), ("prior_an", C.c_int32), ("prior_ac", C.c_void_p),
                ("n_prior_ac", C.c_int), ("user", C.c_void_p), ("PL_typed", C.c_void_p), ("PL_bt", C.c_int)]


class B200Out(C.Structure):
    _fields_ = [("ret", C.c_int), ("als_new", C.c_uint32), ("als_map", C.POINTER(C.c_int8)), ("qual", C.c_float),
                ("ac", C.POINTER(C.c_int32)), ("an", C.c_int), ("site_flags", C.c_uint32),
                ("gts", C.POINTER(C.c_int32)), ("GQs", C.POINTER(C.c_int32)), ("PLs", C.POINTER(C.c_int32)), ("nPLs", C.c_int),
                ("GPs", C.POINTER(C.c_float)), ("user", C.c_void_p),
                ("gts8", C.POINTER(C.c_int8)), ("GQs8", C.POINTER(C.c_int8)), ("PLs16", C.POINTER(C.c_int16))]


HOST_EXPORTS = ["b200_mcall_init", "b200_mcall", "b200_mcall_flush", "b200_mcall_flush_async", "b200_mcall_wait", "b200_mcall_result",
                "b200_mcall_n_ploidy", "b200_mcall_destroy", "b200_set_error_handler"]


def _lib():
    L = mcall.lib()
    L.b200_mcall_init.argtypes = [C.POINTER(B200Call)]
    L.b200_mcall_init.restype = None
    L.b200_mcall.argtypes = [C.POINTER(B200Call), C.POINTER(B200Rec)]
    L.b200_mcall.restype = C.c_int
    for f in (L.b200_mcall_flush, L.b200_mcall_flush_async, L.b200_mcall_wait, L.b200_mcall_n_ploidy):
        f.argtypes = [C.POINTER(B200Call)]
        f.restype = C.c_int
    L.b200_mcall_result.argtypes = [C.POINTER(B200Call), C.c_int, C.POINTER(B200Out)]
    L.b200_mcall_result.restype = C.c_int
    L.b200_mcall_destroy.argtypes = [C.POINTER(B200Call)]
    L.b200_mcall_destroy.restype = None
    return L


def _widen(a, bits):
    lo = -(1 << (bits - 1))
    o = a.astype(np.int32)
    o[a == lo] = np.iinfo(np.int32).min
    o[a == lo + 1] = np.iinfo(np.int32).min + 1
    return o


def replay(params, batch, ploidy_tab=None, max_records=64, typed=False, async_flush=False, stats=None, tie_eps=0.0):
    """Feed a HostBatch record by record through b200_mcall (like vcfcall.c:1089-1148 feeds mcall) and collect the
    results into an abi.HostResult laid out like the C-ABI's, so that the usual comparison helpers apply.
    typed=True: b200_call_t.bcf_typed -- PL goes in as the int8/int16 typed vector a BCF record would hold (int8 when
    every value fits, like bcf_enc_vint chooses) and GT / GQ / PL come back as int8 / int8 / int16 vectors."""
    from . import abi
    L = _lib()
    S = params.nsmpl
    call = B200Call()
    call.nsmpl, call.flag, call.output_tags, call.theta = S, params.flag, params.output_tags, params.theta
    ploidy = np.full(S, 2, np.uint8)
    call.ploidy = ploidy.ctypes.data
    call.nsmpl_grp = params.ngroups
    if params.ngroups > 1:
        call.grp_off, call.grp_smpl = params.grp_off.ctypes.data, params.grp_smpl.ctypes.data
    call.use_prior, call.max_records, call.max_nals, call.device = int(params.use_prior), max_records, params.max_nals, params.device
    call.bcf_typed = int(typed)
    call.async_flush = int(async_flush)
    call.tie_eps = float(tie_eps)
    L.b200_mcall_init(C.byref(call))
    res = abi.HostResult(batch, want_gp=bool(params.output_tags & abi.CALL_FMT_GP))
    done = [0]

    def collect(n):
        out = B200Out()
        for k in range(n):
            assert L.b200_mcall_result(C.byref(call), k, C.byref(out)) == 0
            i = done[0] + k
            res.ret[i], res.site_flags[i] = out.ret, out.site_flags
            if out.ret <= 0:
                continue
            res.als_new[i], res.an[i] = out.als_new, out.an
            res.qual.view(np.uint32)[i] = C.c_uint32.from_buffer(out, B200Out.qual.offset).value      # the exact bits: missing QUAL is a signalling NaN pattern
            res.als_map[i] = np.ctypeslib.as_array(out.als_map, (params.max_nals,))
            res.ac[i] = np.ctypeslib.as_array(out.ac, (params.max_nals,))
            if typed:
                assert not out.gts and not out.PLs
                res.gt[i] = _widen(np.ctypeslib.as_array(out.gts8, (S, 2)), 8)
                if out.GQs8:
                    res.gq[i] = _widen(np.ctypeslib.as_array(out.GQs8, (S,)), 8)
                if out.PLs16:
                    o = batch.pl_off[i]
                    res.pl[o:o + out.nPLs] = _widen(np.ctypeslib.as_array(out.PLs16, (out.nPLs,)), 16)
                    if out.GPs and res.gp is not None:
                        res.gp[o:o + out.nPLs] = np.ctypeslib.as_array(out.GPs, (out.nPLs,))
                continue
            res.gt[i] = np.ctypeslib.as_array(out.gts, (S, 2))
            if out.GQs:
                res.gq[i] = np.ctypeslib.as_array(out.GQs, (S,))
            if out.PLs:
                o = batch.pl_off[i]
                res.pl[o:o + out.nPLs] = np.ctypeslib.as_array(out.PLs, (out.nPLs,))
                if out.GPs and res.gp is not None:
                    res.gp[o:o + out.nPLs] = np.ctypeslib.as_array(out.GPs, (out.nPLs,))
        done[0] += n

    try:
        for i in range(batch.nsites):
            if ploidy_tab is not None and batch.ploidy_id is not None:
                ploidy[:] = np.asarray(ploidy_tab, np.uint8).reshape(-1, S)[batch.ploidy_id[i]]     # set_ploidy()
            call.unseen = int(batch.unseen[i])
            rec = B200Rec()
            pl = np.ascontiguousarray(batch.site_pl(i))
            rec.n_allele, rec.PLs, rec.nPLs = int(batch.nals[i]), pl.ctypes.data, pl.size
            keep = [pl]
            if typed:
                real = pl[pl > np.iinfo(np.int32).min + 1]
                bits = 8 if real.size == 0 or (real.max() <= 127 and real.min() >= -120) else 16
                lo = -(1 << (bits - 1))
                nar = pl.astype(np.int8 if bits == 8 else np.int16)
                nar[pl == np.iinfo(np.int32).min] = lo
                nar[pl == np.iinfo(np.int32).min + 1] = lo + 1
                rec.PLs, rec.PL_typed, rec.PL_bt = None, nar.ctypes.data, bits // 8
                keep.append(nar)
            if params.ngroups > 1:
                ad = np.ascontiguousarray(batch.site_ad(i))
                rec.ADs, rec.nADs = ad.ctypes.data, ad.size
                keep.append(ad)
            else:
                nq = int(batch.nqs[i]) if batch.nqs is not None else int(batch.nals[i])
                q = np.ascontiguousarray(batch.qs[i][:max(nq, 1)])
                rec.QS, rec.nQS = q.ctypes.data, nq
                keep.append(q)
            rec.prior_an = abi.INT32_MISSING
            if batch.prior_an is not None:
                rec.prior_an = int(batch.prior_an[i])
                pac = np.ascontiguousarray(batch.prior_ac[i])
                rec.prior_ac, rec.n_prior_ac = pac.ctypes.data, int(batch.nals[i]) - 1
                keep.append(pac)
            n = L.b200_mcall(C.byref(call), C.byref(rec))
            if n:
                collect(n)
        while True:         # drains the batch in flight (async) and the partial last batch
            n = L.b200_mcall_flush(C.byref(call))
            if n <= 0:
                break
            collect(n)
        if stats is not None:
            stats["n_ploidy"] = L.b200_mcall_n_ploidy(C.byref(call))
    finally:
        L.b200_mcall_destroy(C.byref(call))
    assert done[0] == batch.nsites
    return res
