"""ctypes mirror of include/b200_call.h -- the C host batcher with the reference's hook names
(mcall_init / mcall / mcall_destroy, call.h:131-147).  Used by the tests that replay records one by one."""
import ctypes as C

import numpy as np

from . import mcall


class B200Call(C.Structure):
    _fields_ = [("nsmpl", C.c_int), ("flag", C.c_uint32), ("output_tags", C.c_uint32), ("theta", C.c_double),
                ("ploidy", C.c_void_p), ("unseen", C.c_uint8), ("nsmpl_grp", C.c_int),
                ("grp_off", C.c_void_p), ("grp_smpl", C.c_void_p), ("use_prior", C.c_int),
                ("max_records", C.c_int), ("max_nals", C.c_int), ("device", C.c_int), ("batcher", C.c_void_p)]


class B200Rec(C.Structure):
    _fields_ = [("n_allele", C.c_int), ("PLs", C.c_void_p), ("nPLs", C.c_int), ("QS", C.c_void_p), ("nQS", C.c_int),
                ("ADs", C.c_void_p), ("nADs", C.c_int), ("prior_an", C.c_int32), ("prior_ac", C.c_void_p),
                ("n_prior_ac", C.c_int), ("user", C.c_void_p)]


class B200Out(C.Structure):
    _fields_ = [("ret", C.c_int), ("als_new", C.c_uint32), ("als_map", C.POINTER(C.c_int8)), ("qual", C.c_float),
                ("ac", C.POINTER(C.c_int32)), ("an", C.c_int), ("site_flags", C.c_uint32),
                ("gts", C.POINTER(C.c_int32)), ("GQs", C.POINTER(C.c_int32)), ("PLs", C.POINTER(C.c_int32)), ("nPLs", C.c_int),
                ("GPs", C.POINTER(C.c_float)), ("user", C.c_void_p)]


HOST_EXPORTS = ["b200_mcall_init", "b200_mcall", "b200_mcall_flush", "b200_mcall_result", "b200_mcall_destroy",
                "b200_set_error_handler"]


def _lib():
    L = mcall.lib()
    L.b200_mcall_init.argtypes = [C.POINTER(B200Call)]
    L.b200_mcall_init.restype = None
    L.b200_mcall.argtypes = [C.POINTER(B200Call), C.POINTER(B200Rec)]
    L.b200_mcall.restype = C.c_int
    L.b200_mcall_flush.argtypes = [C.POINTER(B200Call)]
    L.b200_mcall_flush.restype = C.c_int
    L.b200_mcall_result.argtypes = [C.POINTER(B200Call), C.c_int, C.POINTER(B200Out)]
    L.b200_mcall_result.restype = C.c_int
    L.b200_mcall_destroy.argtypes = [C.POINTER(B200Call)]
    L.b200_mcall_destroy.restype = None
    return L


def replay(params, batch, ploidy_tab=None, max_records=64):
    """Feed a HostBatch record by record through b200_mcall (like vcfcall.c:1089-1148 feeds mcall) and collect the
    results into an abi.HostResult laid out like the C-ABI's, so that the usual comparison helpers apply."""
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
            res.als_new[i], res.qual[i], res.an[i] = out.als_new, out.qual, out.an
            res.als_map[i] = np.ctypeslib.as_array(out.als_map, (params.max_nals,))
            res.ac[i] = np.ctypeslib.as_array(out.ac, (params.max_nals,))
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
        n = L.b200_mcall_flush(C.byref(call))
        if n:
            collect(n)
    finally:
        L.b200_mcall_destroy(C.byref(call))
    assert done[0] == batch.nsites
    return res
