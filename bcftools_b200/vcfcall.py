"""ctypes mirror of include/b200_vcfcall.h -- the `bcftools call -m` driver around the B200 path (text VCF in and out,
htslib-free).  `run()` is the whole command on the device; `VcfCall` exposes the two host halves (everything in front
of / behind the likelihood code) so that tests can replay them."""
import ctypes as C

import numpy as np

from . import abi, mcall
from .host_call import B200Call, B200Rec, B200Out

VCFCALL_EXPORTS = ["b200_vc_open", "b200_vc_close", "b200_vc_error", "b200_vc_call_params", "b200_vc_ploidy", "b200_vc_unseen", "b200_vc_output_type",
                   "b200_vc_next", "b200_vc_finish", "b200_vc_flush", "b200_vc_output", "b200_vc_output_clear",
                   "b200_vcfcall_run", "b200_pv4", "b200_vcmp_set_ref", "b200_vcmp_find_allele"]
VCF_EXPORTS = ["b200_str_putsn", "b200_str_puts", "b200_str_putc", "b200_str_putw", "b200_str_putd",
               "b200_vhdr_parse", "b200_vhdr_destroy", "b200_vhdr_def", "b200_vhdr_append", "b200_vhdr_remove", "b200_vhdr_subset",
               "b200_vhdr_format", "b200_vrec_parse", "b200_vrec_new", "b200_vrec_destroy", "b200_vrec_format", "b200_vrec_info",
               "b200_vrec_info_floats", "b200_vrec_info_ints", "b200_vrec_fmt_ints", "b200_vrec_fmt", "b200_vrec_set_info_ints",
               "b200_vrec_set_info_floats", "b200_vrec_set_info_text", "b200_vrec_set_fmt_ints", "b200_vrec_set_fmt_floats",
               "b200_vrec_set_genotypes", "b200_vrec_set_alleles"]


BCFIO_EXPORTS = ["b200_bcfdict_build", "b200_bcfdict_destroy", "b200_bgzf_compress", "b200_bgzf_finish", "b200_bgzf_decompress",
                 "b200_bcf_write_header", "b200_bcf_read_header", "b200_bcf_encode_rec", "b200_bcf_decode_rec",
                 "b200_vcf_text_to_bcf", "b200_bcf_to_vcf_text"]


class B200Str(C.Structure):
    _fields_ = [("s", C.c_void_p), ("l", C.c_size_t), ("m", C.c_size_t)]


def _lib():
    L = mcall.lib()
    L.b200_vc_open.restype = C.c_void_p
    L.b200_vc_open.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    L.b200_vc_close.argtypes = [C.c_void_p]
    L.b200_vc_close.restype = None
    L.b200_vc_error.argtypes = [C.c_void_p]
    L.b200_vc_error.restype = C.c_char_p
    L.b200_vc_call_params.argtypes = [C.c_void_p, C.POINTER(B200Call)]
    L.b200_vc_call_params.restype = None
    L.b200_vc_ploidy.argtypes = [C.c_void_p]
    L.b200_vc_ploidy.restype = C.POINTER(C.c_uint8)
    L.b200_vc_unseen.argtypes = [C.c_void_p]
    L.b200_vc_next.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(B200Rec)]
    L.b200_vc_finish.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(B200Out)]
    L.b200_vc_flush.argtypes = [C.c_void_p]
    L.b200_vc_output.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    L.b200_vc_output.restype = C.c_void_p
    L.b200_vc_output_clear.argtypes = [C.c_void_p]
    L.b200_vc_output_clear.restype = None
    L.b200_vcfcall_run.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    L.b200_pv4.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.b200_str_putd.argtypes = [C.POINTER(B200Str), C.c_double]
    return L


def _argv(args):
    arr = (C.c_char_p * max(len(args), 1))()
    for i, a in enumerate(args):
        arr[i] = a.encode()
    return arr


def run(args, in_path, out_path, device=0):
    """`bcftools call --no-version -O v <args> in_path > out_path`, the calling done on the device."""
    L = _lib()
    err = C.create_string_buffer(1024)
    rc = L.b200_vcfcall_run(len(args), _argv(args), in_path.encode(), out_path.encode(), device, err, len(err))
    if rc:
        raise RuntimeError("b200_vcfcall_run: " + err.value.decode(errors="replace"))


def pv4(i16):
    a = (C.c_float * 16)(*[float(x) for x in i16])
    p = (C.c_float * 4)()
    tested = _lib().b200_pv4(a, p)
    return tested, [float(x) for x in p]


def vcf_to_bcf(text, level=6):
    """text VCF -> BCF2.2 bytes (BGZF, include/b200_bcfio.h)"""
    L = _lib()
    L.b200_vcf_text_to_bcf.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(B200Str)]
    out = B200Str(None, 0, 0)
    if L.b200_vcf_text_to_bcf(text, len(text), level, C.byref(out)):
        raise RuntimeError("b200_vcf_text_to_bcf failed")
    return C.string_at(out.s, out.l)


def bcf_to_vcf(data):
    L = _lib()
    L.b200_bcf_to_vcf_text.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(B200Str)]
    out = B200Str(None, 0, 0)
    if L.b200_bcf_to_vcf_text(data, len(data), C.byref(out)):
        raise RuntimeError("b200_bcf_to_vcf_text failed")
    return C.string_at(out.s, out.l)


def format_float(x):
    s = B200Str(None, 0, 0)
    _lib().b200_str_putd(C.byref(s), float(x))
    return C.string_at(s.s, s.l).decode()


class VcfCall:
    """The host halves of the driver: `next()` yields (handle, record inputs as numpy arrays), `finish()` takes the
    caller's result for that record."""

    def __init__(self, args, vcf_text):
        self.L = _lib()
        self._text = vcf_text if isinstance(vcf_text, bytes) else vcf_text.encode()
        err = C.create_string_buffer(1024)
        self.vc = self.L.b200_vc_open(len(args), _argv(args), self._text, len(self._text), err, len(err))
        if not self.vc:
            raise RuntimeError(err.value.decode(errors="replace"))
        self.call = B200Call()
        self.L.b200_vc_call_params(self.vc, C.byref(self.call))
        self.nsmpl = self.call.nsmpl

    def close(self):
        if self.vc:
            self.L.b200_vc_close(self.vc)
            self.vc = None

    def groups(self):
        n = self.call.nsmpl_grp
        if n <= 1:
            return None
        off = np.ctypeslib.as_array(C.cast(self.call.grp_off, C.POINTER(C.c_uint32)), (n + 1,)).copy()
        smpl = np.ctypeslib.as_array(C.cast(self.call.grp_smpl, C.POINTER(C.c_uint32)), (self.nsmpl,)).copy()
        return [smpl[off[i]:off[i + 1]].tolist() for i in range(n)]

    def ploidy(self):
        return np.ctypeslib.as_array(self.L.b200_vc_ploidy(self.vc), (max(self.nsmpl, 1),))[:self.nsmpl].copy()

    def next(self):
        h, rec = C.c_void_p(), B200Rec()
        rc = self.L.b200_vc_next(self.vc, C.byref(h), C.byref(rec))
        if rc < 0:
            raise RuntimeError(self.L.b200_vc_error(self.vc).decode(errors="replace"))
        if rc == 0:
            return None
        S, A = self.nsmpl, rec.n_allele
        d = dict(handle=h, n_allele=A, unseen=self.L.b200_vc_unseen(self.vc), ploidy=self.ploidy(),
                 pl=np.ctypeslib.as_array(C.cast(rec.PLs, C.POINTER(C.c_int32)), (rec.nPLs,)).copy().reshape(S, -1),
                 qs=None, ad=None, prior_an=rec.prior_an, prior_ac=None)
        if rec.nQS > 0:
            d["qs"] = np.ctypeslib.as_array(C.cast(rec.QS, C.POINTER(C.c_float)), (rec.nQS,)).copy()
        if rec.nADs > 0:
            d["ad"] = np.ctypeslib.as_array(C.cast(rec.ADs, C.POINTER(C.c_int32)), (rec.nADs,)).copy().reshape(S, -1)
        if rec.n_prior_ac > 0:
            d["prior_ac"] = np.ctypeslib.as_array(C.cast(rec.prior_ac, C.POINTER(C.c_int32)), (rec.n_prior_ac,)).copy()
        return d

    def finish(self, handle, res, i=0):
        """res: abi.HostResult holding the record's result at site i (int32 layouts)."""
        o = B200Out()
        S = self.nsmpl
        keep = []

        def p(a, t):
            a = np.ascontiguousarray(a)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(t))
        o.ret, o.als_new, o.qual, o.an, o.site_flags = int(res.ret[i]), int(res.als_new[i]), float(res.qual[i]), int(res.an[i]), int(res.site_flags[i])
        qual = np.array([res.qual[i]], np.float32)
        C.memmove(C.byref(o, B200Out.qual.offset), qual.ctypes.data, 4)     # keep the exact bits (missing is a NaN pattern)
        o.als_map, o.ac = p(res.als_map[i], C.c_int8), p(res.ac[i], C.c_int32)
        o.gts = p(res.gt[i].reshape(-1), C.c_int32)
        if res.gq is not None:
            o.GQs = p(res.gq[i], C.c_int32)
        if o.ret > 0 and not (o.site_flags & abi.SITE_PL_DROPPED):
            pl = res.site_pl(i).reshape(-1)
            o.PLs, o.nPLs = p(pl, C.c_int32), pl.size
            if res.gp is not None:
                o.GPs = p(res.site_gp(i).reshape(-1), C.c_float)
        rc = self.L.b200_vc_finish(self.vc, handle, C.byref(o))
        if rc < 0:
            raise RuntimeError(self.L.b200_vc_error(self.vc).decode(errors="replace"))

    def flush(self):
        if self.L.b200_vc_flush(self.vc) < 0:
            raise RuntimeError(self.L.b200_vc_error(self.vc).decode(errors="replace"))

    def output(self):
        n = C.c_size_t(0)
        p = self.L.b200_vc_output(self.vc, C.byref(n))
        return C.string_at(p, n.value) if n.value else b""
