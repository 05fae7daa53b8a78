"""ctypes loaders for the two CPU oracles (TEST INFRASTRUCTURE ONLY).

  * `port`      : oracle/libmcall_oracle.so -- plain-C restatement (oracle/mcall_oracle.c)
  * `reference` : oracle/_ref/libmcall_ref.so -- the reference's unmodified mcall.c + in-memory htslib shim

Both expose  int f(const mcb_params*, const uint8_t *ploidy_tab, int nploidy, const mcb_batch*,
                    const mcb_result*, int site_beg, int site_end, double *secs).
"""
import ctypes as C
import os
import subprocess
import numpy as np

from bcftools_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libmcall_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmcall_ref.so")
REFERENCE_TREE = "/root/reference"


def build(want_ref=True):
    """Compile the oracles (building the checker is not using it).  The reference build needs
    /root/reference, which only exists in the build container; the GPU box uses the prebuilt .so."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if want_ref and os.path.exists(os.path.join(REFERENCE_TREE, "mcall.c")):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def have_ref():
    return os.path.exists(REF_SO)


_libs = {}


def _load(kind):
    if kind not in _libs:
        path, sym = (PORT_SO, "oracle_mcall_batch") if kind == "port" else (REF_SO, "ref_mcall_batch")
        if not os.path.exists(path):
            build(want_ref=(kind != "port"))
        lib = C.CDLL(path)
        fn = getattr(lib, sym)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(abi.McbParams), C.c_void_p, C.c_int, C.POINTER(abi.McbBatch),
                       C.POINTER(abi.McbResult), C.c_int, C.c_int, C.POINTER(C.c_double)]
        _libs[kind] = fn
    return _libs[kind]


def call(kind, params, batch, ploidy_tab=None, want_gp=False, site_range=None, result=None):
    """Run an oracle over a HostBatch.  Returns (HostResult, seconds spent in the per-record calls)."""
    fn = _load(kind)
    res = result if result is not None else abi.HostResult(batch, want_gp=want_gp)
    p, b, r = params.c_struct(), batch.c_struct(), res.c_struct()
    tab = None if ploidy_tab is None else np.ascontiguousarray(ploidy_tab, np.uint8).reshape(-1, params.nsmpl)
    secs = C.c_double(0)
    beg, end = site_range if site_range else (0, batch.nsites)
    rc = fn(C.byref(p), None if tab is None else tab.ctypes.data, 0 if tab is None else tab.shape[0],
            C.byref(b), C.byref(r), beg, end, C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"oracle {kind} failed: {rc}")
    return res, secs.value
