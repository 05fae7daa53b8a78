"""bcftools_b200 -- B200-native implementation of the `bcftools call -m` hot path (mcall.c).

The product is the CUDA library bcftools_b200/lib/libmcall_b200.so behind the C-ABI of
include/mcall_b200.h; this package only holds its sources (csrc/), the in-tree build recipe and a
thin ctypes mirror used by tests and bench.py.  There is no CPU fallback.
"""
from . import abi  # noqa: F401
