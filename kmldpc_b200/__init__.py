"""kmldpc_b200 — B200-native (sm_100a) Monte-Carlo link path of trganda/kmldpc behind a C ABI.

The product is kmldpc_b200/lib/libkmldpc_b200.so (CUDA kernels + host C++, include/kmldpc_b200.h); this package is
the thin ctypes mirror of the reference's objects.  There is no CPU fallback anywhere in this package."""
from .link import KmlError, LdpcCode, Link, Modem, pack_bits, snr_to_var, unpack_bits  # noqa: F401
from .simulator import Simulator  # noqa: F401

__all__ = ["KmlError", "LdpcCode", "Link", "Modem", "Simulator", "pack_bits", "unpack_bits", "snr_to_var"]
