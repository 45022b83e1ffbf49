"""consent_b200 — B200-native (sm_100a) implementation of CONSENT's per-window correction hot path.

See DESIGN.md.  The product is the CUDA library `libconsent_b200.so` behind the C ABI of
`include/consent_b200.h`; this package is the thin Python host side (ctypes) used by tests and bench.
"""
from ._ffi import Batch, Params, Results  # noqa: F401
