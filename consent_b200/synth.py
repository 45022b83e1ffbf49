"""Seeded synthetic pile windows (consent_b200/host/synth.h) — shared input for the GPU path,
the oracle and the reference harness.  Profiles: SURVEY §8d / BASELINE.md §3.3."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._ffi import Batch, PKG_DIR, cg_synth_spec, load_library

PROFILES = {
    # total error, share substitutions, share insertions (deletions = rest)
    "PB": (0.15, 0.10, 0.60),
    "ONT": (0.10, 0.40, 0.20),
}

_lib = None


def _host():
    global _lib
    if _lib is None:
        _lib = load_library(os.path.join(PKG_DIR, "libconsent_host.so"))
        _lib.cg_synth_max_bases.restype = C.c_uint64
        _lib.cg_synth_max_bases.argtypes = [C.POINTER(cg_synth_spec)]
        _lib.cg_synth_windows.restype = C.c_uint64
        _lib.cg_synth_windows.argtypes = [C.POINTER(cg_synth_spec), C.POINTER(C.c_uint32),
                                          C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
    return _lib


def synth_windows(n_windows: int, n_seqs: int, *, seed: int = 42, profile: str = "PB",
                  truth_len: int = 500, first_window: int = 0, threads: int | None = None) -> Batch:
    err, p_sub, p_ins = PROFILES[profile]
    spec = cg_synth_spec(seed, first_window, n_windows, n_seqs, truth_len, err, p_sub, p_ins)
    lib = _host()
    cap = int(lib.cg_synth_max_bases(C.byref(spec)))
    wsb = np.zeros(n_windows + 1, np.uint32)
    off = np.zeros(n_windows * n_seqs + 1, np.uint64)
    bases = np.empty(max(cap, 1), np.uint8)
    if threads is None:
        threads = min(os.cpu_count() or 1, 32)
    n = lib.cg_synth_windows(C.byref(spec), wsb.ctypes.data_as(C.POINTER(C.c_uint32)),
                             off.ctypes.data_as(C.POINTER(C.c_uint64)),
                             C.cast(bases.ctypes.data, C.c_char_p), threads)
    return Batch(wsb, off, bases[:max(int(n), 1)].copy())
