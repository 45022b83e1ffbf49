"""Seeded synthetic pile windows (consent_b200/host/synth.h) — shared input for the GPU path,
the oracle and the reference harness.  Profiles: SURVEY §8d / BASELINE.md §3.3."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._ffi import Batch, PKG_DIR, Piles, Reads, cg_synth_pile_spec, cg_synth_read_spec, cg_synth_spec, load_library

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
        _lib.cg_synth_reads_bounds.restype = None
        _lib.cg_synth_reads_bounds.argtypes = [C.POINTER(cg_synth_read_spec)] + [C.POINTER(C.c_uint64)] * 4
        _lib.cg_synth_reads.restype = C.c_uint64
        _lib.cg_synth_reads.argtypes = [C.POINTER(cg_synth_read_spec), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                        C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_char_p,
                                        C.POINTER(C.c_uint32)]
        _lib.cg_synth_piles_build.restype = C.c_void_p
        _lib.cg_synth_piles_build.argtypes = [C.POINTER(cg_synth_pile_spec), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib.cg_synth_piles_fetch.restype = None
        _lib.cg_synth_piles_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_char_p] + [C.POINTER(C.c_uint32)] * 4
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


def synth_reads(n_reads: int, n_seqs: int, *, truth_len: int = 3000, seed: int = 42, profile: str = "PB",
                window_size: int = 500, window_overlap: int = 50, first_read: int = 0,
                thin_every: int = 0, thin_seqs: int = 1) -> tuple[Batch, Reads]:
    """Seeded reads with the windows the reference would cut from them (consent_b200/host/synth.h):
    -> (window batch, reads).  Input of the re-anchoring path."""
    err, p_sub, p_ins = PROFILES[profile]
    spec = cg_synth_read_spec(seed, first_read, n_reads, n_seqs, truth_len, window_size, window_overlap,
                              thin_every, thin_seqs, err, p_sub, p_ins)
    lib = _host()
    mw, ms, mb, mr = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.cg_synth_reads_bounds(C.byref(spec), C.byref(mw), C.byref(ms), C.byref(mb), C.byref(mr))
    wsb = np.zeros(mw.value + 1, np.uint32)
    off = np.zeros(ms.value + 1, np.uint64)
    bases = np.empty(max(mb.value, 1), np.uint8)
    rwb = np.zeros(n_reads + 1, np.uint32)
    roff = np.zeros(n_reads + 1, np.uint64)
    rbases = np.empty(max(mr.value, 1), np.uint8)
    wpos = np.zeros(mw.value + 1, np.uint32)
    u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    W = int(lib.cg_synth_reads(C.byref(spec), wsb.ctypes.data_as(u32p), off.ctypes.data_as(u64p),
                               C.cast(bases.ctypes.data, C.c_char_p), rwb.ctypes.data_as(u32p),
                               roff.ctypes.data_as(u64p), C.cast(rbases.ctypes.data, C.c_char_p),
                               wpos.ctypes.data_as(u32p)))
    S = int(wsb[W])
    batch = Batch(wsb[:W + 1].copy(), off[:S + 1].copy(), bases[:max(int(off[S]), 1)].copy())
    reads = Reads(rwb, roff, rbases[:max(int(roff[-1]), 1)].copy(), wpos[:W].copy(), window_size, window_overlap)
    return batch, reads


def synth_piles(n_reads: int, *, genome_len: int = 200_000, read_len: int = 4000, n_piles: int | None = None, seed: int = 42,
                profile: str = "PB", max_support: int = 150, min_overlap: int = 500, min_support: int = 3,
                window_size: int = 500, window_overlap: int = 50) -> Piles:
    """Seeded reads over a random genome with their true overlaps as read piles (consent_b200/host/synth.h): the input of
    the window-extraction path.  Coverage = n_reads * read_len / genome_len."""
    err, p_sub, p_ins = PROFILES[profile]
    n_piles = n_reads if n_piles is None else n_piles
    spec = cg_synth_pile_spec(seed, genome_len, n_reads, read_len, n_piles, max_support, min_overlap, err, p_sub, p_ins)
    lib = _host()
    nb, no = C.c_uint64(), C.c_uint64()
    h = lib.cg_synth_piles_build(C.byref(spec), C.byref(nb), C.byref(no))
    store_off = np.zeros(n_reads + 1, np.uint64)
    store = np.empty(max(nb.value, 1), np.uint8)
    P = min(n_piles, n_reads)
    pile_read, pile_qlen = np.zeros(max(P, 1), np.uint32), np.zeros(max(P, 1), np.uint32)
    pile_ov_begin = np.zeros(P + 1, np.uint32)
    ov = np.zeros((max(no.value, 1), 7), np.uint32)
    u32, u64 = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    lib.cg_synth_piles_fetch(h, store_off.ctypes.data_as(u64), C.cast(store.ctypes.data, C.c_char_p), pile_read.ctypes.data_as(u32),
                             pile_qlen.ctypes.data_as(u32), pile_ov_begin.ctypes.data_as(u32), ov.ctypes.data_as(u32))
    return Piles(store_off, store, pile_read[:P], pile_qlen[:P], pile_ov_begin, ov[:no.value], min_support, window_size, window_overlap)


def synth_paf(piles: Piles, *, seed: int = 1, tie_range: int = 40, blank_every: int = 0, extra_columns: bool = True):
    """A PAF text (bytes) + the read names for `piles` (a pile's overlaps in their given order): what minimap2 would have
    written for these overlaps.  resMatches (column 10) is drawn from `tie_range` values per pile so that the cut to maxSupport
    has ties to break; blank_every > 0 inserts an empty line before every blank_every-th pile (a pile separator for
    getNextReadPile, reference src/alignmentPiles.cpp:29-37).  -> (text, ReadNames)"""
    from ._ffi import ReadNames
    rng = np.random.default_rng(seed)
    names = [f"read_{i}/{(i * 7919) % 1000}" for i in range(piles.n_store)]
    out = []
    ovb = piles.pile_ov_begin
    for p in range(piles.n_piles):
        a, b = int(ovb[p]), int(ovb[p + 1])
        if blank_every and p and p % blank_every == 0:
            out.append("")
        q, qlen = names[int(piles.pile_read[p])], int(piles.pile_qlen[p])
        base = int(rng.integers(100, 3000))
        res = base + rng.integers(0, max(tie_range, 1), size=b - a)
        for i in range(a, b):
            t, st, qs, qe, ts, te, tl = (int(x) for x in piles.overlaps[i])
            r = int(res[i - a])
            line = f"{q}\t{qlen}\t{qs}\t{qe + 1}\t{'-' if st else '+'}\t{names[t]}\t{tl}\t{ts}\t{te + 1}\t{r}\t{r + 57}\t255"
            if extra_columns and (i & 1):
                line += "\ttp:A:S\tcm:i:73\ts1:i:512\tdv:f:0.1201"
            out.append(line)
    text = ("\n".join(out) + "\n").encode() if out else b""
    return text, ReadNames(names)
