"""Host-side mirror of the reference's per-window operator, batched.

`Corrector.correct_windows(batch)` is `computeConsensusReadCorrection` (reference
src/correctionMSA.cpp:29-49; `computeConsensusAssemblyPolishing` :51-70 has the same body) applied to
every window of `batch`, through the C ABI of include/consent_b200.h.  Same argument meaning as the
reference (`Params` = merSize / solidThresh / commonKMers / minAnchors), same results: per window the
mixed-case consensus, the solid k-mer counts (`merCounts` entries >= solidThresh) and whether the raw
template was returned.  Errors are exceptions carrying the library's status code.

There is no CPU path: the CUDA library must be built (`__graft_entry__.build()`) and a GPU present.
"""
from __future__ import annotations

import ctypes as C
import os

from ._ffi import (Batch, CG_N_STAGES, KERNEL_NAMES, cg_kernel_stats, Corrected, PKG_DIR, Params, Piles, PileSet, ReadNames, Reads, Results, STAGE_NAMES,
                   STATUS_NAMES, cg_batch, cg_corrected, cg_counters, cg_params, cg_pile_set, cg_piles, cg_read_names, cg_reads,
                   cg_results, cg_window_set, load_library, results_to_c, window_set_to_py)

LIB_PATH = os.path.join(PKG_DIR, "libconsent_b200.so")

EXPORTS = ("cg_abi_version", "cg_device_count", "cg_create", "cg_destroy", "cg_last_error", "cg_set_option", "cg_pack_bases_2bit",
           "cg_correct_windows", "cg_free_results", "cg_upload", "cg_run", "cg_download", "cg_stage_ms",
           "cg_get_counters", "cg_get_kernel_stats", "cg_debug_dump_window", "cg_run_ms", "cg_chunk_count", "cg_reanchor_reads", "cg_free_corrected", "cg_reanchor_stats",
           "cg_upload_piles", "cg_set_read_store", "cg_download_windows", "cg_free_window_set", "cg_extract_stats",
           "cg_ingest_paf", "cg_free_pile_set", "cg_ingest_stats", "cg_finish_reads", "cg_finish_stats", "cg_finish_resident")


class ConsentError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {what}")
        self.code = code


def bind(lib: C.CDLL) -> C.CDLL:
    H = C.c_void_p
    lib.cg_abi_version.restype = C.c_int
    lib.cg_device_count.restype = C.c_int
    lib.cg_create.restype = C.c_int
    lib.cg_create.argtypes = [C.c_int, C.POINTER(cg_params), C.POINTER(H)]
    lib.cg_destroy.argtypes = [H]
    lib.cg_last_error.restype = C.c_char_p
    lib.cg_last_error.argtypes = [H]
    lib.cg_pack_bases_2bit.restype = None
    lib.cg_pack_bases_2bit.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint8), C.c_int]
    lib.cg_set_option.restype = C.c_int
    lib.cg_set_option.argtypes = [H, C.c_char_p, C.c_longlong]
    for name in ("cg_correct_windows",):
        f = getattr(lib, name)
        f.restype = C.c_int
        f.argtypes = [H, C.POINTER(cg_batch), C.POINTER(cg_results)]
    lib.cg_upload.restype = C.c_int
    lib.cg_upload.argtypes = [H, C.POINTER(cg_batch)]
    lib.cg_run.restype = C.c_int
    lib.cg_run.argtypes = [H]
    lib.cg_download.restype = C.c_int
    lib.cg_download.argtypes = [H, C.POINTER(cg_results)]
    lib.cg_free_results.argtypes = [C.POINTER(cg_results)]
    lib.cg_stage_ms.restype = C.c_int
    lib.cg_stage_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    lib.cg_run_ms.restype = C.c_int
    lib.cg_run_ms.argtypes = [H, C.POINTER(C.c_float)]
    lib.cg_chunk_count.restype = C.c_int
    lib.cg_chunk_count.argtypes = [H]
    lib.cg_get_counters.restype = C.c_int
    lib.cg_get_counters.argtypes = [H, C.POINTER(cg_counters)]
    lib.cg_debug_dump_window.restype = C.c_int
    lib.cg_debug_dump_window.argtypes = [H, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.cg_get_kernel_stats.restype = C.c_int
    lib.cg_get_kernel_stats.argtypes = [H, C.POINTER(cg_kernel_stats)]
    lib.cg_reanchor_reads.restype = C.c_int
    lib.cg_reanchor_reads.argtypes = [H, C.POINTER(cg_batch), C.POINTER(cg_results), C.POINTER(cg_reads), C.POINTER(cg_corrected)]
    lib.cg_free_corrected.argtypes = [C.POINTER(cg_corrected)]
    lib.cg_reanchor_stats.restype = C.c_int
    lib.cg_reanchor_stats.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    lib.cg_upload_piles.restype = C.c_int
    lib.cg_upload_piles.argtypes = [H, C.POINTER(cg_piles)]
    lib.cg_set_read_store.restype = C.c_int
    lib.cg_set_read_store.argtypes = [H, C.c_uint32, C.POINTER(C.c_uint64), C.c_char_p]
    lib.cg_download_windows.restype = C.c_int
    lib.cg_download_windows.argtypes = [H, C.c_int, C.POINTER(cg_window_set)]
    lib.cg_free_window_set.argtypes = [C.POINTER(cg_window_set)]
    lib.cg_extract_stats.restype = C.c_int
    lib.cg_extract_stats.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    lib.cg_ingest_paf.restype = C.c_int
    lib.cg_ingest_paf.argtypes = [H, C.c_char_p, C.c_uint64, C.POINTER(cg_read_names), C.c_uint32, C.POINTER(cg_pile_set)]
    lib.cg_free_pile_set.argtypes = [C.POINTER(cg_pile_set)]
    lib.cg_ingest_stats.restype = C.c_int
    lib.cg_ingest_stats.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    lib.cg_finish_resident.restype = C.c_int
    lib.cg_finish_resident.argtypes = [H, C.c_uint32, C.POINTER(cg_corrected)]
    lib.cg_finish_stats.restype = C.c_int
    lib.cg_finish_stats.argtypes = [H, C.POINTER(C.c_float)]
    lib.cg_finish_reads.restype = C.c_int
    lib.cg_finish_reads.argtypes = [H, C.POINTER(cg_batch), C.POINTER(cg_results), C.POINTER(cg_reads), C.c_uint32, C.POINTER(cg_corrected)]
    return lib


class Corrector:
    """One context = one GPU, one stream, its own workspaces (cg_handle)."""

    def __init__(self, params: Params = Params(), device: int = 0, lib_path: str | None = None):
        self.lib = bind(load_library(lib_path or LIB_PATH))
        self.params = params
        self._h = C.c_void_p()
        cp = params.c()
        rc = self.lib.cg_create(device, C.byref(cp), C.byref(self._h))
        if rc != 0:
            raise ConsentError(rc, (self.lib.cg_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.cg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise ConsentError(rc, (self.lib.cg_last_error(self._h) or b"").decode())

    def _free_results(self, r):
        self.lib.cg_free_results(C.byref(r))

    def set_option(self, key: str, value: int):
        self._check(self.lib.cg_set_option(self._h, key.encode(), int(value)))

    # -- the operator -------------------------------------------------------------------------
    def correct_windows(self, batch: Batch) -> Results:
        """Host buffers in, host buffers out (H2D, every kernel, D2H)."""
        cb, r = batch.c(), cg_results()
        self._check(self.lib.cg_correct_windows(self._h, C.byref(cb), C.byref(r)))
        return Results(r, free=self._free_results)

    # -- re-anchoring: alignConsensus (reference src/correctionAlignment.cpp:47-139) for every read --------------
    def reanchor_reads(self, batch: Batch, results: Results, reads: Reads) -> Corrected:
        """The window consensuses of every read aligned back onto it, overlaps arbitrated, aligned stretches
        replaced by the upper-cased consensus: what `alignConsensus` returns per read (before the trim / drop
        post-filters of processRead).  `results` = what correct_windows(batch) returned (when it is that very
        object the device copies are reused), read r owns windows reads.read_win_begin[r] .. [r+1]."""
        cb, rd, out = batch.c(), reads.c(), cg_corrected()
        live = getattr(results, "_r", None)
        cr = live if live is not None else results_to_c(results)
        self._check(self.lib.cg_reanchor_reads(self._h, C.byref(cb), C.byref(cr), C.byref(rd), C.byref(out)))
        got = Corrected(out)
        self.lib.cg_free_corrected(C.byref(out))
        return got

    def finish_reads(self, batch: Batch, results: Results, reads: Reads, trim_mer: int = 1) -> Corrected:
        """reanchor_reads followed on the device by the tail of processRead (reference src/CONSENT-correction.cpp:49-59):
        trimRead(read, trim_mer) and dropRead — the sequence line of every FASTA record (empty = no record).  trim_mer = 0
        is the proof-file mode (no trimming, nothing dropped)."""
        cb, rd, out = batch.c(), reads.c(), cg_corrected()
        live = getattr(results, "_r", None)
        cr = live if live is not None else results_to_c(results)
        self._check(self.lib.cg_finish_reads(self._h, C.byref(cb), C.byref(cr), C.byref(rd), int(trim_mer), C.byref(out)))
        got = Corrected(out)
        self.lib.cg_free_corrected(C.byref(out))
        return got

    def finish_resident(self, trim_mer: int = 1) -> Corrected:
        """finish_reads for the batch upload_piles() + run() left on the device, with nothing but the corrected reads crossing to the
        host (no download(), no download_windows())."""
        out = cg_corrected()
        self._check(self.lib.cg_finish_resident(self._h, int(trim_mer), C.byref(out)))
        got = Corrected(out)
        self.lib.cg_free_corrected(C.byref(out))
        return got

    def finish_stats(self) -> dict:
        ms = C.c_float(0)
        self._check(self.lib.cg_finish_stats(self._h, C.byref(ms)))
        return {"kernel_ms": float(ms.value)}

    # -- PAF ingest: every getNextReadPile (reference src/alignmentPiles.cpp:22-58) of a PAF text on the device -------------
    def ingest_paf(self, text: bytes, names: ReadNames, max_support: int = 150) -> PileSet:
        """PAF text in, read piles out: lines parsed (src/Overlap.h:26-60), consecutive lines of one query grouped, every pile
        ordered as std::sort(rbegin, rend) by resMatches leaves it and cut to max_support."""
        cn, out = names.c(), cg_pile_set()
        self._check(self.lib.cg_ingest_paf(self._h, text, len(text), C.byref(cn), int(max_support), C.byref(out)))
        got = PileSet(out)
        self.lib.cg_free_pile_set(C.byref(out))
        return got

    def ingest_stats(self) -> dict:
        ms, pms, nb = C.c_float(0), C.c_float(0), C.c_uint64(0)
        self._check(self.lib.cg_ingest_stats(self._h, C.byref(ms), C.byref(pms), C.byref(nb)))
        return {"kernel_ms": float(ms.value), "parse_ms": float(pms.value), "paf_bytes": int(nb.value)}

    def reanchor_stats(self) -> dict:
        ms, cells = C.c_float(0), C.c_uint64(0)
        self._check(self.lib.cg_reanchor_stats(self._h, C.byref(ms), C.byref(cells)))
        return {"kernel_ms": float(ms.value), "dp_cells": int(cells.value)}

    # -- window extraction: phase A of processRead (reference src/alignmentWindows.cpp:27-149) on the device ----------
    def upload_piles(self, piles: Piles):
        """Read store + read piles (overlaps per query read) in, the window batch cut on the device and left resident:
        run() / download() follow as after upload()."""
        cp = piles.c()
        self._check(self.lib.cg_upload_piles(self._h, C.byref(cp)))

    def set_read_store(self, store_off, store_bases):
        """Ship the read store once (cg_set_read_store); later upload_piles(..., resident_store=True) calls cut from it."""
        import numpy as np
        off = np.ascontiguousarray(store_off, dtype=np.uint64)
        bases = np.ascontiguousarray(store_bases, dtype=np.uint8)
        self._check(self.lib.cg_set_read_store(self._h, len(off) - 1, off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                               C.cast(bases.ctypes.data, C.c_char_p)))

    def upload_piles_resident(self, piles: Piles):
        """upload_piles against the store cg_set_read_store left on the device (the piles' own store is not sent)."""
        cp = piles.c()
        cp.store_off = None
        cp.store_bases = None
        self._check(self.lib.cg_upload_piles(self._h, C.byref(cp)))

    def download_windows(self, with_bases: bool = True):
        """The windows cg_upload_piles extracted -> (Batch, Reads, win_end): the piles, and per pile its read, its windows
        and their positions (pilesPos)."""
        ws = cg_window_set()
        self._check(self.lib.cg_download_windows(self._h, int(with_bases), C.byref(ws)))
        out = window_set_to_py(ws, with_bases)
        self.lib.cg_free_window_set(C.byref(ws))
        return out

    def extract_stats(self) -> dict:
        ms, cms, nb = C.c_float(0), C.c_float(0), C.c_uint64(0)
        self._check(self.lib.cg_extract_stats(self._h, C.byref(ms), C.byref(cms), C.byref(nb)))
        return {"kernel_ms": float(ms.value), "copy_ms": float(cms.value), "pile_bytes": int(nb.value)}

    # -- staged (bench: keep the batch resident in HBM, time the kernels alone) ---------------
    def upload(self, batch: Batch):
        cb = batch.c()
        self._check(self.lib.cg_upload(self._h, C.byref(cb)))

    def run(self):
        self._check(self.lib.cg_run(self._h))

    def download(self) -> Results:
        r = cg_results()
        self._check(self.lib.cg_download(self._h, C.byref(r)))
        return Results(r, free=self._free_results)

    # -- instrumentation ----------------------------------------------------------------------
    def stage_ms(self) -> dict:
        ms = (C.c_float * CG_N_STAGES)()
        n = (C.c_uint32 * CG_N_STAGES)()
        self._check(self.lib.cg_stage_ms(self._h, ms, n))
        return {name: {"ms": float(ms[i]), "launches": int(n[i])} for i, name in enumerate(STAGE_NAMES)}

    def pack_2bit(self, batch: Batch, threads: int = 8, pinned: bool = False) -> Batch:
        """The batch with its bases packed 2 bits per base (cg_pack_bases_2bit) — input of a handle with set_option("input_2bit", 1)."""
        import numpy as np
        n = int(batch.n_bases)
        if pinned:
            import torch
            keep = torch.empty((n + 3) // 4 + 8, dtype=torch.uint8).pin_memory()
            out = keep.numpy()
        else:
            keep, out = None, np.empty((n + 3) // 4 + 8, np.uint8)
        self.lib.cg_pack_bases_2bit(C.cast(batch.bases.ctypes.data, C.c_char_p), n, out.ctypes.data_as(C.POINTER(C.c_uint8)), threads)
        nb = Batch.__new__(Batch)
        nb.win_seq_begin, nb.seq_off, nb.bases = batch.win_seq_begin, batch.seq_off, out
        nb._keep = (keep, getattr(batch, "_keep", None))
        return nb

    def dump_window(self, w: int) -> str:
        """Per-stage dump of window w of the last (single-chunk) run, in the oracle's format (lines S M T A R G g c C)."""
        p = C.c_void_p()
        self._check(self.lib.cg_debug_dump_window(self._h, w, C.byref(p)))
        text = C.string_at(p).decode()
        C.CDLL(None).free(p)
        return text

    def kernel_stats(self) -> dict:
        """Per kernel of the last run(): summed launch durations (own CUDA events on the launching stream), launches, and for the
        POA tiers the score-matrix / predecessor-row cells they computed."""
        k = cg_kernel_stats()
        self._check(self.lib.cg_get_kernel_stats(self._h, C.byref(k)))
        out = {name: {"ms": float(k.ms[i]), "launches": int(k.launches[i])} for i, name in enumerate(KERNEL_NAMES)}
        for t, name in enumerate(KERNEL_NAMES[4:8]):
            out[name]["dp_cells"] = int(k.poa_cells[t]); out[name]["dp_pred_cells"] = int(k.poa_pred_cells[t])
        return out

    def run_ms(self) -> float:
        """CUDA-event time of the whole last run() on the library's stream."""
        ms = C.c_float(0)
        self._check(self.lib.cg_run_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def chunk_count(self) -> int:
        """Chunks the uploaded batch is processed in (one launch of every kernel per chunk)."""
        return int(self.lib.cg_chunk_count(self._h))

    def counters(self) -> dict:
        c = cg_counters()
        self._check(self.lib.cg_get_counters(self._h, C.byref(c)))
        return c.as_dict()
