"""Test-only access to the checkers: oracle/liboracle.so (our C restatement) and
oracle/_ref/libconsent_ref_*.so (the unmodified reference).  Product code never imports this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from consent_b200._ffi import (Batch, Corrected, Params, Piles, PileSet, REPO_DIR, ReadNames, Reads, Results, cg_batch,
                               cg_corrected, cg_counters, cg_params, cg_pile_set, cg_piles, cg_read_names, cg_reads, cg_results,
                               cg_window_set, results_to_c, window_set_to_py)

ORACLE_DIR = os.path.join(REPO_DIR, "oracle")


def _cpu_has(flag: str) -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return flag in line.split()
    except OSError:
        pass
    return False


def ref_library_path() -> str | None:
    for name, flag in (("avx2", "avx2"), ("sse41", "sse4_1")):
        p = os.path.join(ORACLE_DIR, "_ref", f"libconsent_ref_{name}.so")
        if os.path.exists(p) and _cpu_has(flag):
            return p
    return None


class _Checker:
    prefix = ""

    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)
        f = getattr(self.lib, self.prefix + "_correct_windows")
        f.restype = C.c_int
        f.argtypes = [C.POINTER(cg_batch), C.POINTER(cg_params), C.c_int, C.c_int,
                      C.POINTER(cg_results), C.POINTER(C.c_double)]
        self._run = f
        self._free = getattr(self.lib, self.prefix + "_free_results")
        self._free.argtypes = [C.POINTER(cg_results)]
        self._dump = getattr(self.lib, self.prefix + "_dump_window")
        self._dump.restype = C.c_void_p
        self._dump.argtypes = [C.POINTER(cg_batch), C.c_uint32, C.POINTER(cg_params)]
        self._msa = getattr(self.lib, self.prefix + "_spoa_msa")
        self._msa.restype = C.c_void_p
        self._msa.argtypes = [C.POINTER(C.c_char_p), C.c_uint32]
        self._free_text = getattr(self.lib, self.prefix + "_free_text")
        self._free_text.argtypes = [C.c_void_p]

    def extract_windows(self, piles: Piles, mer_size: int = 9):
        """Phase A of processRead: window positions + piles of every read pile -> (Batch, Reads, win_end)"""
        f = getattr(self.lib, self.prefix + "_extract_windows")
        f.restype = C.c_int
        f.argtypes = [C.POINTER(cg_piles), C.c_uint, C.POINTER(cg_window_set)]
        free = getattr(self.lib, self.prefix + "_free_window_set")
        free.argtypes = [C.POINTER(cg_window_set)]
        cp, ws = piles.c(), cg_window_set()
        rc = f(C.byref(cp), mer_size, C.byref(ws))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}_extract_windows -> {rc}")
        out = window_set_to_py(ws)
        free(C.byref(ws))
        return out

    def ingest_paf(self, text: bytes, names: ReadNames, max_support: int = 150) -> PileSet:
        """Every getNextReadPile of a PAF text -> PileSet"""
        f = getattr(self.lib, self.prefix + "_ingest_paf")
        f.restype = C.c_int
        f.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(cg_read_names), C.c_uint32, C.POINTER(cg_pile_set)]
        free = getattr(self.lib, self.prefix + "_free_pile_set")
        free.argtypes = [C.POINTER(cg_pile_set)]
        cn, out = names.c(), cg_pile_set()
        rc = f(text, len(text), C.byref(cn), int(max_support), C.byref(out))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}_ingest_paf -> {rc}")
        got = PileSet(out)
        free(C.byref(out))
        return got

    def sort_desc(self, keys) -> np.ndarray:
        """std::sort(rbegin, rend) on records compared by `keys` -> the resulting order (indices)"""
        f = getattr(self.lib, self.prefix + "_sort_desc")
        f.restype = None
        f.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]
        k = np.ascontiguousarray(keys, np.uint32)
        o = np.zeros(max(len(k), 1), np.uint32)
        f(k.ctypes.data_as(C.POINTER(C.c_uint32)) if len(k) else o.ctypes.data_as(C.POINTER(C.c_uint32)), len(k),
          o.ctypes.data_as(C.POINTER(C.c_uint32)))
        return o[:len(k)]

    def finish_reads(self, cor: Corrected, trim_mer: int = 1) -> Corrected:
        """trimRead + dropRead on the strings alignConsensus returned"""
        f = getattr(self.lib, self.prefix + "_finish_reads")
        f.restype = C.c_int
        f.argtypes = [C.POINTER(cg_corrected), C.c_uint32, C.POINTER(cg_corrected)]
        free = getattr(self.lib, "ref_free_corrected" if self.prefix == "ref" else "oracle_free_finished")
        free.argtypes = [C.POINTER(cg_corrected)]
        off = np.ascontiguousarray(cor.read_off, np.uint64)
        b = cor.bases if len(cor.bases) else np.zeros(1, np.uint8)
        cin = cg_corrected(cor.n_reads, off.ctypes.data_as(C.POINTER(C.c_uint64)), C.cast(b.ctypes.data, C.POINTER(C.c_char)), None)
        out = cg_corrected()
        rc = f(C.byref(cin), int(trim_mer), C.byref(out))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}_finish_reads -> {rc}")
        got = Corrected(out)
        free(C.byref(out))
        return got

    def reanchor_reads(self, batch: Batch, res: Results, reads: Reads, params: Params = Params(), threads: int = 1):
        """alignConsensus for every read -> (Corrected, seconds of the compute loop)"""
        f = getattr(self.lib, self.prefix + "_reanchor_reads")
        f.restype = C.c_int
        f.argtypes = [C.POINTER(cg_batch), C.POINTER(cg_results), C.POINTER(cg_reads), C.POINTER(cg_params), C.c_int,
                      C.POINTER(cg_corrected), C.POINTER(C.c_double)]
        free = getattr(self.lib, self.prefix + "_free_corrected")
        free.argtypes = [C.POINTER(cg_corrected)]
        cb, cr, rd, cp, out, sec = batch.c(), results_to_c(res), reads.c(), params.c(), cg_corrected(), C.c_double(0)
        rc = f(C.byref(cb), C.byref(cr), C.byref(rd), C.byref(cp), threads, C.byref(out), C.byref(sec))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}_reanchor_reads -> {rc}")
        got = Corrected(out)
        free(C.byref(out))
        return got, sec.value

    def correct_windows(self, batch: Batch, params: Params = Params(), threads: int = 1,
                        with_status: bool = True):
        """-> (Results, seconds of the compute loop)"""
        cb, cp, r, sec = batch.c(), params.c(), cg_results(), C.c_double(0)
        rc = self._run(C.byref(cb), C.byref(cp), threads, int(with_status), C.byref(r), C.byref(sec))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}_correct_windows -> {rc}")
        out = Results(r)
        self._free(C.byref(r))
        return out, sec.value

    def dump_window(self, batch: Batch, w: int, params: Params = Params()) -> str:
        cb, cp = batch.c(), params.c()
        p = self._dump(C.byref(cb), w, C.byref(cp))
        s = C.string_at(p).decode()
        self._free_text(p)
        return s

    def spoa_msa(self, seqs) -> list[str]:
        arr = (C.c_char_p * len(seqs))(*[s.encode() for s in seqs])
        p = self._msa(arr, len(seqs))
        s = C.string_at(p).decode()
        self._free_text(p)
        return s.split("\n")[:-1] if s else []


class Reference(_Checker):
    """The unmodified reference (oracle/_ref)."""
    prefix = "ref"

    def __init__(self):
        p = ref_library_path()
        if p is None:
            raise FileNotFoundError("oracle/_ref/libconsent_ref_*.so not built (make -C oracle ref)")
        super().__init__(p)
        self.lib.ref_hardware_threads.restype = C.c_int

    def hardware_threads(self) -> int:
        return int(self.lib.ref_hardware_threads())


class Oracle(_Checker):
    """Our plain-C restatement (oracle/consent_oracle.c)."""
    prefix = "oracle"

    def __init__(self):
        p = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(p):
            raise FileNotFoundError("oracle/liboracle.so not built (make -C oracle oracle)")
        super().__init__(p)
        self.lib.oracle_get_counters.argtypes = [C.POINTER(cg_counters)]

    def counters(self) -> dict:
        c = cg_counters()
        self.lib.oracle_get_counters(C.byref(c))
        return c.as_dict()
