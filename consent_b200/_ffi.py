"""ctypes mirror of include/consent_b200.h and consent_b200/host/synth.h.

Only plain C structs live here; nothing in this module computes anything.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)

CG_OK = 0
CG_WINDOW_CONSENSUS = 0
CG_WINDOW_TEMPLATE = 1
CG_N_STAGES = 7
STAGE_NAMES = ("pack", "index", "chain", "split", "poa", "stitch", "polish")
STATUS_NAMES = {
    0: "CG_OK", -1: "CG_ERR_INVALID_ARG", -2: "CG_ERR_NO_DEVICE", -3: "CG_ERR_CUDA",
    -4: "CG_ERR_OUT_OF_MEMORY", -5: "CG_ERR_BAD_BASE", -6: "CG_ERR_CAPACITY", -7: "CG_ERR_STATE",
}


class cg_params(C.Structure):
    _fields_ = [("mer_size", C.c_uint32), ("solid_thresh", C.c_uint32),
                ("common_kmers", C.c_uint32), ("min_anchors", C.c_uint32)]


class cg_batch(C.Structure):
    _fields_ = [("n_windows", C.c_uint32),
                ("win_seq_begin", C.POINTER(C.c_uint32)),
                ("seq_off", C.POINTER(C.c_uint64)),
                ("bases", C.c_char_p)]


class cg_results(C.Structure):
    _fields_ = [("n_windows", C.c_uint32),
                ("cons_off", C.POINTER(C.c_uint64)),
                ("cons", C.POINTER(C.c_char)),
                ("status", C.POINTER(C.c_uint8)),
                ("solid_off", C.POINTER(C.c_uint64)),
                ("solid_kmer", C.POINTER(C.c_uint32)),
                ("solid_count", C.POINTER(C.c_uint32)),
                ("owner_", C.c_void_p)]


class cg_reads(C.Structure):
    _fields_ = [("n_reads", C.c_uint32),
                ("read_win_begin", C.POINTER(C.c_uint32)),
                ("read_off", C.POINTER(C.c_uint64)),
                ("read_bases", C.c_char_p),
                ("win_pos", C.POINTER(C.c_uint32)),
                ("window_size", C.c_uint32),
                ("window_overlap", C.c_uint32)]


class cg_corrected(C.Structure):
    _fields_ = [("n_reads", C.c_uint32),
                ("read_off", C.POINTER(C.c_uint64)),
                ("bases", C.POINTER(C.c_char)),
                ("owner_", C.c_void_p)]


class cg_overlap(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("t_read", "strand", "q_start", "q_end", "t_start", "t_end", "t_length")]


class cg_piles(C.Structure):
    _fields_ = [("n_store", C.c_uint32), ("store_off", C.POINTER(C.c_uint64)), ("store_bases", C.c_char_p),
                ("n_piles", C.c_uint32), ("pile_read", C.POINTER(C.c_uint32)), ("pile_qlen", C.POINTER(C.c_uint32)),
                ("pile_ov_begin", C.POINTER(C.c_uint32)), ("overlaps", C.POINTER(cg_overlap)),
                ("min_support", C.c_uint32), ("window_size", C.c_uint32), ("window_overlap", C.c_uint32)]


class cg_window_set(C.Structure):
    _fields_ = [("batch", cg_batch), ("reads", cg_reads), ("win_end", C.POINTER(C.c_uint32)), ("owner_", C.c_void_p)]


class cg_read_names(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("name_off", C.POINTER(C.c_uint64)), ("names", C.c_char_p)]


class cg_pile_set(C.Structure):
    _fields_ = [("n_piles", C.c_uint32), ("pile_read", C.POINTER(C.c_uint32)), ("pile_qlen", C.POINTER(C.c_uint32)),
                ("pile_ov_begin", C.POINTER(C.c_uint32)), ("overlaps", C.POINTER(cg_overlap)),
                ("res_matches", C.POINTER(C.c_uint32)), ("n_lines", C.c_uint64), ("owner_", C.c_void_p)]


class cg_synth_pile_spec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("genome_len", C.c_uint32), ("n_reads", C.c_uint32), ("read_len", C.c_uint32),
                ("n_piles", C.c_uint32), ("max_support", C.c_uint32), ("min_overlap", C.c_uint32),
                ("err", C.c_double), ("p_sub", C.c_double), ("p_ins", C.c_double)]


class cg_synth_read_spec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("first_read", C.c_uint32), ("n_reads", C.c_uint32),
                ("n_seqs", C.c_uint32), ("truth_len", C.c_uint32),
                ("window_size", C.c_uint32), ("window_overlap", C.c_uint32),
                ("thin_every", C.c_uint32), ("thin_seqs", C.c_uint32),
                ("err", C.c_double), ("p_sub", C.c_double), ("p_ins", C.c_double)]


class cg_counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "windows", "sequences", "bases", "anchors", "regions", "poa_graphs", "alignments",
        "dp_cells", "dp_pred_cells", "solid_kmers", "consensus_bytes", "fallback_windows", "error_windows")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


CG_N_KERNELS = 11
KERNEL_NAMES = ("k_pack", "k_index", "k_chain", "k_split", "k_poa2<C1>", "k_poa2<G>", "k_poa2<W1>", "k_poa2<W2>", "k_poa", "k_polish", "k_out")


class cg_kernel_stats(C.Structure):
    _fields_ = [("ms", C.c_float * CG_N_KERNELS), ("launches", C.c_uint32 * CG_N_KERNELS),
                ("poa_cells", C.c_uint64 * 4), ("poa_pred_cells", C.c_uint64 * 4)]


class cg_synth_spec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("first_window", C.c_uint32), ("n_windows", C.c_uint32),
                ("n_seqs", C.c_uint32), ("truth_len", C.c_uint32),
                ("err", C.c_double), ("p_sub", C.c_double), ("p_ins", C.c_double)]


@dataclass
class Params:
    """The four parameters the path reads (reference src/correctionMSA.cpp:31-32,43-45);
    defaults = CONSENT-correct wrapper defaults (reference CONSENT-correct:42-50)."""
    mer_size: int = 9
    solid_thresh: int = 4
    common_kmers: int = 8
    min_anchors: int = 2

    def c(self) -> cg_params:
        return cg_params(self.mer_size, self.solid_thresh, self.common_kmers, self.min_anchors)


class Batch:
    """W windows of piles in the flat layout of cg_batch (host memory, numpy-owned)."""

    def __init__(self, win_seq_begin: np.ndarray, seq_off: np.ndarray, bases: np.ndarray):
        self.win_seq_begin = np.ascontiguousarray(win_seq_begin, dtype=np.uint32)
        self.seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        assert self.win_seq_begin.ndim == 1 and len(self.win_seq_begin) >= 1
        assert len(self.seq_off) == int(self.win_seq_begin[-1]) + 1
        assert int(self.seq_off[-1]) <= len(self.bases)

    @property
    def n_windows(self) -> int:
        return len(self.win_seq_begin) - 1

    @property
    def n_seqs(self) -> int:
        return int(self.win_seq_begin[-1])

    @property
    def n_bases(self) -> int:
        return int(self.seq_off[-1])

    @classmethod
    def from_piles(cls, piles) -> "Batch":
        """piles: iterable of windows, each a list of str/bytes sequences (piles[w][0] = template)."""
        wsb = [0]
        off = [0]
        chunks = []
        for pile in piles:
            for s in pile:
                b = s.encode() if isinstance(s, str) else bytes(s)
                chunks.append(b)
                off.append(off[-1] + len(b))
            wsb.append(wsb[-1] + len(pile))
        bases = np.frombuffer(b"".join(chunks), dtype=np.uint8) if chunks else np.zeros(0, np.uint8)
        if len(bases) == 0:
            bases = np.zeros(1, np.uint8)
        return cls(np.array(wsb, np.uint32), np.array(off, np.uint64), bases.copy())

    def pile(self, w: int):
        a, b = int(self.win_seq_begin[w]), int(self.win_seq_begin[w + 1])
        raw = self.bases
        return [raw[int(self.seq_off[s]):int(self.seq_off[s + 1])].tobytes().decode() for s in range(a, b)]

    def slice(self, w0: int, w1: int) -> "Batch":
        a, b = int(self.win_seq_begin[w0]), int(self.win_seq_begin[w1])
        o0, o1 = int(self.seq_off[a]), int(self.seq_off[b])
        return Batch(self.win_seq_begin[w0:w1 + 1] - np.uint32(a), self.seq_off[a:b + 1] - np.uint64(o0),
                     self.bases[o0:max(o1, o0 + 1)].copy())

    def c(self) -> cg_batch:
        return cg_batch(self.n_windows,
                        self.win_seq_begin.ctypes.data_as(C.POINTER(C.c_uint32)),
                        self.seq_off.ctypes.data_as(C.POINTER(C.c_uint64)),
                        C.cast(self.bases.ctypes.data, C.c_char_p))


class Reads:
    """R reads and the windows cut from them, in the flat layout of cg_reads (host memory, numpy-owned).
    Read r owns windows [read_win_begin[r], read_win_begin[r+1]) of the window batch it travels with."""

    def __init__(self, read_win_begin, read_off, read_bases, win_pos, window_size: int = 500, window_overlap: int = 50):
        self.read_win_begin = np.ascontiguousarray(read_win_begin, dtype=np.uint32)
        self.read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        self.read_bases = np.ascontiguousarray(read_bases, dtype=np.uint8)
        self.win_pos = np.ascontiguousarray(win_pos, dtype=np.uint32)
        if len(self.read_bases) == 0:
            self.read_bases = np.zeros(1, np.uint8)
        if len(self.win_pos) == 0:
            self.win_pos = np.zeros(1, np.uint32)
        self.window_size, self.window_overlap = int(window_size), int(window_overlap)
        assert len(self.read_off) == len(self.read_win_begin)

    @property
    def n_reads(self) -> int:
        return len(self.read_win_begin) - 1

    @property
    def n_windows(self) -> int:
        return int(self.read_win_begin[-1])

    @classmethod
    def from_lists(cls, reads, windows_per_read, win_pos, window_size=500, window_overlap=50) -> "Reads":
        rwb, off = [0], [0]
        for s, n in zip(reads, windows_per_read):
            rwb.append(rwb[-1] + n)
            off.append(off[-1] + len(s))
        raw = "".join(reads).encode()
        return cls(np.array(rwb, np.uint32), np.array(off, np.uint64), np.frombuffer(raw, np.uint8).copy(),
                   np.array(win_pos, np.uint32), window_size, window_overlap)

    def read(self, r: int) -> str:
        return self.read_bases[int(self.read_off[r]):int(self.read_off[r + 1])].tobytes().decode()

    def c(self) -> cg_reads:
        return cg_reads(self.n_reads, self.read_win_begin.ctypes.data_as(C.POINTER(C.c_uint32)),
                        self.read_off.ctypes.data_as(C.POINTER(C.c_uint64)),
                        C.cast(self.read_bases.ctypes.data, C.c_char_p),
                        self.win_pos.ctypes.data_as(C.POINTER(C.c_uint32)),
                        self.window_size, self.window_overlap)


class Piles:
    """Read store + read piles in the flat layout of cg_piles (host memory, numpy-owned).  overlaps: (n, 7) uint32 rows
    (t_read, strand, q_start, q_end, t_start, t_end, t_length), ends inclusive as src/Overlap.h keeps them."""

    def __init__(self, store_off, store_bases, pile_read, pile_qlen, pile_ov_begin, overlaps,
                 min_support: int = 3, window_size: int = 500, window_overlap: int = 50):
        self.store_off = np.ascontiguousarray(store_off, dtype=np.uint64)
        self.store_bases = np.ascontiguousarray(store_bases, dtype=np.uint8)
        self.pile_read = np.ascontiguousarray(pile_read, dtype=np.uint32)
        self.pile_qlen = np.ascontiguousarray(pile_qlen, dtype=np.uint32)
        self.pile_ov_begin = np.ascontiguousarray(pile_ov_begin, dtype=np.uint32)
        self.overlaps = np.ascontiguousarray(overlaps, dtype=np.uint32).reshape(-1, 7)
        for n in ("store_bases", "pile_read", "pile_qlen"):
            if len(getattr(self, n)) == 0:
                setattr(self, n, np.zeros(1, getattr(self, n).dtype))
        if len(self.overlaps) == 0:
            self.overlaps = np.zeros((1, 7), np.uint32)
        self.min_support, self.window_size, self.window_overlap = int(min_support), int(window_size), int(window_overlap)

    @property
    def n_piles(self) -> int:
        return len(self.pile_ov_begin) - 1

    @property
    def n_store(self) -> int:
        return len(self.store_off) - 1

    def c(self) -> cg_piles:
        u32, u64 = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        return cg_piles(self.n_store, self.store_off.ctypes.data_as(u64), C.cast(self.store_bases.ctypes.data, C.c_char_p),
                        self.n_piles, self.pile_read.ctypes.data_as(u32), self.pile_qlen.ctypes.data_as(u32),
                        self.pile_ov_begin.ctypes.data_as(u32), C.cast(self.overlaps.ctypes.data, C.POINTER(cg_overlap)),
                        self.min_support, self.window_size, self.window_overlap)


class ReadNames:
    """The read names of the store, index = store index (cg_read_names): FASTA headers up to the first blank."""

    def __init__(self, names):
        enc = [n.encode() if isinstance(n, str) else bytes(n) for n in names]
        self.name_off = np.zeros(len(enc) + 1, np.uint64)
        if enc:
            self.name_off[1:] = np.cumsum([len(e) for e in enc])
        self.names = np.frombuffer(b"".join(enc) or b"\0", np.uint8).copy()

    @property
    def n_reads(self) -> int:
        return len(self.name_off) - 1

    def c(self) -> cg_read_names:
        return cg_read_names(self.n_reads, self.name_off.ctypes.data_as(C.POINTER(C.c_uint64)), C.cast(self.names.ctypes.data, C.c_char_p))


class PileSet:
    """Host copy of cg_pile_set: the piles getNextReadPile would return for a PAF text, one after the other."""

    def __init__(self, c: cg_pile_set):
        P = int(c.n_piles)
        self.n_lines = int(c.n_lines)
        self.pile_ov_begin = np.ctypeslib.as_array(c.pile_ov_begin, shape=(P + 1,)).copy()
        n = int(self.pile_ov_begin[P])
        self.pile_read = np.ctypeslib.as_array(c.pile_read, shape=(max(P, 1),))[:P].copy()
        self.pile_qlen = np.ctypeslib.as_array(c.pile_qlen, shape=(max(P, 1),))[:P].copy()
        self.overlaps = (np.ctypeslib.as_array(C.cast(c.overlaps, C.POINTER(C.c_uint32)), shape=(max(n, 1), 7))[:n].copy())
        self.res_matches = np.ctypeslib.as_array(c.res_matches, shape=(max(n, 1),))[:n].copy()

    @property
    def n_piles(self) -> int:
        return len(self.pile_ov_begin) - 1

    def equals(self, o: "PileSet") -> bool:
        return all(np.array_equal(getattr(self, n), getattr(o, n))
                   for n in ("pile_ov_begin", "pile_read", "pile_qlen", "overlaps", "res_matches")) and self.n_lines == o.n_lines

    def first_mismatch(self, o: "PileSet") -> str:
        for n in ("pile_ov_begin", "pile_read", "pile_qlen", "res_matches", "overlaps"):
            x, y = getattr(self, n), getattr(o, n)
            if x.shape != y.shape:
                return f"{n}: shapes {x.shape} / {y.shape}"
            if not np.array_equal(x, y):
                return f"{n}: first difference at {int(np.argmax((x != y).reshape(len(x), -1).any(axis=1)))}"
        return "" if self.n_lines == o.n_lines else f"n_lines {self.n_lines} / {o.n_lines}"

    def piles(self, store_off, store_bases, p0: int = 0, p1: int | None = None, **kw) -> "Piles":
        """Piles [p0, p1) over a read store, as cg_upload_piles takes them."""
        p1 = self.n_piles if p1 is None else p1
        a, b = int(self.pile_ov_begin[p0]), int(self.pile_ov_begin[p1])
        return Piles(store_off, store_bases, self.pile_read[p0:p1], self.pile_qlen[p0:p1],
                     self.pile_ov_begin[p0:p1 + 1] - self.pile_ov_begin[p0], self.overlaps[a:b], **kw)


def window_set_to_py(ws: cg_window_set, with_bases: bool = True):
    """Copies a cg_window_set out: -> (Batch, Reads, win_end)."""
    W, R = int(ws.batch.n_windows), int(ws.reads.n_reads)
    wsb = np.ctypeslib.as_array(ws.batch.win_seq_begin, shape=(W + 1,)).copy()
    S = int(wsb[W])
    soff = np.ctypeslib.as_array(ws.batch.seq_off, shape=(S + 1,)).copy()
    nb = int(soff[S])
    if with_bases and nb:
        bases = np.frombuffer(C.string_at(ws.batch.bases, nb), np.uint8).copy()
    else:
        bases = np.zeros(max(nb, 1), np.uint8)
    rwb = np.ctypeslib.as_array(ws.reads.read_win_begin, shape=(R + 1,)).copy()
    roff = np.ctypeslib.as_array(ws.reads.read_off, shape=(R + 1,)).copy()
    nr = int(roff[R])
    rbases = np.frombuffer(C.string_at(ws.reads.read_bases, nr), np.uint8).copy() if nr else np.zeros(1, np.uint8)
    wpos = np.ctypeslib.as_array(ws.reads.win_pos, shape=(max(W, 1),))[:W].copy()
    wend = np.ctypeslib.as_array(ws.win_end, shape=(max(W, 1),))[:W].copy()
    return (Batch(wsb, soff, bases), Reads(rwb, roff, rbases, wpos, int(ws.reads.window_size), int(ws.reads.window_overlap)), wend)


class Corrected:
    """Host copy of cg_corrected: the re-anchored reads (upper case = replaced by a consensus)."""

    def __init__(self, c: cg_corrected):
        R = int(c.n_reads)
        self.n_reads = R
        self.read_off = np.ctypeslib.as_array(c.read_off, shape=(R + 1,)).copy()
        nb = int(self.read_off[-1])
        self.bases = (np.ctypeslib.as_array(C.cast(c.bases, C.POINTER(C.c_uint8)), shape=(nb,)).copy()
                      if nb else np.zeros(0, np.uint8))

    def read(self, r: int) -> str:
        return self.bases[int(self.read_off[r]):int(self.read_off[r + 1])].tobytes().decode()

    def equals(self, other: "Corrected") -> bool:
        return (self.n_reads == other.n_reads and np.array_equal(self.read_off, other.read_off)
                and np.array_equal(self.bases, other.bases))

    def first_mismatch(self, other: "Corrected"):
        for r in range(min(self.n_reads, other.n_reads)):
            if self.read(r) != other.read(r):
                return r
        return None if self.n_reads == other.n_reads else min(self.n_reads, other.n_reads)

    def digest(self) -> str:
        import hashlib
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(self.read_off).tobytes())
        h.update(np.ascontiguousarray(self.bases).tobytes())
        return h.hexdigest()


def results_to_c(res: "Results") -> cg_results:
    """A cg_results view of a (numpy-backed) Results, for calls that take results as input."""
    def ptr(a, t):
        a = a if len(a) else np.zeros(1, a.dtype)
        return a.ctypes.data_as(C.POINTER(t))
    return cg_results(res.n_windows, ptr(res.cons_off, C.c_uint64), C.cast(ptr(res.cons, C.c_uint8), C.POINTER(C.c_char)),
                      ptr(res.status, C.c_uint8), ptr(res.solid_off, C.c_uint64), ptr(res.solid_kmer, C.c_uint32),
                      ptr(res.solid_count, C.c_uint32), None)


class Results:
    """Host view of cg_results.

    Results(r)            copies every array out of the C buffers (the caller frees them);
    Results(r, free=f)    wraps the C buffers without copying and calls f(r) when this object dies —
                          used by Corrector so that the (pinned, pooled) result buffers of the library
                          are handed back by cg_free_results instead of being copied."""

    def __init__(self, r: cg_results, free=None):
        W = int(r.n_windows)
        self.n_windows = W
        self._r, self._free = (r, free) if free is not None else (None, None)
        keep = (lambda a: a) if free is not None else (lambda a: a.copy())
        self.cons_off = keep(np.ctypeslib.as_array(r.cons_off, shape=(W + 1,)))
        nb = int(self.cons_off[-1])
        self.cons = (keep(np.ctypeslib.as_array(C.cast(r.cons, C.POINTER(C.c_uint8)), shape=(max(nb, 1),))[:nb])
                     if nb else np.zeros(0, np.uint8))
        self.status = keep(np.ctypeslib.as_array(r.status, shape=(max(W, 1),))[:W])
        self.solid_off = keep(np.ctypeslib.as_array(r.solid_off, shape=(W + 1,)))
        ns = int(self.solid_off[-1])
        if ns:
            self.solid_kmer = keep(np.ctypeslib.as_array(r.solid_kmer, shape=(ns,)))
            self.solid_count = keep(np.ctypeslib.as_array(r.solid_count, shape=(ns,)))
        else:
            self.solid_kmer = np.zeros(0, np.uint32)
            self.solid_count = np.zeros(0, np.uint32)

    def __del__(self):
        if getattr(self, "_free", None) is not None:
            f, r = self._free, self._r
            self._free = self._r = None
            for n in ("cons_off", "cons", "status", "solid_off", "solid_kmer", "solid_count"):
                setattr(self, n, None)
            try:
                f(r)
            except Exception:
                pass

    def detach(self) -> "Results":
        """Own copies of every array (the library buffers go back to its pool)."""
        if self._free is not None:
            for n in ("cons_off", "cons", "status", "solid_off", "solid_kmer", "solid_count"):
                setattr(self, n, getattr(self, n).copy())
            f, r = self._free, self._r
            self._free = self._r = None
            f(r)
        return self

    def consensus(self, w: int) -> str:
        return self.cons[int(self.cons_off[w]):int(self.cons_off[w + 1])].tobytes().decode()

    def solid(self, w: int):
        a, b = int(self.solid_off[w]), int(self.solid_off[w + 1])
        return list(zip(self.solid_kmer[a:b].tolist(), self.solid_count[a:b].tolist()))

    def equals(self, other: "Results") -> bool:
        return (self.n_windows == other.n_windows
                and np.array_equal(self.cons_off, other.cons_off)
                and np.array_equal(self.cons, other.cons)
                and np.array_equal(self.status, other.status)
                and np.array_equal(self.solid_off, other.solid_off)
                and np.array_equal(self.solid_kmer, other.solid_kmer)
                and np.array_equal(self.solid_count, other.solid_count))

    def first_mismatch(self, other: "Results"):
        """Index of the first window that differs (consensus, status or solid list), or None."""
        for w in range(min(self.n_windows, other.n_windows)):
            if (self.consensus(w) != other.consensus(w) or self.status[w] != other.status[w]
                    or self.solid(w) != other.solid(w)):
                return w
        return None if self.n_windows == other.n_windows else min(self.n_windows, other.n_windows)

    def stream_digests(self, running=None):
        """Slice-invariant digests of a window stream.  Six running sha256 (consensus lengths, consensus bytes, status; solid-list
        lengths, k-mers, counts): each is the hash of one array of the whole stream, so feeding the slices of a stream in order (pass
        the returned list back in) gives the same values as hashing the stream at once.  stream_digest_pair() folds them into the two
        values tests/golden/stream_digests.json holds for the reference's output."""
        import hashlib
        hs = running if running is not None else [hashlib.sha256() for _ in range(6)]
        n = self.n_windows
        lens = np.diff(self.cons_off[:n + 1].astype(np.int64)).astype(np.uint32)
        c0, c1 = int(self.cons_off[0]), int(self.cons_off[n])
        hs[0].update(lens.tobytes()); hs[1].update(np.ascontiguousarray(self.cons[c0:c1]).tobytes()); hs[2].update(np.ascontiguousarray(self.status[:n]).tobytes())
        ns = np.diff(self.solid_off[:n + 1].astype(np.int64)).astype(np.uint32)
        s0, s1 = int(self.solid_off[0]), int(self.solid_off[n])
        hs[3].update(ns.tobytes()); hs[4].update(np.ascontiguousarray(self.solid_kmer[s0:s1]).tobytes()); hs[5].update(np.ascontiguousarray(self.solid_count[s0:s1]).tobytes())
        return hs

    @staticmethod
    def stream_digest_pair(hs):
        """(consensus digest, solid-list digest) of a finished stream_digests() list."""
        import hashlib
        cons = hashlib.sha256("".join(h.hexdigest() for h in hs[:3]).encode()).hexdigest()
        solid = hashlib.sha256("".join(h.hexdigest() for h in hs[3:]).encode()).hexdigest()
        return cons, solid

    def digest(self) -> str:
        """sha256 over every output byte in a canonical order ("checksum of checksums" for big runs)."""
        import hashlib
        h = hashlib.sha256()
        for a in (self.cons_off, self.cons, self.status, self.solid_off, self.solid_kmer, self.solid_count):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()


def load_library(path: str) -> C.CDLL:
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    return C.CDLL(path, mode=C.RTLD_GLOBAL if False else C.RTLD_LOCAL)
