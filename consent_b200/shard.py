"""Multi-GPU sharding of the window stream (one process per GPU, torch.distributed).

Windows (and the piles they come from) are independent: the reference parallelises them over a thread
pool (src/CONSENT-correction.cpp:76-111, src/CONSENT-polishing.cpp:49-66) and never exchanges anything
between jobs.  So the data path has no collective: rank r owns a contiguous block of windows in input
order, runs the whole path on its own GPU, and the only exchange is the final ordered gather of the
corrected windows to rank 0 — the reference's "print results in submission order"
(src/CONSENT-correction.cpp:100-103) — over NCCL (NVLink 5 / NVSwitch) on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

from ._ffi import Batch, Corrected, Results, cg_results  # noqa: F401


def shard_range(n_windows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [w0, w1) of rank `rank`; block sizes differ by at most one."""
    base, rem = divmod(n_windows, world)
    w0 = rank * base + min(rank, rem)
    return w0, w0 + base + (1 if rank < rem else 0)


def shard_by_bases(batch: Batch, world: int) -> list[tuple[int, int]]:
    """Contiguous blocks balanced by bases (piles of very different depth): greedy prefix split."""
    per_win = np.array([int(batch.seq_off[int(batch.win_seq_begin[w + 1])]) - int(batch.seq_off[int(batch.win_seq_begin[w])])
                        for w in range(batch.n_windows)], dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(per_win)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(batch.n_windows)
    cuts = [min(max(c, 0), batch.n_windows) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


class _Flat:
    """Results as flat numpy arrays (what travels in the gather)."""

    def __init__(self, r: Results):
        self.cons_len = np.diff(r.cons_off).astype(np.int64)
        self.solid_len = np.diff(r.solid_off).astype(np.int64)
        self.cons = r.cons
        self.status = r.status
        self.solid_kmer = r.solid_kmer
        self.solid_count = r.solid_count


def _from_parts(parts) -> Results:
    out = Results.__new__(Results)
    out._r = out._free = None
    cons_len = np.concatenate([p.cons_len for p in parts]) if parts else np.zeros(0, np.int64)
    solid_len = np.concatenate([p.solid_len for p in parts]) if parts else np.zeros(0, np.int64)
    out.n_windows = int(len(cons_len))
    out.cons_off = np.concatenate([[0], np.cumsum(cons_len)]).astype(np.uint64)
    out.solid_off = np.concatenate([[0], np.cumsum(solid_len)]).astype(np.uint64)
    out.cons = np.concatenate([p.cons for p in parts]) if parts else np.zeros(0, np.uint8)
    out.status = np.concatenate([p.status for p in parts]) if parts else np.zeros(0, np.uint8)
    out.solid_kmer = np.concatenate([p.solid_kmer for p in parts]) if parts else np.zeros(0, np.uint32)
    out.solid_count = np.concatenate([p.solid_count for p in parts]) if parts else np.zeros(0, np.uint32)
    return out


_PINNED = {}          # (tag, device index) -> pinned host staging tensor, grown on demand and reused across calls


def _pinned(tag: str, nbytes: int):
    import torch
    key = (tag, torch.cuda.current_device() if torch.cuda.is_available() else -1)
    t = _PINNED.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1) + max(nbytes, 1) // 4, dtype=torch.uint8, pin_memory=True)
        _PINNED[key] = t
    return t


class GatheredParts:
    """What rank 0 holds after gather_results(..., concat=False): one part per rank, in input order, as views into the
    receive buffers (nothing is concatenated; valid until the next gather).  Enough for the ordered writer:
    `for w in range(g.n_windows): g.consensus(w)`."""

    def __init__(self, flats):
        self.parts = flats
        self.first = np.concatenate([[0], np.cumsum([len(p.cons_len) for p in flats])]).astype(np.int64)
        self.n_windows = int(self.first[-1])
        self._offs = [np.concatenate([[0], np.cumsum(p.cons_len)]) for p in flats]

    def consensus(self, w: int) -> str:
        r = int(np.searchsorted(self.first, w, side="right")) - 1
        i = w - int(self.first[r])
        o = self._offs[r]
        return self.parts[r].cons[int(o[i]):int(o[i + 1])].tobytes().decode()

    def status(self, w: int) -> int:
        r = int(np.searchsorted(self.first, w, side="right")) - 1
        return int(self.parts[r].status[w - int(self.first[r])])


def gather_results(local: Results, device=None, group=None, with_solid: bool = True, concat: bool = True):
    """Ordered gather of every rank's results to rank 0 (variable length: sizes first, then padded payloads).

    Returns the concatenated Results on rank 0, None elsewhere.  `device`: torch device of the payload tensors
    ("cuda:N" under NCCL, "cpu" under gloo).  with_solid=False gathers what the reference finally emits — the
    corrected sequences and their status — and leaves the solid k-mer lists on the rank that computed them: their only
    consumer is that rank's own re-anchoring of the window consensuses (src/correctionAlignment.cpp:6-15,103-104),
    and at 150-deep piles they are 50x the bytes of the consensuses.  concat=False returns GatheredParts (views, no
    concatenation on rank 0)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    on_gpu = torch.device(device).type == "cuda"
    f = _Flat(local)
    if not with_solid:
        f.solid_len = np.zeros_like(f.solid_len)
        f.solid_kmer = np.zeros(0, np.uint32)
        f.solid_count = np.zeros(0, np.uint32)
    # one byte payload per rank: [cons_len i64][solid_len i64][status u8][cons u8][solid_kmer u32][solid_count u32]
    chunks = [np.ascontiguousarray(c).reshape(-1).view(np.uint8) for c in
              (f.cons_len, f.solid_len, f.status, f.cons, f.solid_kmer, f.solid_count)] if local.n_windows else []
    nbytes = int(sum(len(c) for c in chunks))
    meta = torch.tensor([local.n_windows, len(f.cons), len(f.solid_kmer), nbytes], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = [[int(x) for x in m.tolist()] for m in metas]
    max_len = max(max(m[3] for m in metas), 1)
    buf = torch.empty(max_len, dtype=torch.uint8, device=device)
    o = 0
    for c in chunks:                                    # straight from the (pinned) result buffers into the send buffer
        if len(c):
            buf[o:o + len(c)].copy_(torch.from_numpy(c), non_blocking=on_gpu)
        o += len(c)
    gathered = [torch.empty(max_len, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0, group=group)
    if on_gpu:
        # the send buffer was filled by non-blocking copies out of library-owned pinned result buffers: they must have been read
        # before the caller may drop `local` (cg_free_results hands them back to the pool for the next call's downloads)
        torch.cuda.current_stream().synchronize()
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        nw, nc, ns, nb = metas[r]
        if on_gpu:
            host = _pinned(f"recv{r}", nb)[:nb]
            host.copy_(gathered[r][:nb], non_blocking=True)
        else:
            host = gathered[r][:nb]
        parts.append((nw, nc, ns, host))
    if on_gpu:
        torch.cuda.current_stream().synchronize()
    flats = []
    for nw, nc, ns, host in parts:
        raw = host.numpy()
        p = _Flat.__new__(_Flat)
        o = 0
        p.cons_len = raw[o:o + 8 * nw].view(np.int64); o += 8 * nw
        p.solid_len = raw[o:o + 8 * nw].view(np.int64); o += 8 * nw
        p.status = raw[o:o + nw]; o += nw
        p.cons = raw[o:o + nc]; o += nc
        p.solid_kmer = raw[o:o + 4 * ns].view(np.uint32); o += 4 * ns
        p.solid_count = raw[o:o + 4 * ns].view(np.uint32); o += 4 * ns
        flats.append(p)
    return _from_parts(flats) if concat else GatheredParts(flats)


# ---- the whole pipeline (BASELINE config 4): read piles sharded over the GPUs, corrected reads gathered in PAF order ---------
def shard_piles(pile_qlen, pile_ov_begin, world: int) -> list[tuple[int, int]]:
    """Contiguous blocks of piles in PAF order, one per rank, balanced by the bases their windows will hold:
    a pile of n overlaps over a read of L bases cuts about L / step windows of about n * coverage fraction sequences, so the
    weight of pile p is qlen[p] * (n_overlaps[p] + 1).  Greedy prefix split, like shard_by_bases."""
    qlen = np.asarray(pile_qlen, dtype=np.int64)
    n_ov = np.diff(np.asarray(pile_ov_begin, dtype=np.int64))
    P = len(qlen)
    cum = np.concatenate([[0], np.cumsum(qlen * (n_ov + 1))])
    total = int(cum[-1])
    cuts = [0] + [int(np.searchsorted(cum, total * r / world, side="left")) for r in range(1, world)] + [P]
    cuts = [min(max(c, 0), P) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def gather_corrected(local: Corrected, device=None, group=None):
    """Ordered gather of every rank's corrected reads (cg_reanchor_reads / cg_finish_reads output) to rank 0: the one
    collective of the pipeline — rank 0 writes the FASTA in PAF order (src/CONSENT-correction.cpp:100-103).  Returns the
    concatenated Corrected on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    on_gpu = torch.device(device).type == "cuda"
    lens = np.diff(local.read_off).astype(np.int64)
    chunks = [lens.view(np.uint8), np.ascontiguousarray(local.bases).view(np.uint8)]
    nbytes = int(sum(len(c) for c in chunks))
    meta = torch.tensor([local.n_reads, len(local.bases), nbytes], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = [[int(x) for x in m.tolist()] for m in metas]
    max_len = max(max(m[2] for m in metas), 1)
    buf = torch.empty(max_len, dtype=torch.uint8, device=device)
    o = 0
    for c in chunks:
        if len(c):
            buf[o:o + len(c)].copy_(torch.from_numpy(c), non_blocking=on_gpu)
        o += len(c)
    gathered = [torch.empty(max_len, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0, group=group)
    if rank != 0:
        return None
    all_lens, all_bases = [], []
    for r in range(world):
        nr, nb, tot = metas[r]
        raw = gathered[r][:tot].cpu().numpy()
        all_lens.append(raw[:8 * nr].view(np.int64))
        all_bases.append(raw[8 * nr:8 * nr + nb])
    out = Corrected.__new__(Corrected)
    lens = np.concatenate(all_lens) if all_lens else np.zeros(0, np.int64)
    out.n_reads = int(len(lens))
    out.read_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    out.bases = np.concatenate(all_bases) if all_bases else np.zeros(0, np.uint8)
    return out
