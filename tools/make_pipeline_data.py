#!/usr/bin/env python
"""Seeded inputs of the two pipeline-scale configurations of BASELINE.json, written as FASTA:

  config4 DIR   5 Mb uniform random genome, fixed 8 kb reads placed uniformly on both strands at 30x (18 750 reads), PacBio channel
                (total error 15 %, sub : ins : del = 10 : 60 : 30)                        -> DIR/reads.fasta
  config5 DIR   50 contigs x 100 kb: truth uniform random, contig = truth with 1 % errors (ONT mix), 30x of 8 kb ONT-channel reads
                (10 %, 40 : 20 : 40) drawn from the truth                                  -> DIR/contigs.fasta, DIR/reads.fasta
  --scale F     shrinks genome / contig count by F (tests use smaller instances of the same generator)

numpy's PCG64 streams are reproducible across machines, so the GPU box regenerates the very files the golden md5s under
tests/golden/pipeline_md5.json were computed from (by tests/golden/make_pipeline_golden.sh, with the unmodified reference binaries)."""
import argparse
import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", np.uint8)
COMP = np.array([3, 2, 1, 0], np.uint8)
PROFILES = {"PB": (0.15, 0.10, 0.60), "ONT": (0.10, 0.40, 0.20)}      # total error, share of substitutions, share of insertions


def channel(rng, codes: np.ndarray, profile: str) -> np.ndarray:
    """i.i.d. per-position error channel: substitution = a different base, insertion = a random base before the
    position, deletion = skip."""
    e, p_sub, p_ins = PROFILES[profile]
    u = rng.random(len(codes))
    sub = u < e * p_sub
    ins = (u >= e * p_sub) & (u < e * (p_sub + p_ins))
    dele = (u >= e * (p_sub + p_ins)) & (u < e)
    out = codes.copy()
    out[sub] = (out[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) % 4
    keep = ~dele
    reps = keep.astype(np.int64) + ins                          # an inserted base in front of a kept or deleted position
    total = int(reps.sum())
    res = np.empty(total, np.uint8)
    pos = np.cumsum(reps) - reps                                # first output slot of every position
    res[pos[ins]] = rng.integers(0, 4, int(ins.sum()), dtype=np.uint8)
    kp = pos + ins
    res[kp[keep]] = out[keep]
    return res


def write_fasta(path, records):
    with open(path, "wb") as f:
        for name, codes in records:
            f.write(b">" + name.encode() + b"\n")
            f.write(ACGT[codes].tobytes())
            f.write(b"\n")


def sample_reads(rng, truth: np.ndarray, n_reads: int, read_len: int, profile: str, prefix: str):
    for i in range(n_reads):
        beg = int(rng.integers(0, len(truth) - read_len + 1))
        seg = truth[beg:beg + read_len]
        if rng.integers(0, 2):
            seg = COMP[seg][::-1]
        yield f"{prefix}{i}", channel(rng, np.ascontiguousarray(seg), profile)


def config4(out: str, scale: float, seed: int = 42):
    rng = np.random.default_rng(seed)
    glen = int(5_000_000 / scale)
    genome = rng.integers(0, 4, glen, dtype=np.uint8)
    n_reads = glen * 30 // 8000
    write_fasta(os.path.join(out, "reads.fasta"), sample_reads(rng, genome, n_reads, 8000, "PB", "read_"))


def config5(out: str, scale: float, seed: int = 43):
    rng = np.random.default_rng(seed)
    n_contigs = max(1, int(50 / scale))
    contigs, reads = [], []
    for c in range(n_contigs):
        truth = rng.integers(0, 4, 100_000, dtype=np.uint8)
        e_save = PROFILES["ONT"]
        PROFILES["draft"] = (0.01, e_save[1], e_save[2])
        contigs.append((f"contig_{c}", channel(rng, truth, "draft")))
        reads.extend(sample_reads(rng, truth, 100_000 * 30 // 8000, 8000, "ONT", f"c{c}_read_"))
    order = rng.permutation(len(reads))                         # reads of all contigs shuffled, as a sequencing run delivers them
    write_fasta(os.path.join(out, "contigs.fasta"), contigs)
    write_fasta(os.path.join(out, "reads.fasta"), (reads[i] for i in order))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=("config4", "config5"))
    ap.add_argument("out")
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    (config4 if a.config == "config4" else config5)(a.out, a.scale)
