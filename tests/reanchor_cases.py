"""Deterministic inputs of the re-anchoring path (test infrastructure): seeded reads with their windows
(consent_b200/host/synth.h) and hand-made edge cases.  A case = (name, window Batch, Reads, Params)."""
from __future__ import annotations

import random

import numpy as np

from consent_b200._ffi import Batch, Params, Reads
from consent_b200.synth import synth_reads
from tests.cases import _mutate, _rand_seq

SEEDED = (  # name, n_reads, n_seqs, truth_len, seed, profile, window_size, window_overlap, thin_every
    ("pb_n8", 6, 8, 2500, 1, "PB", 500, 50, 0),
    ("pb_n4_thin3", 6, 4, 3000, 2, "PB", 500, 50, 3),
    ("pb_n20", 3, 20, 2200, 3, "PB", 500, 50, 0),
    ("pb_n3_thin2", 8, 3, 1800, 4, "PB", 500, 50, 2),
    ("ont_n6", 6, 6, 2600, 5, "ONT", 500, 50, 5),
    ("pb_long_windows", 4, 5, 3200, 6, "PB", 800, 60, 0),          # consensus > 640 rows: two bands in the kernel
    ("pb_small_windows", 5, 6, 1500, 7, "PB", 200, 30, 4),
    ("pb_no_overlap", 4, 6, 2000, 8, "PB", 400, 0, 0),
)


def seeded_cases():
    for name, n_reads, n_seqs, tl, seed, prof, ws, ov, thin in SEEDED:
        batch, reads = synth_reads(n_reads, n_seqs, truth_len=tl, seed=seed, profile=prof, window_size=ws,
                                   window_overlap=ov, thin_every=thin, thin_seqs=1)
        yield name, batch, reads, Params()


def _read_case(rng, truth_len, windows, piles_of):
    """One read = truth through the PB channel; windows = [(start, length)], piles_of(i, template, truth) -> pile."""
    truth = _rand_seq(rng, truth_len)
    read = _mutate(rng, truth, 0.12)
    piles, pos = [], []
    for i, (a, n) in enumerate(windows):
        tpl = read[a:a + n]
        piles.append(piles_of(i, tpl, rng))
        pos.append(a)
    return read, piles, pos


def edge_cases():
    rng = random.Random(77)
    reads, piles, pos, per_read = [], [], [], []

    def deep(i, tpl, rng):
        return [tpl] + [_mutate(rng, tpl, 0.1) for _ in range(7)]

    # 0: a read without windows (CONSENT-correction.cpp:23-25 -> empty output)
    reads.append(_rand_seq(rng, 700)); per_read.append(0)
    # 1: a single window
    r, p, q = _read_case(rng, 900, [(100, 500)], deep)
    reads.append(r); piles += p; pos += q; per_read.append(1)
    # 2: a window whose pile is the template alone (raw template returned), next to corrected ones
    r, p, q = _read_case(rng, 1600, [(0, 500), (450, 500), (900, 500)], lambda i, t, g: [t] if i == 1 else deep(i, t, g))
    reads.append(r); piles += p; pos += q; per_read.append(3)
    # 3: a template shorter than k: the consensus is shorter than merSize, the template is aligned but nothing is replaced
    r, p, q = _read_case(rng, 1500, [(0, 500), (620, 6), (700, 500)], lambda i, t, g: [t, t[:4]] if i == 1 else deep(i, t, g))
    reads.append(r); piles += p; pos += q; per_read.append(3)
    # 4: duplicated window (the "last window" of alignmentWindows.cpp:58-80 can repeat the previous one): full overlap
    r, p, q = _read_case(rng, 1200, [(0, 500), (450, 500), (450, 500)], deep)
    reads.append(r); piles += p; pos += q; per_read.append(3)
    # 5: lower-case read (alignConsensus lower-cases it anyway, :57-58)
    r, p, q = _read_case(rng, 1100, [(0, 500), (450, 500)], deep)
    reads.append(r.lower()); piles += p; pos += q; per_read.append(2)
    # 6: shallow piles that leave weak (lower-case) stretches in the consensus, heavy overlaps
    r, p, q = _read_case(rng, 2000, [(0, 500), (300, 500), (600, 500), (900, 500), (1200, 500)],
                         lambda i, t, g: [t] + [_mutate(g, t, 0.15) for _ in range(3)])
    reads.append(r); piles += p; pos += q; per_read.append(5)
    batch = Batch.from_piles(piles)
    rd = Reads.from_lists(reads, per_read, pos, 500, 50)
    return [("edges", batch, rd, Params())]


def all_cases():
    yield from seeded_cases()
    yield from edge_cases()
