"""Deterministic test piles: seeded synthetic windows plus hand-made edge cases (test infrastructure)."""
from __future__ import annotations

import random

from consent_b200._ffi import Batch
from consent_b200.synth import synth_windows


def _rand_seq(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def _mutate(rng, s, err, p_sub=0.1, p_ins=0.6):
    out = []
    for ch in s:
        u = rng.random()
        if u >= err:
            out.append(ch)
        else:
            v = u / err
            if v < p_sub:
                out.append(rng.choice([c for c in "ACGT" if c != ch]))
            elif v < p_sub + p_ins:
                out.append(rng.choice("ACGT"))
                out.append(ch)
    return "".join(out)


def edge_piles(seed: int = 1):
    """Edge cases of the domain (name, pile).  pile[0] is the template."""
    rng = random.Random(seed)
    truth = _rand_seq(rng, 500)
    cases = []
    cases.append(("single_sequence", [truth]))
    cases.append(("two_identical", [truth, truth]))
    cases.append(("eight_identical", [truth] * 8))
    cases.append(("template_shorter_than_k", ["ACGTAC"] + [_mutate(rng, truth, 0.1) for _ in range(5)]))
    cases.append(("all_shorter_than_k", ["ACGTAC", "ACGTTC", "ACG", "A"]))
    cases.append(("some_reads_shorter_than_k", [truth] + ["ACGT", "GATTACA"] + [_mutate(rng, truth, 0.12) for _ in range(6)]))
    cases.append(("unrelated_reads_no_anchor", [_rand_seq(rng, 300) for _ in range(6)]))
    cases.append(("unrelated_short_no_anchor", [_rand_seq(rng, 60) for _ in range(12)]))
    cases.append(("homopolymer", ["A" * 200] * 6))
    cases.append(("homopolymer_mixed", ["A" * 120 + truth[:100]] + [_mutate(rng, "A" * 120 + truth[:100], 0.1) for _ in range(9)]))
    cases.append(("dinucleotide_repeat", ["AC" * 150] + [_mutate(rng, "AC" * 150, 0.08) for _ in range(7)]))
    cases.append(("tandem_repeat", [(truth[:40] * 8)] + [_mutate(rng, truth[:40] * 8, 0.1) for _ in range(10)]))
    cases.append(("low_error_deep", [truth] + [_mutate(rng, truth, 0.02) for _ in range(40)]))
    cases.append(("high_error", [_mutate(rng, truth, 0.3) for _ in range(25)]))
    cases.append(("ragged_lengths", [truth] + [_mutate(rng, truth[a:b], 0.1) for a, b in ((0, 250), (100, 500), (200, 300), (0, 500), (50, 450), (300, 500), (0, 120), (10, 490))]))
    cases.append(("ont_profile", [_mutate(rng, truth, 0.10, 0.4, 0.2) for _ in range(30)]))
    cases.append(("long_window_1000", [_mutate(rng, _rand_seq(random.Random(seed + 7), 1000), 0.12) for _ in range(12)]))
    cases.append(("short_window_30", [_mutate(rng, truth[:30], 0.1) for _ in range(9)]))
    cases.append(("template_is_outlier", [_rand_seq(rng, 480)] + [_mutate(rng, truth, 0.1) for _ in range(10)]))
    cases.append(("two_haplotypes", [_mutate(rng, truth, 0.05) for _ in range(8)] + [_mutate(rng, truth[:200] + _rand_seq(rng, 40) + truth[240:], 0.05) for _ in range(8)]))
    cases.append(("weak_ends", [truth] + [_mutate(rng, truth[30:470], 0.05) for _ in range(12)]))
    return cases


def counter_width_piles(seed: int = 11):
    """Piles around the width of k_index's byte counters: a 9-mer seen exactly 255 times (stays on the byte path), 256 and 257 times
    (the window is counted again with 32-bit counters), and one seen 255 times next to one seen 300 times in the other pass's key range."""
    rng = random.Random(seed)
    truth = _rand_seq(rng, 300)

    def pile(runs):                       # runs: list of (base, run length, number of reads); a run of n gives n - 8 9-mers
        reads = [truth]
        for base, run, n in runs:
            for _ in range(n):
                cut = rng.randrange(20, 280)
                reads.append(_mutate(rng, truth[:cut], 0.05) + "CG"[base == "C"] + base * run + "CG"[base == "C"] + _mutate(rng, truth[cut:], 0.05))
        return reads
    return [("kmer_255_times", pile([("A", 13, 51)])),                       # 51 x 5
            ("kmer_256_times", pile([("A", 16, 32)])),                       # 32 x 8
            ("kmer_257_times", pile([("A", 16, 32), ("A", 9, 1)])),
            ("kmer_255_and_300_times_two_passes", pile([("A", 13, 51), ("T", 18, 30)]))]


def edge_batch(seed: int = 1) -> Batch:
    return Batch.from_piles([p for _, p in edge_piles(seed)])


SEEDED = (  # (n_seqs, n_windows, seed, profile)
    (1, 6, 3, "PB"), (2, 12, 3, "PB"), (3, 12, 3, "PB"), (5, 10, 3, "PB"), (8, 12, 3, "PB"), (20, 8, 3, "PB"),
    (47, 4, 3, "PB"), (150, 3, 3, "PB"), (20, 6, 4, "ONT"), (60, 3, 4, "ONT"),
)


def seeded_batches():
    for n, w, seed, prof in SEEDED:
        yield f"synth_{prof}_n{n}_w{w}_s{seed}", synth_windows(w, n, seed=seed, profile=prof)


def concat(batches) -> Batch:
    piles = []
    for b in batches:
        for w in range(b.n_windows):
            piles.append(b.pile(w))
    return Batch.from_piles(piles)
