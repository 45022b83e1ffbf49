"""Window extraction (phase A of processRead — getAlignmentWindowsPositions / getAlignmentWindowsSequences,
reference src/alignmentWindows.cpp:5-149) on the device: cg_upload_piles / cg_download_windows.

CPU suite: the oracle (oracle/extract_oracle.c) against the unmodified reference functions (oracle/_ref) on seeded piles and on
the overlaps of the shipped example; the kernel sources (consent_b200/csrc/k_extract.cuh) on the SIMT emulator against the oracle.
GPU suite (-m gpu): the same through libconsent_b200.so on a B200, and the whole chain piles -> corrected reads."""
import gzip
import os

import numpy as np
import pytest

from consent_b200._ffi import Params, Piles
from consent_b200.engine import ConsentError
from consent_b200.synth import synth_piles

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SEEDED = (
    dict(n_reads=60, genome_len=30000, read_len=3000, seed=1),
    dict(n_reads=100, genome_len=40000, read_len=4000, seed=2, max_support=20),
    dict(n_reads=150, genome_len=50000, read_len=2500, seed=3, profile="ONT", min_support=8, window_size=300, window_overlap=0),
    dict(n_reads=40, genome_len=200000, read_len=3000, seed=4),                      # sparse: most piles have no window
    dict(n_reads=80, genome_len=20000, read_len=1500, seed=5, min_support=1, window_size=700, window_overlap=120),
)


def example_piles() -> Piles:
    """The 20 read piles of tests/golden/example_small.paf.gz in the layout of cg_piles (overlaps in PAF order)."""
    names, seqs = {}, []
    with gzip.open(os.path.join(ROOT, "tests", "golden", "example_small_reads.fasta.gz"), "rt") as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                names[line[1:]] = len(seqs)
            elif line:
                seqs.append(line)
    off = np.zeros(len(seqs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    store = np.frombuffer("".join(seqs).encode(), np.uint8)
    pile_read, pile_qlen, ovb, ov, cur = [], [], [0], [], None
    with gzip.open(os.path.join(ROOT, "tests", "golden", "example_small.paf.gz"), "rt") as f:
        for line in f:
            c = line.rstrip("\n").split("\t")
            if c[0] != cur:
                if cur is not None:
                    ovb.append(len(ov))
                cur = c[0]
                pile_read.append(names[c[0]])
                pile_qlen.append(int(c[1]))
            ov.append((names[c[5]], 0 if c[4] == "+" else 1, int(c[2]), int(c[3]) - 1, int(c[7]), int(c[8]) - 1, int(c[6])))
    ovb.append(len(ov))
    return Piles(off, store, pile_read, pile_qlen, ovb, np.array(ov, np.uint32))


def assert_same_windows(got, want, what):
    (b1, r1, e1), (b2, r2, e2) = got, want
    assert b1.n_windows == b2.n_windows, f"{what}: {b1.n_windows} windows vs {b2.n_windows}"
    for name, x, y in (("win_pos", r1.win_pos, r2.win_pos), ("win_end", e1, e2), ("read_win_begin", r1.read_win_begin, r2.read_win_begin),
                       ("win_seq_begin", b1.win_seq_begin, b2.win_seq_begin), ("seq_off", b1.seq_off, b2.seq_off),
                       ("bases", b1.bases[:b1.n_bases], b2.bases[:b2.n_bases]), ("read_off", r1.read_off, r2.read_off),
                       ("read_bases", r1.read_bases[:int(r1.read_off[-1])], r2.read_bases[:int(r2.read_off[-1])])):
        if not np.array_equal(x, y):
            i = int(np.argmax(x[:min(len(x), len(y))] != y[:min(len(x), len(y))])) if len(x) and len(y) else 0
            raise AssertionError(f"{what}: {name} differs (first at {i}; lengths {len(x)} / {len(y)})")


# ------------------------------------------------------------------------------------------------ CPU: the oracle is pinned
def test_oracle_matches_reference(oracle, reference):
    n = 0
    for kw in SEEDED + (dict(n_reads=400, genome_len=100000, read_len=5000, seed=6, max_support=150),):
        p = synth_piles(**kw)
        want = reference.extract_windows(p)
        assert_same_windows(oracle.extract_windows(p), want, f"oracle vs reference, seed {kw['seed']}")
        n += want[0].n_windows
    assert n > 5000
    p = example_piles()
    want = reference.extract_windows(p)
    assert want[0].n_windows > 250 and want[0].n_seqs > 2500
    assert_same_windows(oracle.extract_windows(p), want, "oracle vs reference, shipped example")


def test_oracle_reproduces_the_example_fixture(oracle):
    """tests/golden/example_windows.txt.gz holds piles cut by the reference's own binary code path from the full example PAF
    (reads 21..); the small PAF covers reads 1..20, so only the shapes can be compared: every window is windowSize long, its
    pile starts with the read's own bases, and sequences shorter than k are absent."""
    p = example_piles()
    batch, reads, wend = oracle.extract_windows(p)
    for r in range(reads.n_reads):
        read = reads.read(r)
        for w in range(int(reads.read_win_begin[r]), int(reads.read_win_begin[r + 1])):
            pile = batch.pile(w)
            assert len(pile[0]) == 500 and pile[0] == read[int(reads.win_pos[w]):int(wend[w]) + 1]
            assert all(len(s) >= 9 for s in pile[1:])


def test_oracle_rejects_what_the_reference_cannot_survive(oracle):
    p = synth_piles(n_reads=30, genome_len=10000, read_len=2000, seed=7)
    bad = Piles(p.store_off, p.store_bases, p.pile_read, p.pile_qlen, p.pile_ov_begin, p.overlaps.copy())
    bad.overlaps[0, 3] = bad.pile_qlen[0] + 10                                    # qEnd beyond qLength: getCoverages writes past its array
    with pytest.raises(RuntimeError):
        oracle.extract_windows(bad)


# ------------------------------------------------------------------------------------------------ CPU: kernel sources, emulated
@pytest.mark.parametrize("i", range(len(SEEDED)))
def test_emulated_kernels_match_oracle(emu, oracle, i):
    p = synth_piles(**SEEDED[i])
    cor = emu()
    cor.upload_piles(p)
    assert_same_windows(cor.download_windows(), oracle.extract_windows(p), f"emulated kernels vs oracle, seed {SEEDED[i]['seed']}")
    assert cor.extract_stats()["pile_bytes"] == oracle.extract_windows(p)[0].n_bases


def test_emulated_kernels_on_the_shipped_example(emu, oracle):
    p = example_piles()
    cor = emu()
    cor.upload_piles(p)
    assert_same_windows(cor.download_windows(), oracle.extract_windows(p), "emulated kernels vs oracle, shipped example")


def test_emulated_chain_piles_to_corrected_reads(emu, oracle):
    """cg_upload_piles -> cg_run -> cg_download -> cg_reanchor_reads on one handle = oracle extraction + windows + re-anchoring."""
    p = synth_piles(n_reads=14, genome_len=5000, read_len=1800, seed=9, max_support=8)
    cor = emu()
    cor.upload_piles(p)
    cor.run()
    res = cor.download()
    batch, reads, _ = cor.download_windows()
    got = cor.reanchor_reads(batch, res, reads)
    ob, ord_, _ = oracle.extract_windows(p)
    ores, _ = oracle.correct_windows(ob, threads=8)
    want, _ = oracle.reanchor_reads(ob, ores, ord_, threads=4)
    assert res.equals(ores)
    assert got.equals(want)
    assert batch.n_windows >= 20


def test_emulated_kernels_reject_bad_overlaps(emu):
    p = synth_piles(n_reads=30, genome_len=10000, read_len=2000, seed=7)
    bad = Piles(p.store_off, p.store_bases, p.pile_read, p.pile_qlen, p.pile_ov_begin, p.overlaps.copy())
    bad.overlaps[0, 3] = bad.pile_qlen[0] + 10
    with pytest.raises(ConsentError):
        emu().upload_piles(bad)
    bad2 = Piles(p.store_off, p.store_bases, p.pile_read, p.pile_qlen, p.pile_ov_begin, p.overlaps.copy())
    bad2.overlaps[0, 0] = p.n_store + 5                                            # target read outside the store
    with pytest.raises(ConsentError):
        emu().upload_piles(bad2)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_matches_oracle(gpu, oracle):
    cor = gpu()
    for kw in SEEDED + (dict(n_reads=1200, genome_len=300000, read_len=6000, seed=11, max_support=150),):
        p = synth_piles(**kw)
        cor.upload_piles(p)
        assert_same_windows(cor.download_windows(), oracle.extract_windows(p), f"GPU vs oracle, seed {kw['seed']}")
    p = example_piles()
    cor.upload_piles(p)
    assert_same_windows(cor.download_windows(), oracle.extract_windows(p), "GPU vs oracle, shipped example")


@pytest.mark.gpu
def test_gpu_matches_reference_if_present(gpu, reference):
    cor = gpu()
    p = synth_piles(n_reads=600, genome_len=150000, read_len=5000, seed=12)
    cor.upload_piles(p)
    assert_same_windows(cor.download_windows(), reference.extract_windows(p), "GPU vs the unmodified reference")


@pytest.mark.gpu
def test_gpu_chain_piles_to_corrected_reads(gpu, oracle):
    cor = gpu()
    p = synth_piles(n_reads=500, genome_len=100000, read_len=5000, seed=13, max_support=40)
    cor.upload_piles(p)
    cor.run()
    res = cor.download()
    batch, reads, _ = cor.download_windows()
    got = cor.reanchor_reads(batch, res, reads)
    ob, ord_, _ = oracle.extract_windows(p)
    ores, _ = oracle.correct_windows(ob, threads=os.cpu_count() or 4)
    want, _ = oracle.reanchor_reads(ob, ores, ord_, threads=os.cpu_count() or 4)
    assert res.equals(ores)
    assert got.equals(want)
    assert batch.n_windows > 4000


# ------------------------------------------------------------------------------------------------ polishing-shaped piles
def contig_pile(n_reads: int, contig_len: int, read_len: int, seed: int) -> Piles:
    """ONE pile whose query is a long 'contig' with n_reads short reads mapped all along it (CONSENT-polish: a contig's pile holds
    thousands of overlaps, a window sees its local coverage; maxSupport 20000, reference CONSENT-polish:43)."""
    rng = np.random.default_rng(seed)
    contig = rng.integers(0, 4, contig_len)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    store, off, ov = [acgt[contig]], [0, contig_len], []
    for r in range(n_reads):
        beg = int(rng.integers(0, contig_len - read_len))
        seq = contig[beg:beg + read_len].copy()
        flip = rng.random(read_len) < 0.08
        seq[flip] = (seq[flip] + rng.integers(1, 4, int(flip.sum()))) % 4
        strand = int(rng.integers(0, 2))
        bases = acgt[seq]
        if strand:
            bases = acgt[3 - seq][::-1]
        store.append(bases); off.append(off[-1] + read_len)
        ov.append((r + 1, strand, beg, beg + read_len - 1, 0, read_len - 1, read_len))
    return Piles(np.array(off, np.uint64), np.concatenate(store), np.array([0], np.uint32), np.array([contig_len], np.uint32),
                 np.array([0, n_reads], np.uint32), np.array(ov, np.uint32), 1, 500, 50)


def test_emulated_kernels_cut_a_pile_of_more_overlaps_than_a_window_may_hold(emu, oracle):
    """5000 overlaps in one pile (> CG_N_MAX = 4095 sequences per WINDOW): every window only sees ~40 of them and must extract."""
    p = contig_pile(5000, 60000, 500, seed=21)
    cor = emu()
    cor.upload_piles(p)
    got = cor.download_windows()
    assert_same_windows(got, oracle.extract_windows(p), "deep contig pile")
    assert got[0].n_windows > 100 and got[0].n_seqs / got[0].n_windows < 200


def test_emulated_resident_store_gives_the_same_windows(emu, oracle):
    p = synth_piles(n_reads=40, genome_len=12000, read_len=2000, seed=8)
    cor = emu()
    cor.set_read_store(p.store_off, p.store_bases)
    half = p.n_piles // 2
    want = oracle.extract_windows(p)
    for lo, hi in ((0, half), (half, p.n_piles)):                       # two batches of piles against one resident store
        o0, o1 = int(p.pile_ov_begin[lo]), int(p.pile_ov_begin[hi])
        sub = Piles(p.store_off, p.store_bases, p.pile_read[lo:hi], p.pile_qlen[lo:hi], p.pile_ov_begin[lo:hi + 1] - o0, p.overlaps[o0:o1],
                    p.min_support, p.window_size, p.window_overlap)
        cor.upload_piles_resident(sub)
        assert_same_windows(cor.download_windows(), oracle.extract_windows(sub), f"resident store, piles {lo}..{hi}")
    assert want[0].n_windows > 0
    with pytest.raises(ConsentError):
        emu().upload_piles_resident(p)                                   # no store set on that handle
