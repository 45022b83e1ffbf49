"""PAF ingest (getNextReadPile — Overlap(line), grouping, std::sort(rbegin, rend) by resMatches, cut to
maxSupport; reference src/alignmentPiles.cpp:22-58, src/Overlap.h:26-96) and the post-filters (rank 4: trimRead / dropRead,
src/utils.cpp:60-73,96-128) on the device: cg_ingest_paf / cg_finish_reads.

CPU suite: the oracle (oracle/ingest_oracle.c) against the unmodified reference (oracle/_ref: its own getNextReadPile, its own
std::sort call, its own trimRead / dropRead), then the kernel sources (consent_b200/csrc/k_ingest.cuh, k_reanchor.cuh) on the SIMT
emulator against the oracle.  GPU suite (-m gpu): the same through libconsent_b200.so on a B200, and the whole chain
PAF text -> FASTA sequence lines."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

from consent_b200._ffi import Corrected, ReadNames
from consent_b200.engine import ConsentError
from consent_b200.synth import synth_paf, synth_piles

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PAFS = (
    (dict(n_reads=120, genome_len=30000, read_len=3000, seed=21, max_support=4000), dict(seed=1), (150, 20)),
    (dict(n_reads=200, genome_len=40000, read_len=4000, seed=22, max_support=4000), dict(seed=2, tie_range=3, blank_every=7), (17, 1)),
    (dict(n_reads=90, genome_len=9000, read_len=2500, seed=23, max_support=4000), dict(seed=3, tie_range=1, extra_columns=False), (16, 40)),
    (dict(n_reads=60, genome_len=200000, read_len=3000, seed=24, max_support=4000), dict(seed=4), (150,)),          # sparse: tiny piles
)


def example_paf():
    """The PAF of the 20 read piles of the shipped example (vendored minimap2, flags of CONSENT-correct:185) + its read names."""
    names = []
    with gzip.open(os.path.join(ROOT, "tests", "golden", "example_small_reads.fasta.gz"), "rt") as f:
        for line in f:
            if line.startswith(">"):
                names.append(line[1:].strip().split(" ")[0])
    with gzip.open(os.path.join(ROOT, "tests", "golden", "example_small.paf.gz"), "rb") as f:
        text = f.read()
    return text, ReadNames(names)


def adversarial_keys(oracle, n, div=1):
    """Keys on which libstdc++'s introsort exhausts its depth limit (McIlroy's adversary played against the oracle's restatement)."""
    k = np.zeros(n, np.uint32)
    oracle.lib.oracle_sort_adversary.restype = C.c_long
    hs = oracle.lib.oracle_sort_adversary(n, k.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert hs >= 1
    return (k // div).astype(np.uint32)


def keys_to_paf(keys, n_names=50):
    """One pile whose resMatches column is `keys` (the other columns identify the line: tStart = line number)."""
    names = [f"n{i}" for i in range(n_names)]
    lines = [f"n0\t9000\t10\t8000\t{'+-'[i & 1]}\tn{1 + i % (n_names - 1)}\t9500\t{i}\t{i + 700}\t{int(k)}\t800\t60" for i, k in enumerate(keys)]
    return ("\n".join(lines) + "\n").encode(), ReadNames(names)


def seeded_corrected(seed, n_reads=300):
    """Strings shaped like alignConsensus output: lower-case raw stretches around upper-case corrected ones."""
    rng = np.random.default_rng(seed)
    reads = []
    for r in range(n_reads):
        kind = r % 10
        n = int(rng.integers(1, 4000))
        s = rng.choice(np.frombuffer(b"acgt", np.uint8), size=n)
        if kind == 0:
            pass                                                                     # nothing corrected
        elif kind == 1:
            s[int(rng.integers(0, n))] -= 32                                         # one corrected base: end == beg -> ""
        elif kind == 2:
            s = s - 32                                                               # all corrected
        elif kind == 3:
            a = int(rng.integers(0, n)); s[a:a + max(1, n // 30)] -= 32              # ~3 % corrected, far below 10 % only if gaps follow
            b = int(rng.integers(0, n)); s[b:b + 1] = (s[b:b + 1] | 32) - 32
        else:
            for _ in range(int(rng.integers(1, 6))):
                a = int(rng.integers(0, n)); b = a + int(rng.integers(1, max(2, n // 3)))
                s[a:b] = (s[a:b] | 32) - 32
        reads.append(s.tobytes())
    reads.append(b"")                                                                # a read without windows
    c = Corrected.__new__(Corrected)
    c.n_reads = len(reads)
    c.read_off = np.zeros(len(reads) + 1, np.uint64)
    c.read_off[1:] = np.cumsum([len(x) for x in reads])
    c.bases = np.frombuffer(b"".join(reads), np.uint8).copy()
    return c


# ------------------------------------------------------------------------------------------------ CPU: the oracle is pinned
def test_oracle_sort_matches_std_sort(oracle, reference):
    rng = np.random.default_rng(0)
    for n in list(range(0, 40)) + [100, 257, 1000, 5000]:
        for rep in range(14):
            hi = [1, 2, 3, 5, 10, 1000, 1 << 30][rep % 7]
            k = rng.integers(0, hi + 1, size=n).astype(np.uint32)
            assert np.array_equal(oracle.sort_desc(k), reference.sort_desc(k)), (n, rep)
    for n in (64, 200, 1000, 4096):
        for k in (np.arange(n), np.arange(n)[::-1], np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]), np.arange(n) % 3, np.arange(n) // 7):
            k = np.ascontiguousarray(k, np.uint32)
            assert np.array_equal(oracle.sort_desc(k), reference.sort_desc(k)), n


def test_oracle_sort_matches_std_sort_in_the_heapsort_branch(oracle, reference):
    oracle.lib.oracle_sort_heap_sorts.restype = C.c_long
    for n, div in ((100, 1), (1000, 1), (1000, 4), (5000, 3), (50000, 7)):
        k = adversarial_keys(oracle, n, div)
        before = oracle.lib.oracle_sort_heap_sorts()
        got = oracle.sort_desc(k)
        if div == 1:
            assert oracle.lib.oracle_sort_heap_sorts() > before                     # the replay went through heapsort
        assert np.array_equal(got, reference.sort_desc(k)), (n, div)


def test_oracle_ingest_matches_reference(oracle, reference):
    lines = 0
    for pk, tk, supports in PAFS:
        text, names = synth_paf(synth_piles(**pk), **tk)
        for ms in supports:
            want = reference.ingest_paf(text, names, ms)
            got = oracle.ingest_paf(text, names, ms)
            assert got.equals(want), f"seed {pk['seed']} maxSupport {ms}: {got.first_mismatch(want)}"
            lines += want.n_lines
    assert lines > 20000
    text, names = example_paf()
    want = reference.ingest_paf(text, names, 150)
    assert want.n_piles == 20 and want.n_lines == text.count(b"\n")
    assert oracle.ingest_paf(text, names, 150).equals(want)
    assert oracle.ingest_paf(text, names, 10).equals(reference.ingest_paf(text, names, 10))


def test_oracle_ingest_pile_boundaries(oracle, reference):
    """An empty line splits a pile; the same query coming back later is a new pile; a name listed twice resolves to its last
    entry; a heapsort-deep pile."""
    text, names = keys_to_paf(np.arange(40) % 5)
    rows = text.split(b"\n")[:-1]
    rows = rows[:10] + [b""] + rows[10:25] + [b"", b""] + [r.replace(b"n0\t", b"n7\t", 1) for r in rows[25:30]] + rows[30:]
    text2 = b"\n" + b"\n".join(rows) + b"\n"
    for ms in (150, 4):
        want = reference.ingest_paf(text2, names, ms)
        assert want.n_piles == 4
        assert oracle.ingest_paf(text2, names, ms).equals(want)
    dup = ReadNames([f"n{i}" for i in range(50)] + ["n3", "n0"])
    want = reference.ingest_paf(text, dup, 150)
    assert int(want.pile_read[0]) == 51 and 50 in want.overlaps[:, 0]
    assert oracle.ingest_paf(text, dup, 150).equals(want)
    text3, names3 = keys_to_paf(adversarial_keys(oracle, 3000, 2))
    assert oracle.ingest_paf(text3, names3, 100).equals(reference.ingest_paf(text3, names3, 100))
    assert oracle.ingest_paf(b"", names, 5).n_piles == 0 and reference.ingest_paf(b"", names, 5).n_piles == 0


def test_oracle_rejects_what_the_reference_cannot_survive(oracle):
    text, names = keys_to_paf(np.arange(5))
    for bad in (text[:-1],                                                           # no final newline: getNextReadPile loops forever
                text.replace(b"\t9000\t", b"\tx9\t", 1),                             # stoi throws
                text.replace(b"\t60\n", b"\n", 1),                                   # 11 columns
                text.replace(b"n0\t", b"nobody\t", 1)):                              # unknown read
        with pytest.raises(RuntimeError):
            oracle.ingest_paf(bad, names, 5)


def test_oracle_finish_matches_reference(oracle, reference):
    n = 0
    for seed in (1, 2):
        cor = seeded_corrected(seed)
        for m in (1, 2, 9, 0):
            want = reference.finish_reads(cor, m)
            got = oracle.finish_reads(cor, m)
            assert got.equals(want), f"seed {seed} merSize {m}: read {got.first_mismatch(want)}"
            n += int((want.read_off[1:] != want.read_off[:-1]).sum())
        assert reference.finish_reads(cor, 0).equals(cor)
    assert n > 1000


# ------------------------------------------------------------------------------------------------ CPU: kernel sources, emulated
@pytest.mark.parametrize("i", range(len(PAFS)))
def test_emulated_ingest_matches_oracle(emu, oracle, i):
    pk, tk, supports = PAFS[i]
    text, names = synth_paf(synth_piles(**pk), **tk)
    cor = emu()
    for ms in supports:
        got = cor.ingest_paf(text, names, ms)
        want = oracle.ingest_paf(text, names, ms)
        assert got.equals(want), f"maxSupport {ms}: {got.first_mismatch(want)}"
    assert cor.ingest_stats()["paf_bytes"] == len(text)


def test_emulated_ingest_edge_cases(emu, oracle):
    cor = emu()
    text, names = example_paf()
    assert cor.ingest_paf(text, names, 150).equals(oracle.ingest_paf(text, names, 150))
    text, names = keys_to_paf(np.arange(40) % 5)
    rows = text.split(b"\n")[:-1]
    rows = rows[:10] + [b""] + rows[10:25] + [b"", b""] + [r.replace(b"n0\t", b"n7\t", 1) for r in rows[25:30]] + rows[30:]
    text2 = b"\n" + b"\n".join(rows) + b"\n"
    assert cor.ingest_paf(text2, names, 4).equals(oracle.ingest_paf(text2, names, 4))
    dup = ReadNames([f"n{i}" for i in range(50)] + ["n3", "n0"])
    assert cor.ingest_paf(text, dup, 150).equals(oracle.ingest_paf(text, dup, 150))
    for n, div in ((3000, 2), (700, 1)):                                            # heapsort branch of the replay
        t3, n3 = keys_to_paf(adversarial_keys(oracle, n, div))
        assert cor.ingest_paf(t3, n3, 100).equals(oracle.ingest_paf(t3, n3, 100))
    assert cor.ingest_paf(b"", names, 5).n_piles == 0
    assert cor.ingest_paf(b"\n\n", names, 5).n_piles == 0


def test_emulated_ingest_rejects_bad_text(emu):
    text, names = keys_to_paf(np.arange(5))
    cor = emu()
    for bad in (text[:-1], text.replace(b"\t9000\t", b"\tx9\t", 1), text.replace(b"\t60\n", b"\n", 1), text.replace(b"n0\t", b"nobody\t", 1),
                text.replace(b"\t10\t", b"\t99999999999\t", 1)):
        with pytest.raises(ConsentError):
            cor.ingest_paf(bad, names, 5)
    assert cor.ingest_paf(text, names, 5).n_piles == 1                               # the handle is still usable


def test_emulated_finish_matches_oracle(emu, oracle):
    """cg_finish_reads = cg_reanchor_reads + trimRead + dropRead, on one handle after the window path."""
    from consent_b200.synth import synth_reads
    batch, reads = synth_reads(10, 8, truth_len=2200, seed=31)
    cor = emu()
    live = cor.correct_windows(batch)
    plain = cor.reanchor_reads(batch, live, reads)
    for m in (1, 3, 0):
        got = cor.finish_reads(batch, live, reads, m)
        want = oracle.finish_reads(plain, m)
        assert got.equals(want), f"merSize {m}: read {got.first_mismatch(want)}"
    assert cor.finish_reads(batch, live, reads, 1).read_off[-1] < plain.read_off[-1]          # something was trimmed


def test_emulated_chain_paf_to_fasta_lines(emu, oracle):
    """PAF text -> piles -> windows -> consensuses -> re-anchored, trimmed reads, all through the ABI, against the oracle chain."""
    p = synth_piles(n_reads=14, genome_len=5000, read_len=1800, seed=9, max_support=4000)
    text, names = synth_paf(p, seed=6, tie_range=2)
    cor = emu()
    ps = cor.ingest_paf(text, names, 8)
    cor.upload_piles(ps.piles(p.store_off, p.store_bases))
    cor.run()
    res = cor.download()
    batch, reads, _ = cor.download_windows()
    got = cor.finish_reads(batch, res, reads, 1)
    ops = oracle.ingest_paf(text, names, 8)
    ob, ord_, _ = oracle.extract_windows(ops.piles(p.store_off, p.store_bases))
    ores, _ = oracle.correct_windows(ob, threads=8)
    want = oracle.finish_reads(oracle.reanchor_reads(ob, ores, ord_, threads=4)[0], 1)
    assert ps.equals(ops) and res.equals(ores)
    assert got.equals(want)
    assert batch.n_windows >= 20
    # the same tail without the host round trips: nothing but the corrected reads leaves the device
    for m in (1, 0, 3):
        assert cor.finish_resident(m).equals(cor.finish_reads(batch, res, reads, m)), m
    cor.upload_piles(ps.piles(p.store_off, p.store_bases))
    with pytest.raises(ConsentError):
        cor.finish_resident(1)                                                       # run() has not been called on the new batch
    cor.run()
    assert cor.finish_resident(1).equals(want)                                       # ... and without download() at all


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_ingest_matches_oracle(gpu, oracle):
    cor = gpu()
    for pk, tk, supports in PAFS + ((dict(n_reads=1500, genome_len=300000, read_len=6000, seed=25, max_support=4000), dict(seed=5, tie_range=25), (150, 30)),):
        text, names = synth_paf(synth_piles(**pk), **tk)
        for ms in supports:
            got = cor.ingest_paf(text, names, ms)
            want = oracle.ingest_paf(text, names, ms)
            assert got.equals(want), f"seed {pk['seed']} maxSupport {ms}: {got.first_mismatch(want)}"
    text, names = example_paf()
    assert cor.ingest_paf(text, names, 150).equals(oracle.ingest_paf(text, names, 150))
    for n, div in ((3000, 2), (700, 1), (40000, 5)):
        t3, n3 = keys_to_paf(adversarial_keys(oracle, n, div))
        assert cor.ingest_paf(t3, n3, 100).equals(oracle.ingest_paf(t3, n3, 100))
    assert cor.ingest_paf(b"", names, 5).n_piles == 0


@pytest.mark.gpu
def test_gpu_ingest_matches_reference_if_present(gpu, reference):
    cor = gpu()
    text, names = synth_paf(synth_piles(n_reads=600, genome_len=150000, read_len=5000, seed=26, max_support=4000), seed=7, tie_range=10)
    assert cor.ingest_paf(text, names, 50).equals(reference.ingest_paf(text, names, 50))


@pytest.mark.gpu
def test_gpu_ingest_rejects_bad_text(gpu):
    text, names = keys_to_paf(np.arange(5))
    cor = gpu()
    for bad in (text[:-1], text.replace(b"\t9000\t", b"\tx9\t", 1), text.replace(b"\t60\n", b"\n", 1), text.replace(b"n0\t", b"nobody\t", 1)):
        with pytest.raises(ConsentError):
            cor.ingest_paf(bad, names, 5)
    assert cor.ingest_paf(text, names, 5).n_piles == 1


@pytest.mark.gpu
def test_gpu_chain_paf_to_fasta_lines(gpu, oracle):
    p = synth_piles(n_reads=500, genome_len=100000, read_len=5000, seed=13, max_support=4000)
    text, names = synth_paf(p, seed=8, tie_range=6)
    cor = gpu()
    ps = cor.ingest_paf(text, names, 40)
    cor.upload_piles(ps.piles(p.store_off, p.store_bases))
    cor.run()
    res = cor.download()
    batch, reads, _ = cor.download_windows()
    plain = cor.reanchor_reads(batch, res, reads)
    ops = oracle.ingest_paf(text, names, 40)
    ob, ord_, _ = oracle.extract_windows(ops.piles(p.store_off, p.store_bases))
    ores, _ = oracle.correct_windows(ob, threads=os.cpu_count() or 4)
    oplain, _ = oracle.reanchor_reads(ob, ores, ord_, threads=os.cpu_count() or 4)
    assert ps.equals(ops) and res.equals(ores) and plain.equals(oplain)
    for m in (1, 4, 0):
        assert cor.finish_reads(batch, res, reads, m).equals(oracle.finish_reads(oplain, m)), m
        assert cor.finish_resident(m).equals(oracle.finish_reads(oplain, m)), m
    assert batch.n_windows > 4000
    cor.upload_piles(ps.piles(p.store_off, p.store_bases))
    cor.run()
    assert cor.finish_resident(1).equals(oracle.finish_reads(oplain, 1))             # no download() / download_windows() in between
