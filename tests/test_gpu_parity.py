"""GPU suite (-m gpu): the CUDA path, called through the C ABI (libconsent_b200.so), against the oracle on the same
seeded inputs, against the committed golden vectors (reference outputs), and — at sizes the oracle cannot cover in
seconds — through size-independent properties.  Bit-exact everywhere: the path is integer/byte work."""
import numpy as np
import pytest

from consent_b200._ffi import Batch, Params
from consent_b200.engine import ConsentError
from consent_b200.synth import synth_windows
from tests.cases import concat, edge_batch, edge_piles
from tests.helpers import assert_matches_golden, assert_same, golden_batch

pytestmark = pytest.mark.gpu


def test_gpu_matches_golden(gpu, golden):
    cor = gpu()
    for case in golden["cases"]:
        batch, params = golden_batch(case)
        c = cor if params == Params() else gpu(params)
        assert_matches_golden(c.correct_windows(batch), case)


@pytest.mark.parametrize("n_seqs,n_win,seed,profile", [
    (1, 64, 51, "PB"), (2, 128, 52, "PB"), (3, 128, 53, "PB"), (5, 128, 54, "PB"), (8, 256, 55, "PB"),
    (20, 400, 56, "PB"), (47, 96, 57, "PB"), (150, 64, 58, "PB"), (20, 128, 59, "ONT"), (60, 48, 60, "ONT")])
def test_gpu_matches_oracle(gpu, oracle, n_seqs, n_win, seed, profile):
    batch = synth_windows(n_win, n_seqs, seed=seed, profile=profile)
    cor = gpu()
    got = cor.correct_windows(batch)
    want, _ = oracle.correct_windows(batch, threads=32)
    assert_same(got, want, f"CUDA path vs oracle N={n_seqs} {profile}")
    oracle.lib.oracle_reset_counters()
    oracle.correct_windows(batch, threads=32)
    oc, gc = oracle.counters(), cor.counters()
    for key in oc:
        assert gc[key] == oc[key], key                      # the inputs of the algorithmic-bytes model agree too


def test_gpu_matches_oracle_on_edge_cases(gpu, oracle):
    for seed in (1, 2, 3):
        batch = edge_batch(seed)
        want, _ = oracle.correct_windows(batch, threads=32)
        assert_same(gpu().correct_windows(batch), want, f"edge cases seed {seed}")


def test_gpu_parameter_variants(gpu, oracle):
    batch = synth_windows(48, 12, seed=61)
    for p in (Params(mer_size=7), Params(mer_size=5, solid_thresh=2), Params(mer_size=4), Params(solid_thresh=8),
              Params(common_kmers=3), Params(min_anchors=50), Params(min_anchors=10), Params(mer_size=2, solid_thresh=1)):
        want, _ = oracle.correct_windows(batch, p, threads=32)
        assert_same(gpu(p).correct_windows(batch), want, str(p))


@pytest.mark.parametrize("k", [10, 11, 13, 15])
def test_gpu_long_kmers_hashed_index(gpu, oracle, k):
    """merSize 10 .. 15 (--merSize is a user flag, CONSENT-correct:60-151; the reference allows k <= 15, BMEAN/bmean.cpp:46): k-mers
    counted in the hash table of k_index instead of the direct table; every stage downstream (anchors, polish, re-anchoring) sees
    longer k-mers."""
    from consent_b200.synth import synth_reads
    p = Params(mer_size=k)
    cor = gpu(p)
    batch = concat([synth_windows(24, 20, seed=91), synth_windows(6, 150, seed=92), synth_windows(16, 12, seed=93, profile="ONT"),
                    synth_windows(8, 60, seed=94, profile="ONT"), edge_batch(4)])
    want, _ = oracle.correct_windows(batch, p, threads=32)
    assert_same(cor.correct_windows(batch), want, f"k = {k}")
    rb, reads = synth_reads(12, 15, truth_len=3000, seed=95, profile="ONT")
    live = cor.correct_windows(rb)
    wres, _ = oracle.correct_windows(rb, p, threads=32)
    wreads, _ = oracle.reanchor_reads(rb, wres, reads, p, threads=16)
    assert cor.reanchor_reads(rb, live, reads).equals(wreads)


def test_gpu_mixed_depth_batch_and_chunking(gpu, oracle):
    batch = concat([synth_windows(40, 8, seed=62), synth_windows(30, 3, seed=63), synth_windows(6, 150, seed=64),
                    synth_windows(20, 20, seed=65), edge_batch(4)])
    want, _ = oracle.correct_windows(batch, threads=32)
    assert_same(gpu().correct_windows(batch), want, "one chunk")
    small = gpu(chunk_max_windows=7)
    small.upload(batch)
    small.run()
    assert_same(small.download(), want, "chunks of 7 windows, staged calls")
    small.run()
    assert_same(small.download(), want, "re-run of a resident batch")


def test_gpu_poa_tier_overflow(gpu, oracle):
    batch = concat([synth_windows(16, 8, seed=66), Batch.from_piles([p for n, p in edge_piles(5) if "no_anchor" in n or "outlier" in n])])
    want, _ = oracle.correct_windows(batch, threads=32)
    assert_same(gpu().correct_windows(batch), want, "jobs spread over / re-queued through the POA tiers")
    # two unrelated sequences (1500 and 2100 bases): one whole-window region, too long for every k_poa2 tier -> k_poa (last resort)
    long_pile = Batch.from_piles([["".join("ACGT"[(i * 7 + i // 3) % 4] for i in range(1500)),
                                   "".join("ACGT"[(i * 5 + i // 7 + 1) % 4] for i in range(2100))]])
    want, _ = oracle.correct_windows(long_pile, threads=1)
    assert_same(gpu().correct_windows(long_pile), want, "last-resort tier 1")
    assert_same(gpu(poa_tier1_cells=1 << 20).correct_windows(long_pile), want, "last-resort tier 1 -> 2")
    # no tier can hold it: the window comes back as its raw template with status CG_WINDOW_ERROR, the call succeeds
    got = gpu(poa_tier1_cells=1 << 20, poa_tier2_cells=1 << 20).correct_windows(long_pile)
    assert int(got.status[0]) == 2 and got.consensus(0) == bytes(long_pile.bases[:int(long_pile.seq_off[1])]).decode()      # (its solid k-mer list was already counted and is returned)


def test_gpu_errors(gpu):
    cor = gpu()
    with pytest.raises(ConsentError) as e:
        cor.correct_windows(Batch.from_piles([["ACGTNACGTACGTAGCTAGCTAGCATCGATCGATCGA", "ACGTACGTACGTAGCTAGCTAGC"]]))
    assert e.value.code == -5
    with pytest.raises(ConsentError) as e:
        gpu().run()
    assert e.value.code == -7
    got = cor.correct_windows(Batch.from_piles([["A" * 7000, "ACGT"]]))            # over a stated limit (6000 bases): not corrected, not fatal
    assert int(got.status[0]) == 2 and got.consensus(0) == "A" * 7000              # CG_WINDOW_ERROR: the raw template
    # the context stays usable after an error
    b = synth_windows(4, 5, seed=67)
    assert cor.correct_windows(b).n_windows == 4


# ---- BASELINE-size shapes: properties that do not need the oracle on every window --------------------------------

def test_gpu_config2_shape_properties(gpu, oracle):
    """10k x 20 (config 2) is too slow for the oracle in a test; check 2 000 windows by properties + a 200-window
    oracle sample + invariance of the digest under re-chunking and under a permutation of the windows."""
    W = 2000
    batch = synth_windows(W, 20, seed=42)
    cor = gpu()
    res = cor.correct_windows(batch)
    assert res.n_windows == W
    sample = batch.slice(900, 1100)
    want, _ = oracle.correct_windows(sample, threads=32)
    got = gpu().correct_windows(sample)
    assert_same(got, want, "windows 900..1100 of the config-2 stream")
    for w in range(0, 200, 17):
        assert res.consensus(900 + w) == want.consensus(w)                     # a window's result does not depend on its batch
        assert res.solid(900 + w) == want.solid(w)
    assert gpu(chunk_max_windows=333).correct_windows(batch).digest() == res.digest()
    # domain properties: solid lists sorted, counts >= solid; consensus alphabet; status; length near the truth (500)
    for w in range(0, W, 97):
        s = res.solid(w)
        ks = [k for k, _ in s]
        assert ks == sorted(ks) and len(set(ks)) == len(ks) and all(c >= 4 for _, c in s)
        cons = res.consensus(w)
        assert set(cons) <= set("ACGTacgt") and 400 < len(cons) < 650
    assert int(res.status.sum()) == 0
    # permutation: reversed window order -> reversed results
    rev = Batch.from_piles([batch.pile(w) for w in range(W - 1, W - 201, -1)])
    rres = gpu().correct_windows(rev)
    for i in range(200):
        assert rres.consensus(i) == res.consensus(W - 1 - i)


def test_gpu_config3_shape_properties(gpu, oracle):
    """100k x 150 (config 3): 1 000 windows here; 24 of them against the oracle."""
    W = 1000
    batch = synth_windows(W, 150, seed=42)
    res = gpu().correct_windows(batch)
    sample = batch.slice(500, 524)
    want, _ = oracle.correct_windows(sample, threads=32)
    for w in range(24):
        assert res.consensus(500 + w) == want.consensus(w)
        assert res.solid(500 + w) == want.solid(w)
        assert res.status[500 + w] == want.status[w]
    assert gpu(chunk_max_windows=128).correct_windows(batch).digest() == res.digest()
    assert gpu(chunk_max_windows=100, lanes=1).correct_windows(batch).digest() == res.digest()     # one lane, many chunks
    # self-consistency: correcting a pile whose reads are all the consensus returns the consensus, fully solid
    cons = [res.consensus(w).upper() for w in range(0, 40)]
    again = gpu().correct_windows(Batch.from_piles([[c] * 6 for c in cons]))
    for i, c in enumerate(cons):
        assert again.consensus(i) == c


def test_gpu_config3_stream_is_invariant_under_scheduling(gpu, oracle, reference):
    """A longer slice of the config-3 stream (8 192 windows x 150): the result must not depend on how the batch is cut
    into chunks, on the number of lanes in flight, or on staged vs pipelined calls; a 96-window sample is checked against
    the unmodified reference itself."""
    W = 8192
    batch = synth_windows(W, 150, seed=42, first_window=50000)
    a = gpu().correct_windows(batch)
    assert a.n_windows == W and int(a.status.sum()) == 0
    d = a.digest()
    assert gpu(chunk_max_windows=1500).correct_windows(batch).digest() == d
    assert gpu(lanes=1).correct_windows(batch).digest() == d
    staged = gpu(chunk_max_windows=3000)
    staged.upload(batch)
    staged.run()
    assert staged.download().digest() == d
    staged.run()
    assert staged.download().digest() == d                                  # idempotent on a resident batch
    sample = batch.slice(4000, 4096)
    want, _ = reference.correct_windows(sample, threads=32)
    for w in range(96):
        assert a.consensus(4000 + w) == want.consensus(w)
        assert a.solid(4000 + w) == want.solid(w)
    # work counters of the run equal the oracle's on the same windows (the inputs of the algorithmic-byte model)
    oracle.lib.oracle_reset_counters()
    oracle.correct_windows(sample, threads=1)
    oc = oracle.counters()
    c2 = gpu()
    c2.correct_windows(sample)
    gc = c2.counters()
    for key in ("alignments", "dp_cells", "dp_pred_cells", "poa_graphs", "anchors", "solid_kmers", "consensus_bytes"):
        assert gc[key] == oc[key], key


def _stream_against_golden(gpu, name, step):
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stream_digests.json")))[name]
    W, N = gold["windows"], gold["seqs_per_window"]
    cor, run = gpu(), None
    for w0 in range(0, W, step):
        res = cor.correct_windows(synth_windows(min(step, W - w0), N, seed=gold["seed"], profile=gold["profile"], first_window=w0))
        assert int((res.status == 2).sum()) == 0
        run = res.stream_digests(run)
    cons, solid = res.stream_digest_pair(run)
    assert cons == gold["consensus_sha256"], f"{name}: consensus stream differs from the reference's"
    assert solid == gold["solid_sha256"], f"{name}: solid k-mer stream differs from the reference's"


def test_gpu_config2_whole_stream_equals_the_reference(gpu):
    """BASELINE config 2, all 10 000 windows x 20 sequences: every consensus byte, status and solid (k-mer, count) pair against the digests
    of the unmodified reference's output (tests/golden/stream_digests.json, made by tests/golden/make_stream_digests.py)."""
    _stream_against_golden(gpu, "config2", 4000)


def test_gpu_config3_whole_stream_equals_the_reference(gpu):
    """BASELINE config 3, all 100 000 windows x 150 sequences (the headline workload), fed in slices of 12 500 windows."""
    _stream_against_golden(gpu, "config3", 12500)


def test_gpu_polishing_depth_piles(gpu, oracle):
    """CONSENT-polish piles are far deeper than the corrector's 150 (maxSupport 20000, CONSENT-polish:42-43): more than
    256 sequences (k_split's general path), more than 81 920 k-mer occurrences (k_index's direct-count path, pile too big
    for shared memory), hundreds of segments per region (wide POA tiers)."""
    batch = concat([synth_windows(3, 400, seed=91), synth_windows(2, 300, seed=92, profile="ONT"),
                    Batch.from_piles([synth_windows(1, 150, seed=93).pile(0)[:1] + [s[:200] for s in synth_windows(8, 150, seed=94).pile(0)] * 6])])
    want, _ = oracle.correct_windows(batch, threads=32)
    assert_same(gpu().correct_windows(batch), want, "polishing-depth piles")
    deep = synth_windows(1, 1200, seed=95)
    want, _ = oracle.correct_windows(deep, threads=4)
    assert_same(gpu().correct_windows(deep), want, "1200-deep pile")


def test_gpu_kmer_counts_around_the_byte_counter_width(gpu, oracle):
    """k_index counts in bytes and recounts a window in 32 bits when a k-mer passes 255 occurrences: 255 / 256 / 257 and both key ranges."""
    from tests.cases import counter_width_piles
    cases = counter_width_piles()
    batch = Batch.from_piles([p for _, p in cases])
    want, _ = oracle.correct_windows(batch, threads=4)
    for w, (name, _) in enumerate(cases):
        assert max(c for _, c in want.solid(w)) >= 255, name
    assert_same(gpu().correct_windows(batch), want, "k-mer counts of 255 / 256 / 257")


def test_gpu_position_table_too_big_for_the_pair_bits(gpu, oracle):
    """k_index tells a k-mer held twice by one read from one bit per (read, candidate) in shared memory while reads x candidates fit
    458 752 bits; a deep low-error pile (1 500 reads, ~340 candidates) exceeds that and takes the sweep over the position table."""
    import random
    from tests.cases import _mutate, _rand_seq
    rng = random.Random(17)
    truth = _rand_seq(rng, 500)
    pile = [_mutate(rng, truth, 0.04) for _ in range(1500)]
    batch = Batch.from_piles([pile, pile[:40]])
    want, _ = oracle.correct_windows(batch, threads=8)
    cor = gpu()
    got = cor.correct_windows(batch)
    assert_same(got, want, "1500-deep low-error pile")
    one = gpu()
    one.correct_windows(Batch.from_piles([pile]))
    assert one.counters()["anchors"] * 1500 > 458752          # the deep window's chain alone is longer than the bit table allows candidates for


def test_gpu_two_handles_in_two_threads(gpu, oracle):
    """The ABI is re-entrant per handle (the reference call is made from --nproc threads, src/CONSENT-correction.cpp:76-111)."""
    import threading
    batches = [synth_windows(300, 20, seed=97), synth_windows(120, 150, seed=98)]
    wants = [oracle.correct_windows(b, threads=16)[0] for b in batches]
    got = [None, None]

    def work(i):
        cor = gpu(chunk_max_windows=64)
        for _ in range(2):
            got[i] = cor.correct_windows(batches[i]).detach()
        cor.close()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i in range(2):
        assert_same(got[i], wants[i], f"handle {i} run beside another handle")


def test_gpu_matches_reference_on_real_piles(gpu, example_golden):
    """BASELINE config 1 shape: 300 windows of the shipped example/reads.fasta (piles cut by the reference's own
    minimap2 + alignmentPiles/alignmentWindows code, outputs of the unmodified reference committed as a fixture)."""
    from tests.helpers import assert_matches_golden, example_batch
    batch = example_batch()
    assert_matches_golden(gpu().correct_windows(batch), example_golden)
    assert_matches_golden(gpu(chunk_max_windows=37, lanes=2).correct_windows(batch), example_golden)


def _stage_lines(dump: str):
    return [ln for ln in dump.split("\n") if ln[:2] in ("S ", "M ", "T ", "A ", "R ", "G ", "g ", "c ", "C ")]


def test_gpu_matches_the_oracle_stage_by_stage(gpu, oracle):
    """Every stage, not only the end result (a bug that cancels out downstream would pass an end-to-end comparison): solid list,
    surviving template k-mers, anchor chain, mean distances, each region's segments and consensus, stitched consensus — the CUDA
    path's dump (cg_debug_dump_window) against the oracle's, line for line, on synthetic windows at 8 / 20 / 47 / 150 sequences,
    the degenerate piles and 60 real windows of the shipped example.  (The oracle's dump is pinned line for line against the
    unmodified reference's in tests/test_oracle.py.)"""
    from tests.helpers import example_batch
    ex = example_batch()
    batch = concat([synth_windows(12, 20, seed=71), synth_windows(4, 150, seed=72), synth_windows(12, 8, seed=73), synth_windows(6, 47, seed=74),
                    synth_windows(8, 12, seed=75, profile="ONT"), edge_batch(3), ex.slice(0, 60)])
    for params in (Params(), Params(min_anchors=50), Params(mer_size=11)):
        cor = gpu(params)
        cor.upload(batch)
        cor.run()
        for w in range(batch.n_windows):
            assert _stage_lines(cor.dump_window(w)) == _stage_lines(oracle.dump_window(batch, w, params)), f"window {w}, {params}"
