"""CPU suite, part 3: the product's CUDA kernel sources (consent_b200/csrc/*.cuh), compiled for the SIMT emulator of
tests/emu, against the oracle and the golden vectors.  This is how the kernels are debugged on a box without a GPU;
the `-m gpu` tests repeat the same checks through the real library on a B200."""
import pytest

from consent_b200._ffi import Batch, Params
from consent_b200.engine import ConsentError
from consent_b200.synth import synth_windows
from tests.cases import concat, edge_piles
from tests.helpers import assert_matches_golden, assert_same, golden_batch

FAST_EDGES = {"single_sequence", "two_identical", "template_shorter_than_k", "all_shorter_than_k", "some_reads_shorter_than_k",
              "unrelated_short_no_anchor", "homopolymer", "homopolymer_mixed", "ragged_lengths", "short_window_30",
              "two_haplotypes", "weak_ends", "low_error_deep"}


def test_emulated_kernels_match_golden(emu, golden):
    cor = emu()
    n = 0
    for case in golden["cases"]:
        name = case["name"]
        if name.startswith("edge_") and name[5:] not in FAST_EDGES:
            continue
        if name.startswith("synth_") and any(t in name for t in ("_n20_", "_n60_")):
            continue                                       # long POA rows: covered on the GPU
        batch, params = golden_batch(case)
        c = cor if params == Params() else emu(params)
        assert_matches_golden(c.correct_windows(batch), case)
        n += 1
    assert n >= 20


@pytest.mark.parametrize("n_seqs,n_win,seed", [(1, 8, 31), (2, 16, 32), (3, 16, 33), (8, 8, 34), (47, 3, 36), (150, 3, 37)])
def test_emulated_kernels_match_oracle(emu, oracle, n_seqs, n_win, seed):
    batch = synth_windows(n_win, n_seqs, seed=seed)
    cor = emu()
    got = cor.correct_windows(batch)
    want, _ = oracle.correct_windows(batch, threads=4)
    assert_same(got, want, f"emulated CUDA path vs oracle N={n_seqs}")
    oracle.lib.oracle_reset_counters()
    oracle.correct_windows(batch, threads=1)
    oc, gc = oracle.counters(), cor.counters()
    for key in ("windows", "sequences", "bases", "anchors", "regions", "poa_graphs", "alignments", "dp_cells", "dp_pred_cells",
                "solid_kmers", "consensus_bytes", "fallback_windows"):
        assert gc[key] == oc[key], key


def test_emulated_chunking_and_staged_calls_do_not_change_results(emu, oracle):
    batch = concat([synth_windows(5, 8, seed=41), synth_windows(4, 3, seed=42), synth_windows(2, 47, seed=43)])
    want, _ = oracle.correct_windows(batch, threads=4)
    one = emu()
    assert_same(one.correct_windows(batch), want, "one chunk")
    many = emu(chunk_max_windows=3)
    many.upload(batch)
    many.run()
    assert_same(many.download(), want, "chunks of 3 windows, staged calls")
    many.run()                                              # re-running a resident batch is idempotent
    assert_same(many.download(), want, "second run")


def test_emulated_poa_tier_overflow_requeues_jobs(emu, oracle):
    # N = 8 piles have few anchors: their regions run from a few bases to whole windows, so the jobs spread over the
    # compact and the wide tiers (and some are re-queued from one to the next)
    batch = synth_windows(3, 8, seed=44)
    want, _ = oracle.correct_windows(batch, threads=4)
    assert_same(emu().correct_windows(batch), want, "jobs spread over / re-queued through the POA tiers")
    # two unrelated sequences (1500 and 2100 bases): no anchors -> one whole-window region, too long for every k_poa2 tier -> k_poa
    long_pile = Batch.from_piles([["".join("ACGT"[(i * 7 + i // 3) % 4] for i in range(1500)),
                                   "".join("ACGT"[(i * 5 + i // 7 + 1) % 4] for i in range(2100))]])
    want, _ = oracle.correct_windows(long_pile, threads=1)
    assert_same(emu().correct_windows(long_pile), want, "last-resort tier 1")
    assert_same(emu(poa_tier1_cells=1 << 20).correct_windows(long_pile), want, "last-resort tier 1 -> 2")
    # no tier can hold it: the window comes back as its raw template with status CG_WINDOW_ERROR, the call succeeds
    got = emu(poa_tier1_cells=1 << 20, poa_tier2_cells=1 << 20).correct_windows(long_pile)
    assert int(got.status[0]) == 2 and got.consensus(0) == bytes(long_pile.bases[:int(long_pile.seq_off[1])]).decode()      # (its solid k-mer list was already counted and is returned)


def test_emulated_deep_piles(emu, oracle):
    # > 256 sequences (k_split's general path) and > 81 920 k-mer occurrences (k_index's direct-count path)
    batch = concat([synth_windows(1, 300, seed=81), synth_windows(1, 40, seed=82)])
    want, _ = oracle.correct_windows(batch, threads=4)
    assert_same(emu().correct_windows(batch), want, "300-deep pile")


def test_emulated_kmer_counts_around_the_byte_counter_width(emu, oracle):
    from tests.cases import counter_width_piles
    for name, pile in counter_width_piles():
        batch = Batch.from_piles([pile])
        want, _ = oracle.correct_windows(batch, threads=2)
        assert max(c for _, c in want.solid(0)) >= 255, name
        assert_same(emu().correct_windows(batch), want, name)


def test_emulated_errors(emu):
    cor = emu()
    with pytest.raises(ConsentError) as e:
        cor.correct_windows(Batch.from_piles([["ACGTNACGTACGTAGCTAGCTAGCATCGATCGATCGA", "ACGTACGTACGTAGCTAGCTAGC"]]))
    assert e.value.code == -5                               # CG_ERR_BAD_BASE
    with pytest.raises(ConsentError) as e:
        cor.run() if False else emu().run()
    assert e.value.code == -7                               # CG_ERR_STATE: run before upload
    got = cor.correct_windows(Batch.from_piles([["A" * 7000, "ACGT"]]))            # over a stated limit (6000 bases): not corrected, not fatal
    assert int(got.status[0]) == 2 and got.consensus(0) == "A" * 7000              # CG_WINDOW_ERROR: the raw template


def test_emulated_lifetime_and_empty_batch(emu):
    import numpy as np
    cor = emu()
    empty = Batch(np.zeros(1, np.uint32), np.zeros(1, np.uint64), np.zeros(1, np.uint8))
    r = cor.correct_windows(empty)
    assert r.n_windows == 0 and list(r.cons_off) == [0] and list(r.solid_off) == [0]
    b = synth_windows(3, 5, seed=1)
    r1 = cor.correct_windows(b)
    r2 = cor.correct_windows(b)                      # two result sets alive at once (the pool grows), same content
    assert r1.equals(r2)
    text = r1.consensus(0)
    cor.close()                                      # results outlive their corrector: freed when they die
    assert r1.consensus(0) == text and r2.consensus(2) == r1.consensus(2)
    own = r2.detach()
    assert own.consensus(0) == text


def test_emulated_kernels_on_real_piles(emu, example_golden):
    from tests.helpers import assert_matches_golden, example_piles
    piles = example_piles()[:60]
    res = emu().correct_windows(Batch.from_piles(piles))
    case = {k: (v[:60] if isinstance(v, list) else v) for k, v in example_golden.items()}
    assert_matches_golden(res, case)


def test_emulated_window_over_a_limit_comes_back_as_its_template_and_the_batch_goes_on(emu, oracle):
    """CG_WINDOW_ERROR: a window over a limit of this build (here: 4200 sequences > 4095, and a 2100-base template > 2047 k-mers) is
    returned as its raw template with status 2; its neighbours are corrected as if it were not there (the reference never fails a
    window: src/correctionMSA.cpp:29-49)."""
    import numpy as np
    good1, good2 = synth_windows(2, 8, seed=51), synth_windows(2, 20, seed=52)
    rng = np.random.default_rng(3)
    deep = Batch.from_piles([["".join("ACGT"[i] for i in rng.integers(0, 4, 40))] + ["ACGTACGTACGTAAC"] * 4199])
    long_tpl = Batch.from_piles([["".join("ACGT"[i] for i in rng.integers(0, 4, 2100)), "ACGTACGTACGTAAC", "ACGTTTGACGTACGTAAC"]])
    batch = concat([good1, deep, good2, long_tpl])
    cor = emu()
    got = cor.correct_windows(batch)
    want, _ = oracle.correct_windows(concat([good1, good2]), threads=4)
    idx_good = [0, 1, 3, 4]
    for i, w in enumerate(idx_good):
        assert got.consensus(w) == want.consensus(i) and int(got.status[w]) == int(want.status[i])
        assert got.solid(w) == want.solid(i)
    for w, b in ((2, deep), (5, long_tpl)):
        assert int(got.status[w]) == 2
        assert got.consensus(w) == bytes(b.bases[:int(b.seq_off[1])]).decode()
        assert got.solid(w) == []
    assert cor.counters()["error_windows"] == 2


def test_emulated_2bit_input_and_resident_solid_lists_do_not_change_results(emu, oracle):
    """"input_2bit": a quarter of the bytes over the bus; "results_with_solid" 0: the solid lists stay in HBM, where the
    re-anchoring reads them.  Neither changes a byte of what is computed."""
    from consent_b200.synth import synth_reads
    batch, reads = synth_reads(5, 8, truth_len=2300, seed=77)
    want, _ = oracle.correct_windows(batch, threads=4)
    want_reads, _ = oracle.reanchor_reads(batch, want, reads, threads=4)
    cor = emu(chunk_max_windows=7)                          # several chunks: their base ranges do not start on 4-base boundaries
    cor.set_option("input_2bit", 1)
    got = cor.correct_windows(cor.pack_2bit(batch))
    assert_same(got, want, "2-bit packed input")
    cor.set_option("results_with_solid", 0)
    live = cor.correct_windows(cor.pack_2bit(batch))
    assert all(live.consensus(w) == want.consensus(w) for w in range(want.n_windows))
    assert int(live.solid_off[-1]) == 0
    assert cor.reanchor_reads(batch, live, reads).equals(want_reads)


@pytest.mark.parametrize("k", [11, 15])
def test_emulated_long_kmers_hashed_index(emu, oracle, k):
    """k = 10 .. 15: the hash-table path of k_index (the direct table holds 4^k <= 2^18 keys)."""
    p = Params(mer_size=k)
    batch = concat([synth_windows(3, 20, seed=91), synth_windows(1, 150, seed=92), synth_windows(3, 12, seed=93, profile="ONT")])
    want, _ = oracle.correct_windows(batch, p, threads=4)
    assert_same(emu(p).correct_windows(batch), want, f"k = {k}")


def _stage_lines(dump: str):
    return [ln for ln in dump.split("\n") if ln[:2] in ("S ", "M ", "T ", "A ", "R ", "G ", "g ", "c ", "C ")]


def test_emulated_kernels_match_the_oracle_stage_by_stage(emu, oracle):
    """Not only the end result: solid list, surviving template k-mers, anchor chain, mean distances, every region's segments and
    consensus, and the stitched consensus of the CUDA path equal the oracle's dump line for line (cg_debug_dump_window)."""
    from tests.cases import edge_piles
    batch = concat([synth_windows(3, 20, seed=71), synth_windows(1, 150, seed=72), synth_windows(3, 8, seed=73),
                    Batch.from_piles([p for n, p in edge_piles(3) if n in ("two_haplotypes", "unrelated_short_no_anchor", "weak_ends")])])
    for params in (Params(), Params(min_anchors=50)):
        cor = emu(params)
        cor.upload(batch)
        cor.run()
        for w in range(batch.n_windows):
            assert _stage_lines(cor.dump_window(w)) == _stage_lines(oracle.dump_window(batch, w, params)), f"window {w}, {params}"


def test_integer_comparable_tests_equal_the_fp64_ones(entry):
    """cg_common.cuh evaluates comparable(x, mean) (bmean.cpp:286-295) and comparable(x, deciles) (:264-282) in integers on the device;
    the product's own functions, compiled for the host, agree with the fp64 forms on every argument of a dense grid and around powers of two."""
    import ctypes as C
    lib = C.CDLL(entry.build_emu())
    lib.emu_comparable_selfcheck.restype = C.c_uint64
    lib.emu_comparable_selfcheck.argtypes = [C.c_uint32, C.c_uint32]
    assert lib.emu_comparable_selfcheck(1500, 1500) == 0
