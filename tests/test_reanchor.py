"""Consensus re-anchoring (alignConsensus + the SSW local aligner, reference
src/correctionAlignment.cpp:47-139, BMEAN/Complete-Striped-Smith-Waterman-Library/src/ssw.c).

CPU suite: the oracle (oracle/reanchor_oracle.c) against the unmodified reference and against the committed golden
reads; the kernel sources (consent_b200/csrc/k_reanchor.cuh) on the SIMT emulator against the oracle.
GPU suite (-m gpu): the same checks through libconsent_b200.so on a B200."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from consent_b200._ffi import Params, Reads, Results
from consent_b200.engine import ConsentError
from consent_b200.synth import synth_reads
from tests.reanchor_cases import all_cases, edge_cases, seeded_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def reanchor_golden():
    with open(os.path.join(ROOT, "tests", "golden", "reanchor_golden.json")) as f:
        return {c["name"]: c for c in json.load(f)["cases"]}


def assert_same_reads(got, want, what):
    if not got.equals(want):
        r = got.first_mismatch(want)
        a, b = got.read(r), want.read(r)
        j = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
        raise AssertionError(f"{what}: read {r} differs at base {j} (lengths {len(a)} / {len(b)}):\n got : {a[max(0, j - 30):j + 30]}\n want: {b[max(0, j - 30):j + 30]}")


def check_golden(cor, case, batch, reads):
    dig = hashlib.sha256(batch.bases[:batch.n_bases].tobytes() + reads.read_bases.tobytes()).hexdigest()[:24]
    assert dig == case["input_digest"], "generators no longer reproduce the golden inputs"
    assert [int(cor.read_off[r + 1] - cor.read_off[r]) for r in range(cor.n_reads)] == case["lengths"]
    for r, s in enumerate(case["reads"]):
        assert cor.read(r) == s, f"{case['name']}: read {r} differs from the reference's"
    assert cor.digest()[:24] == case["digest"]


# ------------------------------------------------------------------------------------------------ CPU: the oracle is pinned
def test_oracle_matches_reference(oracle, reference):
    """The restatement against the reference's own alignConsensus (oracle/_ref), incl. the rarely taken branches."""
    n_windows = 0
    for name, batch, reads, params in all_cases():
        res, _ = reference.correct_windows(batch, params, threads=8, with_status=False)
        want, _ = reference.reanchor_reads(batch, res, reads, params, threads=4)
        got, _ = oracle.reanchor_reads(batch, res, reads, params, threads=4)
        assert_same_reads(got, want, f"oracle vs reference, {name}")
        n_windows += batch.n_windows
    batch, reads = synth_reads(120, 5, truth_len=3000, seed=21, thin_every=6, thin_seqs=1)
    res, _ = reference.correct_windows(batch, Params(), threads=os.cpu_count() or 4, with_status=False)
    want, _ = reference.reanchor_reads(batch, res, reads, threads=8)
    got, _ = oracle.reanchor_reads(batch, res, reads, threads=8)
    assert_same_reads(got, want, "oracle vs reference, 120 reads")
    b = (C.c_uint64 * 8)()
    oracle.lib.oracle_reanchor_branches(b)
    assert b[0] > 500 and b[1] > 200 and b[2] > 30 and b[3] > 3, f"arbitration branches not exercised: {list(b)}"
    assert n_windows + batch.n_windows > 1000


def test_oracle_matches_golden(oracle, reanchor_golden):
    """Holds where /root/reference is absent: window consensuses from the window oracle, reads against the golden file."""
    for name, batch, reads, params in all_cases():
        res, _ = oracle.correct_windows(batch, params, threads=8, with_status=False)
        assert res.digest()[:24] == reanchor_golden[name]["windows_digest"], f"{name}: window consensuses differ from the reference's"
        cor, _ = oracle.reanchor_reads(batch, res, reads, params, threads=4)
        check_golden(cor, reanchor_golden[name], batch, reads)


def test_reanchored_reads_have_the_expected_shape(oracle):
    """Domain sanity of the generator + path: well covered reads come back almost entirely upper case and close to
    the truth length; a read without windows comes back empty."""
    batch, reads = synth_reads(4, 12, truth_len=2500, seed=5)
    res, _ = oracle.correct_windows(batch, Params(), threads=8, with_status=False)
    cor, _ = oracle.reanchor_reads(batch, res, reads, threads=4)
    for r in range(cor.n_reads):
        s = cor.read(r)
        assert sum(c.isupper() for c in s) / len(s) > 0.95
        assert abs(len(s) - 2500) < 125                       # raw PB reads are ~ 2610 long, corrected ones ~ 2500
    for name, batch, rd, params in edge_cases():
        res, _ = oracle.correct_windows(batch, params, threads=4, with_status=False)
        cor, _ = oracle.reanchor_reads(batch, res, rd, params)
        assert cor.read(0) == ""


# ------------------------------------------------------------------------------------------------ CPU: kernel sources, emulated
@pytest.mark.parametrize("which", ["pb_n8", "pb_n3_thin2", "ont_n6", "pb_long_windows", "pb_no_overlap", "edges"])
def test_emulated_kernel_matches_oracle_and_golden(emu, oracle, reanchor_golden, which):
    cor = emu()
    for name, batch, reads, params in all_cases():
        if name != which:
            continue
        res, _ = oracle.correct_windows(batch, params, threads=8, with_status=False)
        want, _ = oracle.reanchor_reads(batch, res, reads, params, threads=4)
        got = cor.reanchor_reads(batch, res, reads)
        assert_same_reads(got, want, f"emulated kernel vs oracle, {name}")
        check_golden(got, reanchor_golden[name], batch, reads)
        oracle.lib.oracle_reanchor_cells.restype = C.c_uint64
        assert cor.reanchor_stats()["dp_cells"] == int(oracle.lib.oracle_reanchor_cells())


def _very_long_windows():
    # windows of 1 100 bases: alignments of more than 959 bases leave the halfword (score, row) keys of k_reanchor's scan (KEY16) for
    # the compare path, and need two bands of rows
    return synth_reads(2, 4, truth_len=3600, seed=21, profile="PB", window_size=1100, window_overlap=80)


def test_emulated_kernel_on_alignments_too_long_for_halfword_keys(emu, oracle, reference):
    batch, reads = _very_long_windows()
    res, _ = oracle.correct_windows(batch, Params(), threads=8, with_status=False)
    want, _ = oracle.reanchor_reads(batch, res, reads, Params(), threads=4)
    if reference is not None:
        ref, _ = reference.reanchor_reads(batch, res, reads, Params())
        assert_same_reads(want, ref, "oracle vs reference, 1100-base windows")
    assert_same_reads(emu().reanchor_reads(batch, res, reads), want, "emulated kernel, 1100-base windows")


def test_emulated_resident_results_are_reused(emu, oracle):
    """results of correct_windows on the same handle: the device copies are used, same reads as with host copies."""
    batch, reads = synth_reads(3, 4, truth_len=1700, seed=12, thin_every=4, thin_seqs=1)
    cor = emu()
    live = cor.correct_windows(batch)
    got_live = cor.reanchor_reads(batch, live, reads)
    host = Results(live._r)                                     # copies
    want, _ = oracle.reanchor_reads(batch, host, reads, threads=2)
    assert_same_reads(got_live, want, "resident results")
    other = emu()
    assert_same_reads(other.reanchor_reads(batch, host, reads), want, "uploaded results")


def test_reanchor_rejects_inconsistent_arguments(emu, oracle):
    batch, reads = synth_reads(2, 3, truth_len=1200, seed=3)
    res, _ = oracle.correct_windows(batch, Params(), threads=2, with_status=False)
    bad = Reads(reads.read_win_begin[:-1], reads.read_off[:-1], reads.read_bases, reads.win_pos)   # drops the last read's windows
    with pytest.raises(ConsentError):
        emu().reanchor_reads(batch, res, bad)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_matches_oracle_and_golden(gpu, oracle, reanchor_golden):
    cor = gpu()
    for name, batch, reads, params in all_cases():
        res, _ = oracle.correct_windows(batch, params, threads=8, with_status=False)
        want, _ = oracle.reanchor_reads(batch, res, reads, params, threads=4)
        got = cor.reanchor_reads(batch, res, reads)
        assert_same_reads(got, want, f"GPU vs oracle, {name}")
        check_golden(got, reanchor_golden[name], batch, reads)
        oracle.lib.oracle_reanchor_cells.restype = C.c_uint64
        assert cor.reanchor_stats()["dp_cells"] == int(oracle.lib.oracle_reanchor_cells())


@pytest.mark.gpu
def test_gpu_matches_reference_if_present(gpu, reference):
    cor = gpu()
    batch, reads = synth_reads(40, 6, truth_len=3000, seed=31, thin_every=7, thin_seqs=1)
    res, _ = reference.correct_windows(batch, Params(), threads=os.cpu_count() or 4, with_status=False)
    want, _ = reference.reanchor_reads(batch, res, reads, threads=8)
    assert_same_reads(cor.reanchor_reads(batch, res, reads), want, "GPU vs the unmodified reference")


@pytest.mark.gpu
def test_gpu_alignments_too_long_for_halfword_keys(gpu, oracle):
    batch, reads = _very_long_windows()
    res, _ = oracle.correct_windows(batch, Params(), threads=8, with_status=False)
    want, _ = oracle.reanchor_reads(batch, res, reads, Params(), threads=4)
    assert_same_reads(gpu().reanchor_reads(batch, res, reads), want, "1100-base windows")


@pytest.mark.gpu
def test_gpu_full_path_windows_then_reads(gpu, oracle):
    """correct_windows -> reanchor_reads on one handle (device-resident results), a few thousand windows; the oracle
    re-anchors the GPU's own window results, so this isolates the re-anchoring kernel at scale."""
    cor = gpu()
    batch, reads = synth_reads(300, 12, truth_len=4000, seed=41, thin_every=9, thin_seqs=1)
    live = cor.correct_windows(batch)
    got = cor.reanchor_reads(batch, live, reads)
    host = Results(live._r)
    want, _ = oracle.reanchor_reads(batch, host, reads, threads=os.cpu_count() or 4)
    assert_same_reads(got, want, "GPU resident path vs oracle")
    other = gpu()
    assert_same_reads(other.reanchor_reads(batch, host, reads), want, "GPU upload path vs oracle")
    up = np.frombuffer(got.bases.tobytes(), np.uint8)
    assert ((up >= 65) & (up <= 90)).mean() > 0.9


@pytest.mark.gpu
def test_gpu_reanchor_is_deterministic_and_order_free(gpu, oracle):
    """Reads are independent: re-anchoring any subset gives the same reads (scheduling invariance)."""
    cor = gpu()
    batch, reads = synth_reads(64, 5, truth_len=2600, seed=51)
    res, _ = oracle.correct_windows(batch, Params(), threads=os.cpu_count() or 4, with_status=False)
    a = cor.reanchor_reads(batch, res, reads)
    b = cor.reanchor_reads(batch, res, reads)
    assert a.equals(b)
    # the second half of the reads alone
    r0 = 32
    w0 = int(reads.read_win_begin[r0])
    sub_batch = batch.slice(w0, batch.n_windows)
    sub_reads = Reads(reads.read_win_begin[r0:] - np.uint32(w0), reads.read_off[r0:] - reads.read_off[r0],
                      reads.read_bases[int(reads.read_off[r0]):], reads.win_pos[w0:])
    sub_res, _ = oracle.correct_windows(sub_batch, Params(), threads=os.cpu_count() or 4, with_status=False)
    c = cor.reanchor_reads(sub_batch, sub_res, sub_reads)
    for r in range(sub_reads.n_reads):
        assert c.read(r) == a.read(r0 + r)
