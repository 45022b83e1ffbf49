"""Shared checks (test infrastructure)."""
import hashlib

import numpy as np

from consent_b200._ffi import Batch, Params
from tests.cases import seeded_batches


def solid_digest(res, w):
    a, b = int(res.solid_off[w]), int(res.solid_off[w + 1])
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(res.solid_kmer[a:b]).tobytes())
    h.update(np.ascontiguousarray(res.solid_count[a:b]).tobytes())
    return h.hexdigest()[:24]


def example_piles():
    """The 300 real-data piles of tests/golden/example_windows.txt.gz (pile[0] = template)."""
    import gzip
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example_windows.txt.gz")
    piles, left = [], 0
    with gzip.open(path, "rt") as f:
        for line in f:
            line = line.rstrip("\n")
            if left == 0:
                assert line.startswith("W "), line[:20]
                left = int(line[2:])
                piles.append([])
            else:
                piles[-1].append(line)
                left -= 1
    return piles


def example_batch():
    return Batch.from_piles(example_piles())


def golden_batch(case):
    """(Batch, Params) of a golden case."""
    p = Params(**case["params"])
    if case["kind"] == "explicit":
        return Batch.from_piles(case["piles"]), p
    for name, batch in seeded_batches():
        if name == case["name"]:
            dig = hashlib.sha256(batch.bases[:batch.n_bases].tobytes()).hexdigest()[:24]
            assert dig == case["input_digest"], "synthetic generator is not reproducing the golden inputs"
            return batch, p
    raise KeyError(case["name"])


def assert_matches_golden(res, case):
    n = len(case["consensus"])
    assert res.n_windows == n
    for w in range(n):
        assert res.consensus(w) == case["consensus"][w], f"{case['name']} window {w}: consensus differs"
        assert int(res.status[w]) == case["status"][w], f"{case['name']} window {w}: status differs"
        assert int(res.solid_off[w + 1] - res.solid_off[w]) == case["n_solid"][w], f"{case['name']} window {w}: #solid differs"
        assert solid_digest(res, w) == case["solid_digest"][w], f"{case['name']} window {w}: solid k-mer list differs"


def assert_same(got, want, what=""):
    if not got.equals(want):
        w = got.first_mismatch(want)
        msg = f"{what}: first differing window {w}"
        if w is not None and w < min(got.n_windows, want.n_windows):
            msg += f"\n got : {got.consensus(w)[:300]}\n want: {want.consensus(w)[:300]}\n status {got.status[w]} vs {want.status[w]}; #solid {len(got.solid(w))} vs {len(want.solid(w))}"
        raise AssertionError(msg)
