"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py
The fixtures pin the oracle (and through it the CUDA path) where /root/reference is absent (the GPU box)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from consent_b200._ffi import Batch, Params  # noqa: E402
from tests.cases import edge_piles, seeded_batches  # noqa: E402
from tests.refs import Reference  # noqa: E402


def solid_digest(res, w):
    a, b = int(res.solid_off[w]), int(res.solid_off[w + 1])
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(res.solid_kmer[a:b]).tobytes())
    h.update(np.ascontiguousarray(res.solid_count[a:b]).tobytes())
    return h.hexdigest()[:24]


def record(ref, batch, params):
    res, _ = ref.correct_windows(batch, params, threads=os.cpu_count() or 1)
    return {
        "consensus": [res.consensus(w) for w in range(batch.n_windows)],
        "status": [int(x) for x in res.status],
        "n_solid": [int(res.solid_off[w + 1] - res.solid_off[w]) for w in range(batch.n_windows)],
        "solid_digest": [solid_digest(res, w) for w in range(batch.n_windows)],
    }


def main():
    ref = Reference()
    out = {"reference": "morispi/CONSENT v2.2.2 fad9e342 (BMEAN 40ab186c, spoa 4.0.0 0ed1abf3)", "cases": []}
    default = Params()
    for name, batch in seeded_batches():
        c = {"name": name, "kind": "synth", "params": default.__dict__,
             "input_digest": hashlib.sha256(batch.bases[:batch.n_bases].tobytes()).hexdigest()[:24]}
        c.update(record(ref, batch, default))
        out["cases"].append(c)
    for name, pile in edge_piles():
        batch = Batch.from_piles([pile])
        c = {"name": "edge_" + name, "kind": "explicit", "params": default.__dict__, "piles": [pile]}
        c.update(record(ref, batch, default))
        out["cases"].append(c)
    # parameter variants on one explicit batch (first 4 windows of the N=8 synthetic case)
    from consent_b200.synth import synth_windows
    b8 = synth_windows(4, 8, seed=9)
    piles = [b8.pile(w) for w in range(4)]
    for pname, p in (("k7", Params(mer_size=7)), ("k5_solid2", Params(mer_size=5, solid_thresh=2)),
                     ("solid8", Params(solid_thresh=8)), ("common3", Params(common_kmers=3)),
                     ("min_anchors_50", Params(min_anchors=50)), ("polish_wrapper", Params(min_anchors=10))):
        c = {"name": "params_" + pname, "kind": "explicit", "params": p.__dict__, "piles": piles}
        c.update(record(ref, Batch.from_piles(piles), p))
        out["cases"].append(c)
    # per-stage dumps of a few windows (pins the oracle stage by stage)
    dumps = []
    for name, batch in seeded_batches():
        if name.split("_")[2] in ("n3", "n8", "n47"):
            dumps.append({"case": name, "window": 0, "text": ref.dump_window(batch, 0, default)})
    out["dumps"] = dumps
    path = os.path.join(ROOT, "tests", "golden", "golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["cases"]), "cases,", len(dumps), "dumps;",
          "fallback windows:", sum(sum(c["status"]) for c in out["cases"]))


if __name__ == "__main__":
    main()
