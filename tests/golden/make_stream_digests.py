#!/usr/bin/env python
"""tests/golden/stream_digests.json: digests of what the UNMODIFIED reference (oracle/_ref) returns for the full synthetic streams of
BASELINE configs 2 and 3 (seed 42, PB): 10 000 windows x 20 sequences and 100 000 windows x 150 sequences.  bench.py and the GPU
tests compare the CUDA path's whole output with them (no spot checks).  ~6 min on 8 cores.  Digest = Results.stream_digests()."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from consent_b200.synth import synth_windows  # noqa: E402
from tests.refs import Reference  # noqa: E402

ref = Reference()
threads = os.cpu_count() or 8
out = {"note": "sha256 stream digests (Results.stream_digests) of the unmodified reference's output; generator tests/golden/make_stream_digests.py"}
for name, W, N, step in (("config2", 10000, 20, 2500), ("config3", 100000, 150, 2000)):
    run = None
    for w0 in range(0, W, step):
        batch = synth_windows(min(step, W - w0), N, seed=42, profile="PB", first_window=w0)
        res, _ = ref.correct_windows(batch, threads=threads)
        run = res.stream_digests(run)
        print(name, w0 + batch.n_windows, flush=True)
    cons, solid = res.stream_digest_pair(run)
    out[name] = {"windows": W, "seqs_per_window": N, "seed": 42, "profile": "PB", "consensus_sha256": cons, "solid_sha256": solid}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "stream_digests.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
