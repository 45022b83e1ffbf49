"""Generates tests/golden/reanchor_golden.json from the UNMODIFIED reference (oracle/_ref: the reference's own
alignConsensus + vendored SSW library, oracle/Makefile).  Run in the build container only:
    python tests/golden/make_reanchor_golden.py
Per case: the window consensuses come from the reference's computeConsensusReadCorrection, the corrected reads from
its alignConsensus; stored = every corrected read of the small cases, digests of the others."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.reanchor_cases import all_cases  # noqa: E402
from tests.refs import Reference  # noqa: E402


def main():
    ref = Reference()
    out = {"reference": "morispi/CONSENT v2.2.2 fad9e342 (alignConsensus, SSW library 1.2.3)", "cases": []}
    for name, batch, reads, params in all_cases():
        res, _ = ref.correct_windows(batch, params, threads=os.cpu_count() or 1, with_status=False)
        cor, _ = ref.reanchor_reads(batch, res, reads, params, threads=4)
        out["cases"].append({
            "name": name,
            "input_digest": hashlib.sha256(batch.bases[:batch.n_bases].tobytes() + reads.read_bases.tobytes()).hexdigest()[:24],
            "windows_digest": res.digest()[:24],
            "digest": cor.digest()[:24],
            "lengths": [int(cor.read_off[r + 1] - cor.read_off[r]) for r in range(cor.n_reads)],
            "reads": [cor.read(r) for r in range(cor.n_reads)] if batch.n_windows <= 24 else [cor.read(0)],
        })
    path = os.path.join(ROOT, "tests", "golden", "reanchor_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
