"""Reference outputs for the real-data piles of tests/golden/example_windows.txt.gz (300 windows of the shipped
example/reads.fasta, cut by the reference's own code: make_example_windows.sh).  Build container only:
    bash tests/golden/make_example_windows.sh && python tests/golden/make_example_golden.py
Writes tests/golden/example_golden.json: per window the consensus, status, #solid and a digest of the solid list, as
returned by the UNMODIFIED reference (oracle/_ref)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from consent_b200._ffi import Params  # noqa: E402
from tests.golden.make_golden import record  # noqa: E402
from tests.helpers import example_batch  # noqa: E402
from tests.refs import Reference  # noqa: E402


def main():
    batch = example_batch()
    out = {"reference": "morispi/CONSENT v2.2.2 fad9e342; piles: example/reads.fasta, minimap2 2.17 PB flags, reads 21.. of the PAF",
           "name": "example_reads_300_windows", "params": Params().__dict__}
    out.update(record(Reference(), batch, Params()))
    path = os.path.join(ROOT, "tests", "golden", "example_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes;", batch.n_windows, "windows;", sum(out["status"]), "template fall-backs")


if __name__ == "__main__":
    main()
