"""A small resident workload for profilers / sanitizers: python tools/profile_workload.py [windows] [seqs per window] [runs]
(config-3 shape by default: 150 sequences per window).  Prints the run time, the per-stage CUDA-event times and the result digest."""
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200.synth import synth_windows  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 150
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cor = Corrector(device=0)
batch = synth_windows(W, N, seed=42)
cor.upload(batch)
for _ in range(runs):
    cor.run()
print(cor.run_ms(), cor.stage_ms(), cor.download().digest())
