"""The wide POA tier's workload for profilers / sanitizers: windows of 20 sequences (whole-window POA).  python tools/w1_profile_workload.py [windows]"""
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200.synth import synth_windows  # noqa: E402

cor = Corrector(device=0)
batch = synth_windows(int(sys.argv[1]) if len(sys.argv) > 1 else 4000, 20, seed=42)
cor.upload(batch)
cor.run()
cor.run()
print(cor.run_ms(), cor.stage_ms(), cor.download().digest())
