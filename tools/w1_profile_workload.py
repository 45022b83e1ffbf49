import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from consent_b200.engine import Corrector
from consent_b200.synth import synth_windows
cor = Corrector(device=0)
batch = synth_windows(4000, 20, seed=42)
cor.upload(batch); cor.run(); cor.run()
print(cor.run_ms(), cor.stage_ms())
