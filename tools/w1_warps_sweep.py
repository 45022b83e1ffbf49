"""python tools/w1_warps_sweep.py [windows] [n_seqs]: resident warps of the wide tier vs throughput"""
import json, os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200.synth import synth_windows  # noqa: E402
W = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
batch = synth_windows(W, N, seed=42)
for per_sm in (4, 8, 12, 16, 20, 24):
    cor = Corrector(device=0)
    cor.set_option("poa_wide1_warps", 148 * per_sm)
    cor.upload(batch); cor.run()
    ms = []
    for _ in range(3):
        cor.run(); ms.append(cor.run_ms())
    print(json.dumps({"w1_warps_per_sm": per_sm, "run_ms": round(min(ms), 2), "windows_per_s": round(W / min(ms) * 1e3), "poa_ms": round(cor.stage_ms()["poa"]["ms"], 1)}), flush=True)
    cor.close()
