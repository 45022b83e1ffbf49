"""Chunk-size sweep of the resident window path: python tools/chunk_sweep.py [windows] (config-3 shape; results never depend on the chunking)."""
import json, os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200.synth import synth_windows  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
batch = synth_windows(W, 150, seed=42)
for chunk, budget_gb in ((8192, 8), (16384, 8), (22000, 12), (32768, 16)):
    cor = Corrector(device=0)
    cor.set_option("chunk_max_windows", chunk)
    cor.set_option("chunk_budget_bytes", budget_gb << 30)
    cor.upload(batch)
    cor.run()
    ms = []
    for _ in range(2):
        cor.run(); ms.append(cor.run_ms())
    print(json.dumps({"chunk_max_windows": chunk, "chunks": cor.chunk_count() if hasattr(cor, "chunk_count") else None, "run_ms": round(min(ms), 1),
                      "windows_per_s": round(W / min(ms) * 1e3), "digest": cor.download().digest()[:12]}), flush=True)
    cor.close()
