"""Resident warps of the wide POA tiers against windows/s at shallow piles (N = 20 / 40: the whole-window POA dominates, SURVEY config 2).
python tools/poa_tier_sweep.py [windows]"""
import json
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200.synth import synth_windows  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cor = Corrector(device=0)
for n_seqs in (20, 40):
    batch = synth_windows(W, n_seqs, seed=42)
    cor.upload(batch)
    cor.run()
    base = cor.download().digest()
    for w1, w2, g in ((0, 0, 0), (2368, 0, 0), (3552, 0, 0), (4736, 0, 0), (2368, 592, 0), (4736, 1184, 0), (2368, 0, 4440)):
        if w1:
            cor.set_option("poa_wide1_warps", w1)
        if w2:
            cor.set_option("poa_wide2_warps", w2)
        if g:
            cor.set_option("poa_g_warps", g)
        ms = []
        for _ in range(3):
            cor.run()
            ms.append(cor.run_ms())
        same = cor.download().digest() == base
        print(json.dumps({"n_seqs": n_seqs, "windows": W, "w1_warps": w1 or "default", "w2_warps": w2 or "default", "g_warps": g or "default",
                          "run_ms": round(min(ms), 2), "windows_per_s": round(W / min(ms) * 1e3), "same_results": same,
                          "stages": {k: round(v["ms"], 1) for k, v in cor.stage_ms().items()}}), flush=True)
    cor.close()
    cor = Corrector(device=0)
