"""Shallow-pile throughput (the wide POA tier): python tools/w1_quick.py [windows] [lib ...]"""
import json, os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200._ffi import Params  # noqa: E402
from consent_b200.synth import synth_windows  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
libs = sys.argv[2:] or [None]
for n_seqs in (20, 40):
    batch = synth_windows(W, n_seqs, seed=42)
    for lib in libs:
        cor = Corrector(Params(), lib_path=lib) if lib else Corrector(device=0)
        cor.upload(batch)
        cor.run()
        ms = []
        for _ in range(3):
            cor.run()
            ms.append(cor.run_ms())
        print(json.dumps({"lib": lib or "default", "n_seqs": n_seqs, "windows": W, "run_ms": round(min(ms), 2), "windows_per_s": round(W / min(ms) * 1e3),
                          "digest": cor.download().digest()[:12], "stages": {k: round(v["ms"], 1) for k, v in cor.stage_ms().items()}}), flush=True)
        cor.close()
