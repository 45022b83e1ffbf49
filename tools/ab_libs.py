import sys, os, json
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from consent_b200.engine import Corrector
from consent_b200._ffi import Params
from consent_b200.synth import synth_windows
batch = synth_windows(16384, 150, seed=42)
for lib in sys.argv[1:]:
    cor = Corrector(Params(), lib_path=lib)
    cor.upload(batch); cor.run()
    ms=[]
    for _ in range(3):
        cor.run(); ms.append(cor.run_ms())
    print(lib, round(min(ms),2), {k: round(v["ms"],2) for k,v in cor.stage_ms().items()}, cor.download().digest()[:12], flush=True)
    cor.close()
