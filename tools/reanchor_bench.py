import sys, time, os, json
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from consent_b200.synth import synth_reads
from consent_b200.engine import Corrector
from consent_b200._ffi import Results
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
lib = sys.argv[2] if len(sys.argv) > 2 else None        # another build of the library (A/B runs)
from consent_b200._ffi import Params
t=time.time(); batch, reads = synth_reads(n_reads, 12, truth_len=8000, seed=42); print("gen", time.time()-t, batch.n_windows, "windows", flush=True)
cor = Corrector(Params(), lib_path=lib) if lib else Corrector(device=0)
live = cor.correct_windows(batch)
for i in range(3):
    t=time.time(); got = cor.reanchor_reads(batch, live, reads); dt=time.time()-t
    st = cor.reanchor_stats()
    print(json.dumps({"mode":"resident","wall_s":dt, **st, "win_per_s_kernel": batch.n_windows/(st["kernel_ms"]/1e3), "gcups": st["dp_cells"]/st["kernel_ms"]/1e6}), flush=True)
host = Results(live._r)
for i in range(2):
    t=time.time(); got2 = cor.reanchor_reads(batch, host, reads); dt=time.time()-t
    st = cor.reanchor_stats()
    print(json.dumps({"mode":"upload","wall_s":dt, **st}), flush=True)
print("same", got.equals(got2))
