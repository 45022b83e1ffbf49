"""Debug: per-job phase clocks of the POA tiers (builds alt/libconsent_timing.so with -DCG_POA_TIMING if it is not there; run from the repo root).
python tools/poa_job_timing.py [windows] [n_seqs]"""
import ctypes as C, os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200._ffi import Params  # noqa: E402
from consent_b200.synth import synth_windows  # noqa: E402
import numpy as np  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = "alt/libconsent_timing.so"
if not os.path.exists(lib):                              # the instrumented build (per-job phase clocks): not part of build()
    import subprocess
    os.makedirs("alt", exist_ok=True)
    subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                    "-fmad=false", "-DCG_POA_TIMING", "-Xcompiler", "-fPIC", "-shared", "-I", "include", "-I", "consent_b200/csrc", "-o", lib,
                    "consent_b200/csrc/consent_b200.cu", "-ldl"], check=True)
cor = Corrector(Params(), lib_path=lib)
cor.upload(synth_windows(W, N, seed=42))
cor.run()
cor.lib.cg_debug_dump_jobs(b"/dev/null")
cor.run()
print("run_ms", cor.run_ms(), {k: round(v["ms"], 1) for k, v in cor.stage_ms().items()})
out = "gpurun_out/poa_jobs_%d_%d.txt" % (W, N)
os.makedirs("gpurun_out", exist_ok=True)
n = cor.lib.cg_debug_dump_jobs(out.encode())
a = np.loadtxt(out, skiprows=1, dtype=np.int64)
print("jobs", n)
names = ["dp", "maxtie", "traceback", "update", "splice", "dfs", "setup", "vote"]
for tier in sorted(set(a[:, 0])):
    t = a[a[:, 0] == tier]
    dur = t[:, 7] - t[:, 6]
    tot = t[:, 8:16].sum(0)
    print(f"tier VCAP={tier}: jobs {len(t)}  duration cycles: mean {dur.mean():.0f} p50 {np.median(dur):.0f} p90 {np.percentile(dur, 90):.0f} max {dur.max()}  sum {dur.sum() / 1e9:.2f} Gcyc")
    print("   phase share:", ", ".join(f"{nm} {v / tot.sum():.1%}" for nm, v in zip(names, tot)))
    span = t[:, 7].max() - t[:, 6].min()
    print(f"   first start -> last end: {span} cycles; warps x span = {span / 1e9:.3f} Gcyc per warp")
    big = t[np.argsort(-dur)[:5]]
    for j in big:
        print("   big:", dict(w=int(j[1]), rg=int(j[2]), nseg=int(j[3]), V=int(j[4]), maxL=int(j[5]), start=int(j[6] - t[:, 6].min()), dur=int(j[7] - j[6])), dict(zip(names, (int(x) for x in j[8:16]))))
