"""Where a fresh process spends its first seconds: python tools/startup_probe.py  (prints seconds per step; no torch import)."""
import ctypes as C
import os
import sys
import time

t0 = time.time()
root = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(os.path.join(root, "consent_b200", "libconsent_b200.so"))
t1 = time.time()


class Params(C.Structure):
    _fields_ = [("mer_size", C.c_uint32), ("solid_thresh", C.c_uint32), ("common_kmers", C.c_uint32), ("min_anchors", C.c_uint32)]


lib.cg_device_count.restype = C.c_int
n = lib.cg_device_count()
t2 = time.time()
h = C.c_void_p()
p = Params(9, 4, 8, 2)
rc = lib.cg_create(0, C.byref(p), C.byref(h))
t3 = time.time()
h2 = C.c_void_p()
rc2 = lib.cg_create(0, C.byref(p), C.byref(h2))
t4 = time.time()
print({"dlopen_s": round(t1 - t0, 3), "device_count_s (cuInit)": round(t2 - t1, 3), "devices": n, "first cg_create_s (context + module)": round(t3 - t2, 3),
       "second cg_create_s": round(t4 - t3, 3), "rc": (rc, rc2)})
