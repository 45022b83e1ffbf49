import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, time
from consent_b200.engine import Corrector
from consent_b200._ffi import Params
from consent_b200.synth import synth_windows
from tests.refs import Oracle
o=Oracle()
c=Corrector(Params(), lib_path="tests/emu/libconsent_emu.so")
for k,v in {"poa_c1_warps": 16, "poa_g_warps": 16, "poa_wide1_warps": 4, "poa_wide2_warps": 4, "poa_tier1_warps": 2, "poa_tier2_warps": 2, "poa_tier1_cells": 8 << 20, "poa_tier2_cells": 16 << 20}.items(): c.set_option(k,v)
for n,nw,seed in ((150,40,42),(20,30,7),(8,40,9),(47,20,11),(3,50,12),(150,30,77)):
    b=synth_windows(nw,n,seed=seed)
    t=time.time(); got=c.correct_windows(b); t1=time.time()-t
    want,_=o.correct_windows(b,threads=8)
    ok=got.equals(want)
    print(n,nw,seed,"OK" if ok else f"MISMATCH at {got.first_mismatch(want)}", f"{t1:.1f}s", flush=True)
    o.lib.oracle_reset_counters(); o.correct_windows(b,threads=1)
    oc,gc=o.counters(),c.counters()
    bad=[k for k in oc if oc[k]!=gc[k]]
    if bad: print("  counter mismatch",bad,[(oc[k],gc[k]) for k in bad])
