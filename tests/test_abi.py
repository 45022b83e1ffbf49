"""CPU suite, part 2: the C-ABI library loads, exports every symbol the header declares, and refuses to run
without a CUDA device (no CPU fallback behind the ABI)."""
import ctypes as C
import os
import re

import pytest

from consent_b200 import engine
from consent_b200._ffi import Params, cg_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "consent_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_boundary():
    fns = header_functions()
    assert {"cg_create", "cg_destroy", "cg_correct_windows", "cg_upload", "cg_run", "cg_download", "cg_free_results",
            "cg_last_error", "cg_stage_ms", "cg_run_ms", "cg_get_counters", "cg_set_option", "cg_abi_version", "cg_device_count"} <= set(fns)


def test_library_exports_every_declared_symbol(gpu_lib):
    lib = C.CDLL(gpu_lib)
    for fn in header_functions():
        assert hasattr(lib, fn), f"{fn} declared in include/consent_b200.h but not exported"
    assert set(header_functions()) == set(engine.EXPORTS)
    lib.cg_abi_version.restype = C.c_int
    assert lib.cg_abi_version() == 5


def test_cuda_library_holds_sm100a_kernels(gpu_lib):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", gpu_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_behind_the_abi(gpu_lib):
    """Without a GPU cg_create must fail loudly (CG_ERR_NO_DEVICE), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = engine.bind(C.CDLL(gpu_lib))
    h = C.c_void_p()
    p = Params().c()
    rc = lib.cg_create(0, C.byref(p), C.byref(h))
    assert rc == -2 and not h.value
    assert b"no usable CUDA device" in lib.cg_last_error(None)
    with pytest.raises(engine.ConsentError):
        engine.Corrector()


def test_product_never_touches_the_oracle():
    """Nothing under consent_b200/ may import, link or load oracle/ or the emulator."""
    pkg = os.path.join(ROOT, "consent_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(d, f)).read()
                assert "liboracle" not in text and "consent_oracle" not in text and "libconsent_emu" not in text, f
                assert "tests.refs" not in text and "from tests" not in text, f


def test_invalid_parameters_are_rejected(gpu_lib):
    lib = engine.bind(C.CDLL(gpu_lib))
    h = C.c_void_p()
    for bad in (cg_params(1, 4, 8, 2), cg_params(16, 4, 8, 2), cg_params(9, 0, 8, 2)):
        assert lib.cg_create(0, C.byref(bad), C.byref(h)) == -1
    # k = 10 .. 15 are valid since round 2 (hashed k-mer index): on a box without a GPU the only possible outcome is CG_ERR_NO_DEVICE —
    # the product has no CPU path
    k12 = cg_params(12, 4, 8, 2)
    rc = lib.cg_create(0, C.byref(k12), C.byref(h))
    assert rc in (0, -2)
    if rc == 0:
        lib.cg_destroy(h)
    else:
        assert b"no CPU path" in lib.cg_last_error(None)
