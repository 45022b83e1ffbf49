import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def entry():
    import __graft_entry__ as g
    g.build_host()
    g.build_oracle()
    return g


@pytest.fixture(scope="session")
def oracle(entry):
    from tests.refs import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference(entry):
    from tests.refs import Reference, ref_library_path
    if ref_library_path() is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference) or host lacks SSE4.1")
    return Reference()


@pytest.fixture(scope="session")
def example_golden():
    """Reference outputs on 300 real-data windows of the shipped example (tests/golden/make_example_golden.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "example_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def emu(entry):
    """The product's kernel sources running on the SIMT emulator (tests/emu) — CPU-side check of the real code."""
    from consent_b200.engine import Corrector
    path = entry.build_emu()

    def make(params=None, **options):
        from consent_b200._ffi import Params
        c = Corrector(params or Params(), lib_path=path)
        opts = {"poa_c1_warps": 16, "poa_g_warps": 16, "poa_wide1_warps": 4, "poa_wide2_warps": 4,
                "poa_tier1_warps": 2, "poa_tier2_warps": 2, "poa_tier1_cells": 8 << 20, "poa_tier2_cells": 16 << 20}
        opts.update(options)
        for k, v in opts.items():
            c.set_option(k, v)
        return c
    return make


@pytest.fixture(scope="session")
def gpu_lib(entry):
    return entry.build_cuda()


@pytest.fixture(scope="session")
def gpu(gpu_lib):
    from consent_b200.engine import Corrector

    def make(params=None, **options):
        from consent_b200._ffi import Params
        c = Corrector(params or Params(), device=0)
        for k, v in options.items():
            c.set_option(k, v)
        return c
    return make
