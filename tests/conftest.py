import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than ~10 s on CPU")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (oracle) and the product library once per session."""
    from oracle import oracle_py
    if not os.path.exists(oracle_py.LIB_PATH):
        oracle_py.build()
    from kwage_b200 import build as kbuild
    kbuild.build()
    emul_so = os.path.join(ROOT, "tests", "host_emul", "libkwage_emul.so")
    src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
    hdrs = [os.path.join(ROOT, "kwage_b200", "csrc", h) for h in ("bitops.cuh", "crc_tables.h")]
    if (not os.path.exists(emul_so)) or os.path.getmtime(emul_so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", emul_so, src], check=True)
    yield


def load_golden(name):
    with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def pytest_collection_modifyitems(config, items):
    # a plain `pytest` on a CPU box skips the GPU tests instead of failing them (the driver selects with -m)
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def has_gpu():
    try:
        from kwage_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False
