"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/kwage_cuda.h
declares, validates arguments, and fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from kwage_b200 import capi
from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "kwage_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kwg_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 28
    L = capi.lib()
    for n in names:
        assert hasattr(L, n), "libkwage_cuda.so does not export " + n
    assert sorted(capi.EXPORTS) == names


def test_version_and_launch_counter():
    assert b"sm_100a" in capi.lib().kwg_version()
    assert capi.launch_count() >= 0


def test_built_for_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_validation_precedes_device_use():
    h = C.c_void_p()
    L = capi.lib()
    assert L.kwg_bloom_create(C.byref(h), 0, 33, 1, 20, 24) == capi.KWG_ERR_INVALID_ARG      # k > 32 (word.h:10)
    assert b"word.h" in L.kwg_last_error()
    assert L.kwg_bloom_create(C.byref(h), 0, 31, 16, 20, 24) == capi.KWG_ERR_INVALID_ARG     # min count > 15
    assert L.kwg_bloom_create(C.byref(h), 0, 31, 1, 17, 24) == capi.KWG_ERR_INVALID_ARG      # Lc < 18
    assert L.kwg_bloom_create(C.byref(h), 0, 31, 0, 20, 24) == capi.KWG_ERR_INVALID_ARG      # min count 0
    assert L.kwg_bloom_create_raw(C.byref(h), 0, 31, 9, 20) == capi.KWG_ERR_INVALID_ARG      # > 8 hashes (hash.cpp:243)
    assert L.kwg_transpose(0, None, 4, 64, None) == capi.KWG_ERR_INVALID_ARG


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_fails_loudly_without_a_device():
    with pytest.raises(capi.KwageError) as e:
        capi.BloomBuilder(31, raw_num_hash=3, raw_log2_len=20)
    assert e.value.code == capi.KWG_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.KwageError):
        capi.transpose([np.zeros(8, np.uint8)], 64)
    with pytest.raises(capi.KwageError):
        capi.Database.load(np.zeros((1 << 10, 1), np.uint8), 31, 3, 10, 8)
