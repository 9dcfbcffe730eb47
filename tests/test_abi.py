"""The C-ABI library loads and exports every symbol include/ytb200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

import pytest

from yt_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ytb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ytb200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_something():
    syms = header_symbols()
    assert len(syms) >= 25 and "ytb200_frame_cast" in syms


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_abi_version_and_error_string(lib):
    lib.ytb200_abi_version.restype = ctypes.c_int
    assert lib.ytb200_abi_version() == 1
    lib.ytb200_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.ytb200_last_error(), bytes)


def test_no_gpu_fails_loudly():
    """Without a device the product path raises; it never falls back to a CPU implementation."""
    L = _lib.load()
    n = ctypes.c_int(-1)
    rc = L.ytb200_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    from yt_b200 import device
    with pytest.raises(_lib.YTB200Error):
        device.Context(0)


def test_product_never_imports_oracle():
    """Nothing under yt_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "yt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"\boracle\b", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
