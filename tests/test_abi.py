"""The C-ABI libraries load on a CPU-only box and export every symbol their headers
declare (no compute calls here)."""
import ctypes
import os
import re

import pytest

import saamge_b200 as sab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, txt)))


def test_gpu_abi_symbols():
    lib = sab.gpu_lib()
    names = declared("saamge_b200.h", "sa_gpu_")
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert missing == []


def test_level_desc_layout_matches_ctypes_mirror():
    from saamge_b200 import cabi

    lib = sab.gpu_lib()
    lib.sa_gpu_level_desc_size.restype = ctypes.c_size_t
    assert lib.sa_gpu_level_desc_size() == ctypes.sizeof(cabi.LevelDesc)


def test_driver_abi_symbols():
    lib = sab.host_lib()
    names = declared("saamge_b200_driver.h", "sa_drv_")
    missing = [n for n in names if not hasattr(lib, n)]
    assert missing == []


def test_no_silent_cpu_fallback():
    """Without a CUDA device the context cannot be created and says so."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = sab.gpu_lib()
    ctx = ctypes.c_void_p()
    rc = lib.sa_gpu_ctx_create(0, ctypes.byref(ctx))
    assert rc != 0
    assert b"no CPU fallback" in lib.sa_gpu_last_error() or b"CUDA" in lib.sa_gpu_last_error()


def test_product_does_not_reference_oracle():
    """Nothing under saamge_b200/ or include/ may include, link or import oracle/."""
    bad = []
    for base in ("saamge_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".so", ".log", ".pyc")):
                    continue
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"#include\s+\"[^\"]*oracle|import\s+oracle|liboracle|sa_orc_", txt):
                    bad.append(os.path.join(dp, fn))
    assert bad == []
