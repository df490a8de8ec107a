"""CPU: the C-ABI library loads and exports every symbol include/x265b200.h declares, and the
compute path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import importlib
import os

import pytest

pkg = importlib.import_module("x265-yuuki-asuna_b200")


def test_library_exports_every_declared_symbol():
    L = pkg.load()
    syms = pkg.header_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    L = pkg.load()
    if L.x265b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.X265B200Error):
        pkg.Ctx(0)
    # compute entry points refuse a NULL context instead of computing on the host
    rc = L.x265b200_pixelcmp_host(None, 0, 8, 8, 8, None, ctypes.c_size_t(0), ctypes.c_int64(0), None,
                                  ctypes.c_size_t(0), ctypes.c_int64(0), None, None, ctypes.c_int64(0), None)
    assert rc != 0


def test_product_does_not_reference_oracle():
    """The product sources must never include, link or call anything under oracle/."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pdir = os.path.join(root, "x265-yuuki-asuna_b200")
    bad = []
    for dp, _, fns in os.walk(pdir):
        for fn in fns:
            if fn.endswith((".cu", ".cuh", ".cpp", ".h", ".py", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if "oracle" in txt and "orc_" in txt or "liboracle" in txt or "libx265ref" in txt:
                    bad.append(fn)
    assert not bad, bad
