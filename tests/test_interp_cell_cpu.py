"""CPU: the staged cell form of the 8-bit luma pp interpolation (csrc/interp_cell.cuh -- one thread per 4x4 output cell on packed
words, switched on by X265B200_INTERP_FAST=1) executed on the host for every thread of its grid and compared with the oracle's
luma_hpp / luma_vpp / luma_hvpp (ipfilter.cpp:79-118, 164-203, 362-369); 32-bit accesses are checked for alignment and bounds."""
import ctypes
import importlib
import os
import subprocess

import numpy as np
import pytest

from util import oracle, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("x265-yuuki-asuna_b200")
KINDS = {"hpp": (0, 0), "vpp": (2, 2), "hvpp": (6, 6)}          # name -> (X265B200_IP_* kind, orc_interp kind)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ic") / "interp_cell_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "x265-yuuki-asuna_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), "-o", so, os.path.join(ROOT, "tests", "host_emu", "interp_cell_emu.cpp")], check=True)
    return ctypes.CDLL(so)


@pytest.mark.parametrize("name", sorted(KINDS))
@pytest.mark.parametrize("w,h", [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (16, 12), (12, 16), (24, 32), (64, 16), (48, 64)])
def test_cell_interpolation_equals_oracle(emu, name, w, h):
    kind, okind = KINDS[name]
    rng = np.random.default_rng(w * 131 + h + kind)
    S, R, pad = 160, 160, 16                                      # source plane (stride multiple of 4), blocks stay `pad` inside
    src = rng.integers(0, 256, R * S, dtype=np.uint8)
    src[rng.integers(0, R * S, 4000)] = 255
    src[rng.integers(0, R * S, 4000)] = 0
    njobs = 23
    DS = 68 if (w + h) % 8 else 67                                # destination pitch: multiple of 4 or not (byte-store path)
    dst = np.zeros(njobs * 64 * DS + 64, dtype=np.uint8)
    jobs = np.zeros(njobs, dtype=pkg.INTERP_JOB)
    for j in range(njobs):
        x = int(rng.integers(pad, S - pad - w)); y = int(rng.integers(pad, R - pad - h))
        jobs[j]["srcOff"] = y * S + x
        jobs[j]["dstOff"] = j * 64 * DS + (int(rng.integers(0, 4)) if j % 3 == 0 else 0)        # aligned and misaligned destinations
        jobs[j]["idxX"] = int(rng.integers(1, 4))
        jobs[j]["idxY"] = int(rng.integers(1, 4))
    want = dst.copy()
    O = oracle.orc()
    for j in range(njobs):
        O.orc_interp(8, okind, 8, w, h, ctypes.c_void_p(src.ctypes.data + int(jobs[j]["srcOff"])), ctypes.c_ssize_t(S),
                     ctypes.c_void_p(want.ctypes.data + int(jobs[j]["dstOff"])), ctypes.c_ssize_t(DS), int(jobs[j]["idxX"]), int(jobs[j]["idxY"]), 0)
    rc = emu.ic_run(vp(src), ctypes.c_size_t(src.nbytes), ctypes.c_int64(S), vp(dst), ctypes.c_size_t(dst.nbytes), ctypes.c_int64(DS),
                    vp(jobs), ctypes.c_int64(njobs), kind, w, h)
    assert rc == 0
    assert np.array_equal(dst, want), int(np.count_nonzero(dst != want))
