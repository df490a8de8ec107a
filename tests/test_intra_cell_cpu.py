"""CPU: the staged cell form of the 8-bit all-modes intra prediction (csrc/intra_cell.cuh -- one thread per 16 output bytes, no
shared memory; switched on by X265B200_INTRA_FAST=1) executed on the host for every thread of its grid and compared with the
oracle's per-mode predictions (intrapred.cpp:31-204 restatement) in the all_angs layout: smoothed / raw neighbours chosen by the
distance thresholds, horizontal modes un-transposed, planar + DC in front for the 35-mode form.  Accesses are bounds-checked."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from util import oracle, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("xc") / "intra_cell_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "x265-yuuki-asuna_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), "-o", so, os.path.join(ROOT, "tests", "host_emu", "intra_cell_emu.cpp")], check=True)
    return ctypes.CDLL(so)


def _expected(nbr, filt, n, log2N, bLuma, all35):
    O = oracle.orc()
    N = 1 << log2N
    L = 4 * N + 1
    thr = {3: 7, 4: 1, 5: 0}[log2N]
    modes = list(range(35)) if all35 else list(range(2, 35))
    exp = np.zeros((n, len(modes), N, N), dtype=np.uint8)
    for i in range(n):
        for k, mode in enumerate(modes):
            if mode == 0:
                src, bf = filt, 0
            elif mode == 1:
                src, bf = nbr, bLuma
            else:
                src, bf = (filt if min(abs(mode - 26), abs(mode - 10)) > thr else nbr), bLuma
            e = np.empty((N, N), dtype=np.uint8)
            O.orc_intra_pred(8, log2N, mode, bf, ctypes.c_void_p(src.ctypes.data + i * L), vp(e), ctypes.c_ssize_t(N))
            exp[i, k] = e.T if 2 <= mode < 18 else e
    return exp


@pytest.mark.parametrize("log2N", [3, 4, 5])
@pytest.mark.parametrize("all35,bLuma,fill,misalign", [(1, 1, "rand", 0), (1, 0, "rand", 0), (0, 1, "rand", 0), (1, 1, "extreme", 0), (1, 1, "rand", 4), (0, 0, "rand", 3)])
def test_cell_intra_equals_oracle(emu, log2N, all35, bLuma, fill, misalign):
    N = 1 << log2N
    L = 4 * N + 1
    n = 7
    rng = np.random.default_rng(1000 + log2N * 10 + all35)
    nbr = rng.integers(0, 256, n * L, dtype=np.uint8) if fill == "rand" else rng.choice(np.array([0, 255], dtype=np.uint8), n * L)
    filt = np.zeros_like(nbr)
    O = oracle.orc()
    for i in range(n):
        O.orc_intra_filter(8, log2N, ctypes.c_void_p(nbr.ctypes.data + i * L), ctypes.c_void_p(filt.ctypes.data + i * L))
    NM = 35 if all35 else 33
    total = n * NM * N * N
    buf = np.full(total + 32, 0xCD, dtype=np.uint8)
    base = (-buf.ctypes.data) % 16 + misalign                       # 16-byte aligned destination, or deliberately not
    emu.xc_run(vp(nbr), vp(filt), ctypes.c_size_t(nbr.nbytes), ctypes.c_void_p(buf.ctypes.data + base), ctypes.c_size_t(total), log2N, bLuma, all35, ctypes.c_int64(n))
    got = buf[base:base + total].reshape(n, NM, N, N)
    exp = _expected(nbr, filt, n, log2N, bLuma, all35)
    bad = np.argwhere(got != exp)
    assert not len(bad), (log2N, all35, bLuma, bad[0].tolist(), len(bad))
    assert np.all(buf[:base] == 0xCD) and np.all(buf[base + total:] == 0xCD)
