"""Shared helpers for the parity tests: TestBench-shaped fixtures and oracle call wrappers.

Fixture shapes mirror the reference's own unit test (source/test/pixelharness.cpp:31-101,
pixelharness.h:36-45): three buffers per type -- random, all-min, all-max -- of
BUFFSIZE = 64*(64+64) + 32*100 pixels, stride 64, block i starting at offset 32*i;
sad_x3/x4 use the unaligned reference stride FENC_STRIDE-5 = 59 (pixelharness.cpp:150,178)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle  # noqa: E402  (test infrastructure)

BUFFSIZE = 64 * (64 + 64) + 32 * 100
STRIDE = 64
ITERS = 24
INCR = 32
ssz = ctypes.c_ssize_t

LUMA_PU_SIZES = [
    (4, 4), (8, 8), (16, 16), (32, 32), (64, 64),
    (8, 4), (4, 8), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64),
    (16, 12), (12, 16), (16, 4), (4, 16), (32, 24), (24, 32), (32, 8), (8, 32),
    (64, 48), (48, 64), (64, 16), (16, 64),
]
CU_SIZES = [4, 8, 16, 32, 64]


def pdtype(depth):
    return np.uint8 if depth == 8 else np.uint16


def pixel_buffers(depth, seed=1234, size=BUFFSIZE):
    """[random, all-min, all-max] pixel buffers (pixelharness.cpp:33-65)."""
    rng = np.random.default_rng(seed)
    pmax = (1 << depth) - 1
    dt = pdtype(depth)
    return [rng.integers(0, pmax + 1, size, dtype=np.int64).astype(dt),
            np.zeros(size, dtype=dt), np.full(size, pmax, dtype=dt)]


def short_buffers(depth, seed=4321, size=BUFFSIZE, full_range=False):
    """int16 residual-range buffers: random +-PIXEL_MAX, all -PIXEL_MAX, all +PIXEL_MAX."""
    rng = np.random.default_rng(seed)
    pmax = 32767 if full_range else (1 << depth) - 1
    lo = -32768 if full_range else -pmax
    return [rng.integers(lo, pmax + 1, size, dtype=np.int64).astype(np.int16),
            np.full(size, lo, dtype=np.int16), np.full(size, pmax, dtype=np.int16)]


def vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def vpo(a, off):
    """void* to element `off` of array a."""
    return ctypes.c_void_p(a.ctypes.data + int(off) * a.itemsize)


def block_offsets(n=ITERS, incr=INCR):
    return np.arange(n, dtype=np.int64) * incr


# ---- oracle wrappers (restatement) ---------------------------------------------------------
def orc_cmp(kind, depth, w, h, A, sa, B, sb, offA, offB):
    """kind: 'sad' 'satd' 'sa8d' 'sa8d8' 'sse_pp' 'sse_ss' 'ssd_s' -> list of python ints."""
    L = oracle.orc()
    out = []
    for oa, ob in zip(offA, offB):
        pa, pb = vpo(A, oa), vpo(B, ob)
        if kind == "sad":
            v = L.orc_sad(depth, w, h, pa, ssz(sa), pb, ssz(sb))
        elif kind == "satd":
            v = L.orc_satd(depth, w, h, pa, ssz(sa), pb, ssz(sb))
        elif kind == "sa8d":
            v = L.orc_sa8d(depth, w, h, int(w >= 16 and h >= 16), pa, ssz(sa), pb, ssz(sb))
        elif kind == "sa8d8":
            v = L.orc_sa8d(depth, w, h, 0, pa, ssz(sa), pb, ssz(sb))
        elif kind == "sse_pp":
            v = L.orc_sse_pp(depth, w, h, pa, ssz(sa), pb, ssz(sb))
        elif kind == "sse_ss":
            v = L.orc_sse_ss(depth, w, h, pa, ssz(sa), pb, ssz(sb))
        elif kind == "ssd_s":
            v = L.orc_ssd_s(depth, w, pa, ssz(sa))
        else:
            raise ValueError(kind)
        out.append(int(v))
    return out


def ref_cmp(kind, depth, w, h, A, sa, B, sb, offA, offB):
    """Same through the compiled reference (oracle/_ref); None when it is not available or the
    reference table has no such slot."""
    R = oracle.ref(depth)
    if R is None:
        return None
    out = []
    for oa, ob in zip(offA, offB):
        pa, pb = vpo(A, oa), vpo(B, ob)
        if kind in ("sad", "satd"):
            if (w, h) not in LUMA_PU_SIZES:
                return None
            v = R.ref_pixelcmp(0 if kind == "sad" else 1, LUMA_PU_SIZES.index((w, h)), pa, ssz(sa), pb, ssz(sb))
        elif kind == "sa8d":
            if w != h or w not in CU_SIZES:
                return None
            v = R.ref_pixelcmp(2, CU_SIZES.index(w), pa, ssz(sa), pb, ssz(sb))
        elif kind == "sa8d8":
            if (w, h) == (8, 8):
                v = R.ref_pixelcmp(5, 2, pa, ssz(sa), pb, ssz(sb))     # chroma[420].cu[16x16].sa8d = sa8d8<8,8>
            elif (w, h) == (8, 16):
                v = R.ref_pixelcmp(6, 2, pa, ssz(sa), pb, ssz(sb))     # chroma[422].cu[16x16].sa8d = sa8d8<8,16>
            else:
                return None
        elif kind in ("sse_pp", "sse_ss", "ssd_s"):
            if w != h or w not in CU_SIZES:
                return None
            v = R.ref_sse({"sse_pp": 0, "sse_ss": 1, "ssd_s": 2}[kind], CU_SIZES.index(w), pa, ssz(sa), pb, ssz(sb))
        else:
            raise ValueError(kind)
        out.append(int(v))
    return out
