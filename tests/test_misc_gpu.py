"""GPU parity (batched form, n > 1) for the small table entries next to the hot path: ssim_4x4x2_core / ssim_end_4 (float,
BIT-exact: every operation individually rounded in the reference's order), planeClipAndMax (10-bit), propagateCost incl. the
degenerate intraCost == 0 inputs (x86 conversion semantics)."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle
from util import pdtype, vp, vpo, ssz

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("depth", [8, 10])
def test_ssim_primitives_bit_exact(ctx, depth):
    R = oracle.ref(depth)
    R.ref_ssim_end4.restype = ctypes.c_float
    rng = np.random.default_rng(depth)
    n = 300
    hi = 16 * 2 * ((1 << depth) - 1)
    s0 = np.zeros((n, 5, 4), dtype=np.int32); s1 = np.zeros((n, 5, 4), dtype=np.int32)
    for arr in (s0, s1):
        arr[:, :, 0] = rng.integers(0, hi, (n, 5)); arr[:, :, 1] = rng.integers(0, hi, (n, 5))
        arr[:, :, 2] = rng.integers(0, hi * ((1 << depth) - 1) // 2, (n, 5)); arr[:, :, 3] = rng.integers(0, hi * ((1 << depth) - 1) // 4, (n, 5))
    widths = rng.integers(1, 5, n).astype(np.int32)
    d0, d1, dW, dO = ctx.to_device(s0), ctx.to_device(s1), ctx.to_device(widths), ctx.empty(n * 4)
    ctx.ssim_end4_dev(depth, d0, d1, dW, n, dO)
    got = dO.download(np.float32)
    exp = np.array([R.ref_ssim_end4(vp(s0[i]), vp(s1[i]), int(widths[i])) for i in range(n)], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    # ssim_4x4x2_core over a list of block pairs
    S = 80
    a = rng.integers(0, 1 << depth, S * 40).astype(pdtype(depth)); b = rng.integers(0, 1 << depth, S * 40).astype(pdtype(depth))
    off = (rng.integers(0, 30, n) * S + rng.integers(0, 60, n)).astype(np.int64)
    dA, dB, dOff, dS = ctx.to_device(a), ctx.to_device(b), ctx.to_device(off), ctx.empty(n * 32)
    ctx.ssim_4x4x2_dev(depth, dA, S, dB, S, dOff, dOff, n, dS)
    gs = dS.download(np.int32).reshape(n, 2, 4)
    for i in range(n):
        e = np.zeros((2, 4), dtype=np.int32)
        R.ref_ssim_core(vpo(a, int(off[i])), ssz(S), vpo(b, int(off[i])), ssz(S), vp(e))
        assert np.array_equal(gs[i], e), i
    for bfr in (d0, d1, dW, dO, dA, dB, dOff, dS):
        bfr.free()


def test_plane_clip_and_max_10bit(ctx):
    R = oracle.ref(10)
    rng = np.random.default_rng(4)
    S, Wd, Hh = 208, 200, 37
    src = rng.integers(0, 1024, S * Hh).astype(np.uint16)
    e = src.copy(); esum = ctypes.c_uint64(0)
    emax = R.ref_plane_clip_max(vp(e), ssz(S), Wd, Hh, ctypes.byref(esum), 64, 940)
    assert emax >= 0
    dS, dO = ctx.to_device(src), ctx.empty(16)
    ctx.plane_clip_max_dev(10, dS, S, Wd, Hh, 64, 940, dO.ptr, dO.ptr + 8)
    out = dO.download(np.uint64)
    assert np.array_equal(dS.download(np.uint16), e) and int(out[0]) == esum.value and int(out[1] & 0xffffffff) == emax
    dS.free(); dO.free()


def test_propagate_cost_batch(ctx):
    R = oracle.ref(8)
    rng = np.random.default_rng(8)
    n = 5000
    pin = rng.integers(0, 65536, n).astype(np.uint16)
    intra = rng.integers(0, 40000, n).astype(np.int32)
    intra[::97] = 0                                               # division by zero -> x86 "integer indefinite"
    inter = rng.integers(0, 65536, n).astype(np.uint16)
    invq = rng.integers(1, 70000, n).astype(np.int32)
    exp = np.zeros(n, dtype=np.int32)
    R.ref_propagate_cost(vp(exp), vp(pin), vp(intra), vp(inter), vp(invq), ctypes.c_double(0.75 * 256), n)
    bufs = [ctx.to_device(x) for x in (pin, intra, inter, invq)]
    dD = ctx.empty(n * 4)
    ctx.propagate_cost_dev(dD, bufs[0], bufs[1], bufs[2], bufs[3], 0.75 * 256, n)
    assert np.array_equal(dD.download(np.int32), exp)
    for bfr in bufs + [dD]:
        bfr.free()
