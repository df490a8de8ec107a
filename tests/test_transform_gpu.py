"""GPU parity: DCT/IDCT/DST (IMMA kernels), quant/nquant/dequant/count_nonzero vs the reference C
primitives (oracle/_ref) with the TestBench input shapes (source/test/mbdstharness.cpp:54-115:
residual-range +-PIXEL_MAX inputs, all-min / all-max buffers, srcStride = width) plus full-range
int16 inputs that exercise the forward wrap (dct.cpp:113) and the inverse saturation (dct.cpp:257)."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle
from util import short_buffers, vpo, ssz

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu

SIZES = [(0, 4), (1, 8), (2, 16), (3, 32), (4, 4)]     # (sizeIdx, N); 4 = DST


def _ref(depth):
    R = oracle.ref(depth)
    assert R is not None, "oracle/_ref/libx265ref*.so missing (built by oracle/Makefile in the authoring container)"
    return R


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("full", [False, True])
def test_forward_dct(ctx, depth, full):
    R = _ref(depth)
    for sizeIdx, N in SIZES:
        for bi, buf in enumerate(short_buffers(depth, seed=11 + sizeIdx, size=N * N * 40 + 64, full_range=full)):
            n = 37
            for (blockStride, stride, base) in [(N * N, N, 0), (N * N + 6, N + 2, 1)]:
                if base + (n - 1) * blockStride + (N - 1) * stride + N > len(buf):
                    n = (len(buf) - base - (N - 1) * stride - N) // blockStride
                dS = ctx.to_device(buf)
                dD = ctx.empty(n * N * N * 2)
                ctx.dct_dev(sizeIdx, depth, dS.ptr + base * 2, blockStride, stride, dD, n)
                got = dD.download(np.int16).reshape(n, N * N)
                exp = np.empty((n, N * N), dtype=np.int16)
                for i in range(n):
                    R.ref_dct(sizeIdx, vpo(buf, base + i * blockStride), vpo(exp, i * N * N), ssz(stride))
                assert np.array_equal(got, exp), (depth, full, sizeIdx, bi, stride)
                dS.free(); dD.free()


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("full", [False, True])
def test_inverse_dct(ctx, depth, full):
    R = _ref(depth)
    for sizeIdx, N in SIZES:
        for bi, buf in enumerate(short_buffers(depth, seed=23 + sizeIdx, size=N * N * 33, full_range=full)):
            n = 33
            for (blockStride, stride) in [(N * N, N), (N * (N + 3) + 1, N + 3)]:
                dS = ctx.to_device(buf)
                out_len = n * blockStride + 8
                dD = ctx.to_device(np.full(out_len, 0x3bcd, dtype=np.int16))
                ctx.idct_dev(sizeIdx, depth, dS, dD, blockStride, stride, n)
                got = dD.download(np.int16)
                exp = np.full(out_len, 0x3bcd, dtype=np.int16)       # poison: catches out-of-block writes
                for i in range(n):
                    R.ref_idct(sizeIdx, vpo(buf, i * N * N), vpo(exp, i * blockStride), ssz(stride))
                assert np.array_equal(got, exp), (depth, full, sizeIdx, bi, stride)
                dS.free(); dD.free()


def test_dct_idct_roundtrip_large(ctx):
    """size-independent property at scale: IDCT(DCT(x)) == x for 8-bit residuals (|x| <= 255),
    2^15 blocks of 32x32 (the integer HEVC transform pair is exact to +-1 only after quantisation,
    so we check against the reference on a sample and the energy-compaction DC identity on all)."""
    rng = np.random.default_rng(3)
    n, N = 1 << 12, 32
    x = rng.integers(-255, 256, n * N * N, dtype=np.int64).astype(np.int16)
    dX = ctx.to_device(x)
    dC = ctx.empty(x.nbytes)
    dY = ctx.empty(x.nbytes)
    ctx.dct_dev(3, 8, dX, N * N, N, dC, n)
    ctx.idct_dev(3, 8, dC, dY, N * N, N, n)
    c = dC.download(np.int16).reshape(n, N, N)
    y = dY.download(np.int16).reshape(n, N, N)
    R = _ref(8)
    for i in (0, 1, n // 2, n - 1):
        e = np.empty(N * N, dtype=np.int16)
        R.ref_dct(3, vpo(x, i * N * N), vpo(e, 0), ssz(N))
        assert np.array_equal(c[i].ravel(), e)
    for i in (0, 7, n // 3, n - 1):
        e = np.empty(N * N, dtype=np.int16)
        R.ref_idct(3, vpo(c, i * N * N), vpo(e, 0), ssz(N))
        assert np.array_equal(y[i].ravel(), e)
    # the HEVC integer pair reconstructs to within a few LSB without quantisation
    assert np.abs(y.astype(np.int32) - x.reshape(n, N, N)).max() <= 8
    for b in (dX, dC, dY):
        b.free()


@pytest.mark.parametrize("depth", [8, 10])
def test_quant_nquant_dequant(ctx, depth):
    R = _ref(depth)
    rng = np.random.default_rng(17)
    for log2 in (2, 3, 4, 5):
        numCoeff = 1 << (2 * log2)
        n = 19
        coef = rng.integers(-32768, 32768, n * numCoeff, dtype=np.int64).astype(np.int16)
        coef[:numCoeff] = 0
        coef[numCoeff:2 * numCoeff] = -32768
        qc = rng.integers(1, 1 << 15, numCoeff, dtype=np.int64).astype(np.int32)
        transformShift = 15 - depth - log2
        for qp_per in (0, 3, 7):
            qBits = 14 + qp_per + transformShift
            add = 171 << (qBits - 9)
            dC, dQ = ctx.to_device(coef), ctx.to_device(qc)
            dDu, dOut, dSig = ctx.empty(n * numCoeff * 4), ctx.empty(n * numCoeff * 2), ctx.empty(n * 4)
            ctx.quant_dev(dC, dQ, dDu, dOut, qBits, add, numCoeff, n, dSig)
            gq, gd, gs = dOut.download(np.int16), dDu.download(np.int32), dSig.download(np.uint32)
            eq = np.empty_like(gq); ed = np.empty_like(gd); es = np.empty_like(gs)
            for i in range(n):
                es[i] = R.ref_quant(vpo(coef, i * numCoeff), vpo(qc, 0), vpo(ed, i * numCoeff), vpo(eq, i * numCoeff), qBits, add, numCoeff)
            assert np.array_equal(gq, eq) and np.array_equal(gd, ed) and np.array_equal(gs, es), (depth, log2, qp_per)
            ctx.quant_dev(dC, dQ, None, dOut, qBits, add, numCoeff, n, dSig, nquant=True)
            gq, gs = dOut.download(np.int16), dSig.download(np.uint32)
            for i in range(n):
                es[i] = R.ref_nquant(vpo(coef, i * numCoeff), vpo(qc, 0), vpo(eq, i * numCoeff), qBits, add, numCoeff)
            assert np.array_equal(gq, eq) and np.array_equal(gs, es), ("nquant", depth, log2, qp_per)
            # dequant_normal / scaling on the quantised levels
            dCo = ctx.empty(n * numCoeff * 2)
            q = eq.copy(); q[::3] = -q[::3]
            dQl = ctx.to_device(q)
            for scale, shift in [(40 << qp_per, 6), (72 * 64, 10), (45, 1)]:
                ctx.dequant_normal_dev(dQl, dCo, numCoeff, n, scale, shift)
                g = dCo.download(np.int16); e = np.empty_like(g)
                for i in range(n):
                    R.ref_dequant_normal(vpo(q, i * numCoeff), vpo(e, i * numCoeff), numCoeff, scale, shift)
                assert np.array_equal(g, e), ("dequant_normal", scale, shift)
            deq = rng.integers(16, 1 << 12, numCoeff, dtype=np.int64).astype(np.int32)
            dDq = ctx.to_device(deq)
            for per, shift in [(0, 3), (5, 1), (9, 2), (2, 6)]:
                ctx.dequant_scaling_dev(dQl, dDq, dCo, numCoeff, n, per, shift)
                g = dCo.download(np.int16); e = np.empty_like(g)
                for i in range(n):
                    R.ref_dequant_scaling(vpo(q, i * numCoeff), vpo(deq, 0), vpo(e, i * numCoeff), numCoeff, per, shift)
                assert np.array_equal(g, e), ("dequant_scaling", per, shift)
            dCnt = ctx.empty(n * 4)
            ctx.count_nonzero_dev(dQl, numCoeff, n, dCnt)
            g = dCnt.download(np.int32)
            assert list(g) == [int(np.count_nonzero(q[i * numCoeff:(i + 1) * numCoeff])) for i in range(n)]
            for b in (dC, dQ, dDu, dOut, dSig, dCo, dQl, dDq, dCnt):
                b.free()
