"""GPU parity: the element-wise "glue" entries in their BATCHED form
(many jobs per launch, ragged offsets and strides) against the reference's own table (oracle/_ref).  The one-call
form of every slot is additionally pinned by the reference's own TestBench (tests/test_testbench_gpu.py)."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle
from util import pdtype, pixel_buffers, short_buffers, ssz, vpo

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu

S = 96          # plane stride of the fixtures (elements)


def _ref(depth):
    R = oracle.ref(depth)
    assert R is not None, "oracle/_ref missing"
    R.ref_var.restype = ctypes.c_uint64
    return R


def _offsets(rng, n, w, h, rows):
    return np.array([int(rng.integers(0, rows - h)) * S + int(rng.integers(0, S - w)) for _ in range(n)], dtype=np.int64)


def _run(ctx, op, depth, w, h, dst, dstStride, s0, st0, s1, st1, jobs, p=(0, 0, 0, 0)):
    dD = ctx.to_device(dst)
    d0 = ctx.to_device(s0) if s0 is not None else None
    d1 = ctx.to_device(s1) if s1 is not None else None
    dJ = ctx.to_device(jobs)
    ctx.glue_dev(op, depth, w, h, dD, dstStride, d0, st0, d1, st1, dJ, len(jobs), *p)
    out = dD.download(dst.dtype)
    for b in (dD, d0, d1, dJ):
        if b is not None:
            b.free()
    return out


@pytest.mark.parametrize("depth", [8, 10])
def test_weight_pp_sp(ctx, depth):
    R = _ref(depth)
    dt = pdtype(depth)
    rows = 140
    src = pixel_buffers(depth, seed=7, size=S * rows)[0]
    ssrc = short_buffers(depth, seed=8, size=S * rows)[0]
    corr = 14 - depth
    for (w, h, w0, shift, offset) in [(64, 32, 61, 6 + corr, 3), (48, 17, 1, 0 + corr, -5), (16, 64, 127, 7 + corr, 20)]:
        rnd = (1 << (shift - 1)) & ~((1 << corr) - 1)
        off = np.array([5, 70 * S + 11], dtype=np.int64)            # disjoint blocks: the batched result is order-independent
        jobs = np.zeros(2, dtype=pkg.GLUE_JOB)
        jobs["dstOff"] = off; jobs["src0Off"] = off
        got = _run(ctx, pkg.GL_WEIGHT_PP, depth, w, h, np.zeros(S * rows, dtype=dt), S, src, S, None, 0, jobs, (w0, rnd, shift, offset))
        exp = np.zeros(S * rows, dtype=dt)
        for o in off:
            R.ref_weight_pp(vpo(src, int(o)), vpo(exp, int(o)), ssz(S), w, h, w0, rnd, shift, offset)
        assert np.array_equal(got, exp), ("weight_pp", depth, w, h)
        got = _run(ctx, pkg.GL_WEIGHT_SP, depth, w, h, np.zeros(S * rows, dtype=dt), S, ssrc, S, None, 0, jobs, (w0, rnd, shift, offset))
        exp = np.zeros(S * rows, dtype=dt)
        for o in off:
            R.ref_weight_sp(vpo(ssrc, int(o)), vpo(exp, int(o)), ssz(S), ssz(S), w, h, w0, rnd, shift, offset)
        assert np.array_equal(got, exp), ("weight_sp", depth, w, h)


@pytest.mark.parametrize("depth", [8, 10])
def test_avg_copy_resid(ctx, depth):
    R = _ref(depth)
    dt = pdtype(depth)
    rows = 140
    a, _, _ = pixel_buffers(depth, seed=21, size=S * rows)
    b, _, _ = pixel_buffers(depth, seed=22, size=S * rows)
    sa = short_buffers(depth, seed=23, size=S * rows)[0]
    sb = (short_buffers(depth, seed=24, size=S * rows)[0].astype(np.int32) * 8).clip(-32768, 32767).astype(np.int16)
    # non-overlapping tiling of 64x64 slots so that the batched result is order-independent
    slots = np.array([r * 64 * S + c * 32 for r in range(2) for c in range(1)], dtype=np.int64)
    for (w, h) in [(64, 64), (32, 24), (16, 4), (12, 16), (8, 8), (4, 4)]:
        part = R.ref_partition_from_sizes(w, h)
        jobs = np.zeros(len(slots), dtype=pkg.GLUE_JOB)
        jobs["dstOff"] = slots; jobs["src0Off"] = slots + 3; jobs["src1Off"] = slots + S + 1
        got = _run(ctx, pkg.GL_PIXELAVG_PP, depth, w, h, np.zeros(S * rows, dtype=dt), S, a, S, b, S, jobs)
        exp = np.zeros(S * rows, dtype=dt)
        for j in jobs:
            R.ref_pixelavg_pp(part, vpo(exp, int(j["dstOff"])), ssz(S), vpo(a, int(j["src0Off"])), ssz(S), vpo(b, int(j["src1Off"])), ssz(S))
        assert np.array_equal(got, exp), ("pixelavg", depth, w, h)
        got = _run(ctx, pkg.GL_ADDAVG, depth, w, h, np.zeros(S * rows, dtype=dt), S, sa, S, sb, S, jobs)
        exp = np.zeros(S * rows, dtype=dt)
        for j in jobs:
            R.ref_addavg(part, vpo(sa, int(j["src0Off"])), vpo(sb, int(j["src1Off"])), vpo(exp, int(j["dstOff"])), ssz(S), ssz(S), ssz(S))
        assert np.array_equal(got, exp), ("addAvg", depth, w, h)
    # numpy restatements for the pure copies / shifts / residuals (formulas of pixel.cpp:393-491, :759-840)
    w = h = 16
    jobs = np.zeros(2, dtype=pkg.GLUE_JOB)
    jobs["dstOff"] = [0, 20 * S + 5]; jobs["src0Off"] = [7, 40 * S + 9]; jobs["src1Off"] = [S + 2, 60 * S + 1]
    A2, B2, SA2 = a.reshape(rows, S).astype(np.int32), b.reshape(rows, S).astype(np.int32), sa.reshape(rows, S).astype(np.int32)

    def blk(P, o):
        return P[o // S:o // S + h, o % S:o % S + w]

    got = _run(ctx, pkg.GL_SUB_PS, depth, w, h, np.zeros(S * rows, dtype=np.int16), S, a, S, b, S, jobs).reshape(rows, S)
    for j in jobs:
        assert np.array_equal(blk(got, int(j["dstOff"])), blk(A2, int(j["src0Off"])) - blk(B2, int(j["src1Off"])))
    got = _run(ctx, pkg.GL_ADD_PS, depth, w, h, np.zeros(S * rows, dtype=dt), S, a, S, sa, S, jobs).reshape(rows, S)
    for j in jobs:
        assert np.array_equal(blk(got, int(j["dstOff"])), np.clip(blk(A2, int(j["src0Off"])) + blk(SA2, int(j["src1Off"])), 0, (1 << depth) - 1))
    for op, sh in [(pkg.GL_CPY1DTO2D_SHL, 3), (pkg.GL_CPY1DTO2D_SHR, 2)]:
        flat = sa[:2 * w * h].copy()
        jj = np.zeros(2, dtype=pkg.GLUE_JOB); jj["dstOff"] = [0, 30 * S + 4]; jj["src0Off"] = [0, w * h]
        got = _run(ctx, op, depth, w, h, np.zeros(S * rows, dtype=np.int16), S, flat, w, None, 0, jj, (sh, 0, 0, 0)).reshape(rows, S)
        for k, j in enumerate(jj):
            src = flat[k * w * h:(k + 1) * w * h].reshape(h, w).astype(np.int32)
            e = (src << sh) if op == pkg.GL_CPY1DTO2D_SHL else ((src + (1 << (sh - 1))) >> sh)
            assert np.array_equal(blk(got, int(j["dstOff"])), e.astype(np.int16))
    got = _run(ctx, pkg.GL_TRANSPOSE, depth, w, w, np.zeros(2 * w * w, dtype=dt), w, a, S, None, 0,
               np.array([(0, 11, 0), (w * w, 33 * S + 2, 0)], dtype=pkg.GLUE_JOB)).reshape(2, w, w)
    assert np.array_equal(got[0], blk(A2, 11).T) and np.array_equal(got[1], blk(A2, 33 * S + 2).T)


@pytest.mark.parametrize("depth", [8, 10])
def test_var_psy_copycnt(ctx, depth):
    R = _ref(depth)
    rows = 140
    rng = np.random.default_rng(31)
    for bi, (a, b) in enumerate(zip(pixel_buffers(depth, seed=41, size=S * rows), reversed(pixel_buffers(depth, seed=42, size=S * rows)))):
        for idx, N in enumerate([4, 8, 16, 32, 64]):
            n = 13
            offA, offB = _offsets(rng, n, N, N, rows), _offsets(rng, n, N, N, rows)
            dA, dB, dOA, dOB = ctx.to_device(a), ctx.to_device(b), ctx.to_device(offA), ctx.to_device(offB)
            dO = ctx.empty(n * 8)
            ctx.var_dev(depth, N, dA, S, dOA, n, dO)
            got = dO.download(np.uint64)
            exp = np.array([R.ref_var(idx, vpo(a, int(o)), ssz(S)) for o in offA], dtype=np.uint64)
            assert np.array_equal(got, exp), ("var", depth, N, bi)
            ctx.psy_cost_dev(depth, N, dA, S, dB, S, dOA, dOB, n, dO)
            got = dO.download(np.int32)[:n]
            exp = np.array([R.ref_psy_cost(idx, vpo(a, int(oa)), ssz(S), vpo(b, int(ob)), ssz(S)) for oa, ob in zip(offA, offB)], dtype=np.int32)
            assert np.array_equal(got, exp), ("psy", depth, N, bi)
            for x in (dA, dB, dOA, dOB, dO):
                x.free()
    resi = short_buffers(depth, seed=51, size=S * rows)[0]
    resi[rng.integers(0, len(resi), len(resi) // 2)] = 0
    for idx, N in enumerate([4, 8, 16, 32]):
        n = 11
        off = _offsets(rng, n, N, N, rows)
        dR, dOff, dC, dN = ctx.to_device(resi), ctx.to_device(off), ctx.empty(n * N * N * 2), ctx.empty(n * 4)
        ctx.copy_cnt_dev(N, dC, dR, S, dOff, n, dN)
        gc, gn = dC.download(np.int16).reshape(n, N * N), dN.download(np.uint32)
        for i, o in enumerate(off):
            e = np.empty(N * N, dtype=np.int16)
            cnt = R.ref_copy_cnt(idx, vpo(e, 0), vpo(resi, int(o)), ssz(S))
            assert np.array_equal(gc[i], e) and gn[i] == cnt
        for x in (dR, dOff, dC, dN):
            x.free()


def test_denoise_dct_batch(ctx):
    R = _ref(8)
    rng = np.random.default_rng(61)
    for numCoeff in (16, 64, 256, 1024):
        n = 7
        coef = rng.integers(-2000, 2001, n * numCoeff).astype(np.int16)
        offset = rng.integers(0, 300, numCoeff).astype(np.uint16)
        res0 = rng.integers(0, 1 << 20, numCoeff).astype(np.uint32)
        dC, dR, dF = ctx.to_device(coef), ctx.to_device(res0), ctx.to_device(offset)
        ctx.denoise_dct_dev(dC, dR, dF, numCoeff, n)
        gc, gr = dC.download(np.int16), dR.download(np.uint32)
        ec, er = coef.copy(), res0.copy()
        R.ref_denoise_dct.restype = None
        for i in range(n):                      # the reference processes TU after TU into the same accumulators
            R.ref_denoise_dct(vpo(ec, i * numCoeff), vpo(er, 0), vpo(offset, 0), numCoeff)
        assert np.array_equal(gc, ec) and np.array_equal(gr, er), numCoeff
        for x in (dC, dR, dF):
            x.free()


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("full", [False, True])
def test_lowpass_dct(ctx, depth, full):
    """lowPassDct8/16/32_c (lowpassdct.cpp:33-111) incl. the int16 wrap of the 2x2 sums on full-range input."""
    R = _ref(depth)
    for sizeIdx, N in [(1, 8), (2, 16), (3, 32)]:
        for buf in short_buffers(depth, seed=71 + sizeIdx, size=N * (N + 3) * 21 + 64, full_range=full):
            n, stride, blockStride = 19, N + 3, N * (N + 3) + 1
            dS, dD = ctx.to_device(buf), ctx.empty(n * N * N * 2)
            ctx.lowpass_dct_dev(sizeIdx, depth, dS, blockStride, stride, dD, n)
            got = dD.download(np.int16).reshape(n, N * N)
            exp = np.empty((n, N * N), dtype=np.int16)
            for i in range(n):
                R.ref_lowpass_dct(sizeIdx, vpo(buf, i * blockStride), vpo(exp, i * N * N), ssz(stride))
            assert np.array_equal(got, exp), (depth, full, N)
            dS.free(); dD.free()
