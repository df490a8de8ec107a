"""GPU parity for the --me sea row of SURVEY.md 8a (a-3): the SEA integral planes (FrameFilter::computeMEIntegral driving
integral_inith / integral_initv, framefilter.cpp:39-140,722-825), pu[].ads (pixel.cpp:121-165) and the X265_SEA branch of
MotionEstimate::motionEstimate (motion.cpp:1242-1395), all against the reference's own compiled objects (oracle/_ref)."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle
from me_util import REF_ME_JOB, make_jobs, synth_pair
from util import LUMA_PU_SIZES, pdtype, vp

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu

INTER_SIZES = [s for s in LUMA_PU_SIZES if s != (4, 4)]
PW, PH = pkg.SEA_PLANE_W, pkg.SEA_PLANE_H


def ref_planes(depth, plane, stride, rows, padX, padY, maxHeight):
    """the 12 planes as the reference builds them (allocated like the pixel plane; unwritten elements stay 0xAAAAAAAA)"""
    R = oracle.ref(depth)
    out = [np.full(stride * rows, 0xAAAAAAAA, dtype=np.uint32) for _ in range(12)]
    origin = padY * stride + padX
    arr = (ctypes.c_void_p * 12)(*[p.ctypes.data + origin * 4 for p in out])
    R.ref_sea_integrals(ctypes.c_void_p(plane.ctypes.data + origin * plane.itemsize), ctypes.c_ssize_t(stride), padX, padY, maxHeight, arr)
    return out


def gpu_planes(ctx, depth, dPlane, itemsize, stride, rows, padX, padY, maxHeight):
    origin = padY * stride + padX
    bufs = [ctx.empty(stride * rows * 4) for _ in range(12)]
    ctx.sea_integral_dev(depth, dPlane.ptr + origin * itemsize, stride, padX, padY, maxHeight, [b.ptr + origin * 4 for b in bufs])
    return bufs


@pytest.mark.parametrize("depth", [8, 10])
def test_sea_integral_planes(ctx, depth):
    W, H, padX, padY = 192, 128, 96, 80
    stride, rows = W + 2 * padX, H + 2 * padY
    rng = np.random.default_rng(5)
    plane = rng.integers(0, 1 << depth, stride * rows).astype(pdtype(depth))
    exp = ref_planes(depth, plane, stride, rows, padX, padY, H)
    dP = ctx.to_device(plane)
    bufs = gpu_planes(ctx, depth, dP, plane.itemsize, stride, rows, padX, padY, H)
    for k in range(12):
        got = bufs[k].download(np.uint32).reshape(rows, stride)
        e = exp[k].reshape(rows, stride)
        w, h = PW[k], PH[k]
        lo, hi = 1, rows - 1 - h
        assert np.array_equal(got[lo:hi + 1, :stride - w], e[lo:hi + 1, :stride - w]), "plane %d (%dx%d)" % (k, w, h)
        assert not got[0].any() and not e[0].any()                      # the memset row (framefilter.cpp:757)
        assert not got[hi + 1:].any() and not got[:, stride - w:].any()  # never finalised by the reference: defined as 0 here
    # the box-sum property, independent of the reference: element (r, c) = sum of the w x h pixels below/right of it
    k = 4
    got = bufs[k].download(np.uint32).reshape(rows, stride)
    p2 = plane.reshape(rows, stride).astype(np.int64)
    for (r, c) in ((1, 0), (17, 33), (rows - 1 - 16, stride - 17)):
        assert got[r, c] == p2[r:r + 16, c:c + 16].sum()
    for b in bufs + [dP]:
        b.free()


@pytest.mark.parametrize("depth", [8, 10])
def test_integral_row_primitives(ctx, depth):
    R = oracle.ref(depth)
    stride = 416
    rng = np.random.default_rng(9)
    for n in (4, 8, 12, 16, 24, 32):
        pix = rng.integers(0, 1 << depth, stride + 64).astype(pdtype(depth))
        sums = rng.integers(0, 1 << 32, stride * (n + 2), dtype=np.uint64).astype(np.uint32)
        e = sums.copy()
        R.ref_integral_inith(n, ctypes.c_void_p(e.ctypes.data + stride * 4), vp(pix), ctypes.c_ssize_t(stride))
        R.ref_integral_initv(n, vp(e), ctypes.c_ssize_t(stride))
        dS, dP = ctx.to_device(sums), ctx.to_device(pix)
        ctx.integral_inith_dev(depth, n, dS.ptr + stride * 4, dP, stride)
        ctx.integral_initv_dev(n, dS, stride)
        got = dS.download(np.uint32)
        assert np.array_equal(got, e), n
        dS.free(); dP.free()


def test_ads_vs_reference(ctx):
    R = oracle.ref(8)
    rng = np.random.default_rng(21)
    stride, width = 512, 116
    sums = rng.integers(0, 40000, stride * 40).astype(np.uint32)
    cost = rng.integers(0, 300, width).astype(np.uint16)
    dS, dC = ctx.to_device(sums), ctx.to_device(cost)
    for (w, h), kind in (((8, 8), 1), ((16, 12), 1), ((16, 8), 2), ((8, 16), 2), ((16, 16), 4), ((32, 32), 4), ((64, 16), 4), ((24, 32), 4)):
        part = LUMA_PU_SIZES.index((w, h))
        delta = 8 * stride if kind != 1 and (w, h) != (64, 16) else 8
        n = 16
        jobs = np.zeros(n, dtype=pkg.ADS_JOB)
        jobs["sumsOff"] = rng.integers(0, stride * 20, n)
        jobs["thresh"] = rng.integers(2000, 60000, n)
        jobs["encDC"] = rng.integers(0, 40000, (n, 4))
        dJ, dM, dN = ctx.to_device(jobs), ctx.empty(n * width * 2), ctx.empty(n * 4)
        ctx.ads_dev(kind, w >> 1, dS, delta, dC, width, dJ, n, dM, dN)
        gotN, gotM = dN.download(np.int32), dM.download(np.int16).reshape(n, width)
        for i in range(n):
            enc = jobs["encDC"][i].astype(np.int32).copy()
            mvs = np.zeros(width + 8, dtype=np.int16)
            cnt = R.ref_ads(part, vp(enc), ctypes.c_void_p(sums.ctypes.data + int(jobs["sumsOff"][i]) * 4), int(delta), vp(cost), vp(mvs), width,
                            int(jobs["thresh"][i]))
            assert cnt == gotN[i], (w, h, i)
            assert np.array_equal(mvs[:cnt], gotM[i, :cnt]), (w, h, i)
        for b in (dJ, dM, dN):
            b.free()
    dS.free(); dC.free()


def _run_sea(ctx, depth, subme, merange, qp, sizes, seed, n_per_size=6, mvp_span=24, motion=(5, -3), maxSlices=1):
    W, H = 192, 128
    pad = 64 + merange + 32
    cur, ref, S, origin = synth_pair(W, H, pad, depth=depth, seed=seed, motion=motion)
    rows = H + 2 * pad
    rng = np.random.default_rng(seed + 1)
    job = make_jobs(pkg, W, H, sizes, merange, rng, n_per_size=n_per_size, mvp_span=mvp_span)
    item = cur.itemsize
    # reference: its own planes, its own MotionEstimate
    R = oracle.ref(depth)
    rp = ref_planes(depth, ref, S, rows, pad, pad, H)
    rj = np.zeros(len(job), dtype=REF_ME_JOB)
    for f in ("puX", "puY", "w", "h", "mvminX", "mvminY", "mvmaxX", "mvmaxY", "mvpX", "mvpY", "numCand", "mvc"):
        rj[f] = job[f]
    arr = (ctypes.c_void_p * 12)(*[p.ctypes.data + origin * 4 for p in rp])
    R.ref_me_batch_sea(ctypes.c_void_p(cur.ctypes.data + origin * item), ctypes.c_ssize_t(S), ctypes.c_void_p(ref.ctypes.data + origin * item),
                       ctypes.c_ssize_t(S), arr, vp(rj), ctypes.c_int64(len(rj)), int(subme), int(merange), int(qp), int(maxSlices), 4)
    # backend: planes built on the device from the same reference plane
    dC, dR, dJ = ctx.to_device(cur), ctx.to_device(ref), ctx.to_device(job)
    bufs = gpu_planes(ctx, depth, dR, item, S, rows, pad, pad, H)
    dPtrs = ctx.to_device(np.array([b.ptr + origin * 4 for b in bufs], dtype=np.uint64))
    lam = pkg.lambda_for_qp(qp, depth)
    ctx.me_batch_sea_dev(depth, dC.ptr + origin * item, S, dR.ptr + origin * item, S, dPtrs, dJ, len(job), 64, 64, subme, merange, lam, maxSlices)
    out = dJ.download(pkg.ME_JOB)
    bad = np.nonzero((out["outMvX"] != rj["outMvX"]) | (out["outMvY"] != rj["outMvY"]) | (out["outCost"] != rj["outCost"]))[0]
    msg = ""
    if len(bad):
        i = bad[0]
        msg = "job %d %s: got mv (%d,%d) cost %d, reference mv (%d,%d) cost %d; %d/%d differ" % (
            i, job[i], out["outMvX"][i], out["outMvY"][i], out["outCost"][i], rj["outMvX"][i], rj["outMvY"][i], rj["outCost"][i], len(bad), len(job))
    for b in bufs + [dC, dR, dJ, dPtrs]:
        b.free()
    assert not len(bad), msg


@pytest.mark.parametrize("subme", [0, 2])
def test_me_sea_all_shapes_8bit(ctx, subme):
    _run_sea(ctx, 8, subme, 24, 30, INTER_SIZES, seed=700 + subme)


def test_me_sea_10bit(ctx):
    _run_sea(ctx, 10, 2, 16, 32, INTER_SIZES, seed=720, n_per_size=3)


def test_me_sea_wide_range_and_mvp(ctx):
    # merange 57 (medium), large predictors: the doubly-offset p_cost_mvx / p_cost_mvy tables of the reference matter
    _run_sea(ctx, 8, 1, 57, 26, [(8, 8), (16, 16), (32, 32), (64, 64), (32, 16), (16, 32), (64, 48), (24, 32)], seed=740, n_per_size=4, mvp_span=60,
             motion=(-9, 6))


def test_me_sea_needs_planes(ctx):
    job = np.zeros(1, dtype=pkg.ME_JOB)
    job["w"], job["h"] = 16, 16
    dJ = ctx.to_device(job)
    with pytest.raises(pkg.X265B200Error):
        ctx.me_batch_dev(8, dJ, 64, dJ, 64, dJ, 1, 64, 64, pkg.ME_SEA, 2, 16, 1.0)
    dJ.free()
