"""GPU parity: interpolation filters (all 8 kinds, luma 8-tap for the 25 PU sizes, chroma 4-tap for
the 4:2:0 shapes) and intra prediction (35 modes x 4 sizes, filter, all-angles) vs the reference C
primitives.  Shapes follow source/test/ipfilterharness.cpp:59-90 (random strides, src offset by 3
rows, poisoned outputs compared over the whole buffer) and intrapredharness.cpp."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle
from util import LUMA_PU_SIZES, pdtype, vpo, ssz

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu


def _ref(depth):
    R = oracle.ref(depth)
    assert R is not None, "oracle/_ref missing"
    return R


def _pix(depth, n, seed, mode="rand"):
    rng = np.random.default_rng(seed)
    pmax = (1 << depth) - 1
    if mode == "rand":
        return rng.integers(0, pmax + 1, n, dtype=np.int64).astype(pdtype(depth))
    return np.full(n, 0 if mode == "min" else pmax, dtype=pdtype(depth))


def _shorts(depth, n, seed):
    rng = np.random.default_rng(seed)
    # the int16 intermediates the ps filters can produce
    return rng.integers(-8192 - 2000, 8192 + 2000, n, dtype=np.int64).astype(np.int16)


KINDS = [("hpp", 0), ("hps", 1), ("vpp", 2), ("vps", 3), ("vsp", 4), ("vss", 5), ("hvpp", 6), ("p2s", 7)]


def _run_interp(ctx, R, depth, csp, part, w, h, taps, kind, kidx, srcStride, dstStride, seed):
    src_is_short = kind in ("vsp", "vss")
    dst_is_short = kind in ("hps", "vps", "vss", "p2s")
    rows_out = h + (taps - 1 if kind == "hps" else 0)
    src_len = (h + 16) * srcStride + 32
    src = _shorts(depth, src_len, seed) if src_is_short else _pix(depth, src_len, seed)
    src_off = 8 * srcStride + 8
    dst_len = (rows_out + 2) * dstStride + 8
    dt = np.int16 if dst_is_short else pdtype(depth)
    fracs = range(1, 4) if taps == 8 else range(1, 8)
    for cx in fracs:
        cy = (cx % 3) + 1 if taps == 8 else cx
        isRowExt = 1 if kind == "hps" else 0
        poison = np.full(dst_len, 0x4d4d if dt != np.uint8 else 0xcd, dtype=dt)
        dS, dD = ctx.to_device(src), ctx.to_device(poison)
        job = np.zeros(1, dtype=pkg.INTERP_JOB)
        job["srcOff"], job["dstOff"], job["idxX"], job["idxY"] = src_off, dstStride + 2, cx, cy
        dJ = ctx.to_device(job)
        ctx.interp_dev(kidx, taps, depth, w, h, dS, srcStride, dD, dstStride, dJ, 1, isRowExt)
        got = dD.download(dt)
        exp = poison.copy()
        R.ref_interp(kidx, csp, part, vpo(src, src_off), ssz(srcStride), vpo(exp, dstStride + 2), ssz(dstStride), cx,
                     cy if kind == "hvpp" else isRowExt)
        assert np.array_equal(got, exp), (depth, csp, kind, w, h, cx, cy, srcStride, dstStride)
        for b in (dS, dD, dJ):
            b.free()


@pytest.mark.parametrize("depth", [8, 10])
def test_luma_interp_all_sizes(ctx, depth):
    R = _ref(depth)
    rng = np.random.default_rng(101)
    for part, (w, h) in enumerate(LUMA_PU_SIZES):
        for kind, kidx in KINDS:
            srcStride = w + 16 + int(rng.integers(0, 9))
            dstStride = w + int(rng.integers(0, 7)) if kind != "hvpp" else w + 3
            _run_interp(ctx, R, depth, -1, part, w, h, 8, kind, kidx, srcStride, dstStride, seed=part * 8 + kidx)


@pytest.mark.parametrize("depth", [8, 10])
def test_chroma420_interp(ctx, depth):
    R = _ref(depth)
    rng = np.random.default_rng(202)
    for part, (lw, lh) in enumerate(LUMA_PU_SIZES):
        w, h = lw // 2, lh // 2
        for kind, kidx in KINDS:
            if kind == "hvpp":
                continue
            if not R.ref_interp_available(kidx, 1, part):
                continue
            srcStride = w + 12 + int(rng.integers(0, 9))
            dstStride = w + int(rng.integers(0, 7))
            _run_interp(ctx, R, depth, 1, part, w, h, 4, kind, kidx, srcStride, dstStride, seed=1000 + part * 8 + kidx)


def test_interp_batch_many_jobs(ctx):
    """many jobs in one launch, different fractions per job (the subpel-ME access pattern)."""
    R = _ref(8)
    rng = np.random.default_rng(7)
    W, H, S = 256, 128, 288
    plane = rng.integers(0, 256, (H + 32) * S, dtype=np.int64).astype(np.uint8)
    w = h = 16
    n = 200
    job = np.zeros(n, dtype=pkg.INTERP_JOB)
    job["srcOff"] = (16 + rng.integers(0, H - h, n)) * S + 16 + rng.integers(0, W - w - 16, n)
    job["dstOff"] = np.arange(n) * w * h
    job["idxX"] = rng.integers(1, 4, n)
    job["idxY"] = rng.integers(1, 4, n)
    dP, dJ, dD = ctx.to_device(plane), ctx.to_device(job), ctx.empty(n * w * h)
    ctx.interp_dev(pkg.IP_HVPP, 8, 8, w, h, dP, S, dD, w, dJ, n)
    got = dD.download(np.uint8).reshape(n, w * h)
    part = LUMA_PU_SIZES.index((w, h))
    for i in range(n):
        e = np.empty(w * h, dtype=np.uint8)
        R.ref_interp(6, -1, part, vpo(plane, job["srcOff"][i]), ssz(S), vpo(e, 0), ssz(w), int(job["idxX"][i]), int(job["idxY"][i]))
        assert np.array_equal(got[i], e), i
    for b in (dP, dJ, dD):
        b.free()


def test_interp_multi_segments_equal_single_launches(ctx):
    """x265b200_interp_multi_dev: segments of different kinds / block sizes in shared launches (and one the packed-word kernel does not
    take, forwarded) give what one x265b200_interp_dev call per segment gives."""
    rng = np.random.default_rng(17)
    W, H, S = 256, 128, 288
    plane = rng.integers(0, 256, (H + 32) * S, dtype=np.int64).astype(np.uint8)
    dP = ctx.to_device(plane)
    specs = [(pkg.IP_HPP, 8, 8, 150), (pkg.IP_VPP, 16, 16, 90), (pkg.IP_HVPP, 32, 32, 40), (pkg.IP_HVPP, 8, 8, 0), (pkg.IP_HPP, 64, 64, 7),
             (pkg.IP_HVPP, 12, 16, 33), (pkg.IP_HPS, 16, 16, 20), (pkg.IP_VPP, 8, 32, 25)] + [(pkg.IP_HPP, 16, 8, 11)] * 12
    segs, singles, keep = [], [], []
    for kind, w, h, n in specs:
        job = np.zeros(max(n, 1), dtype=pkg.INTERP_JOB)[:n]
        job["srcOff"] = (16 + rng.integers(0, H - h, n)) * S + 16 + rng.integers(0, W - w - 16, n)
        job["dstOff"] = np.arange(n) * w * h
        job["idxX"] = rng.integers(1, 4, n)
        job["idxY"] = rng.integers(1, 4, n)
        px = 2 if kind == pkg.IP_HPS else 1
        dJ = ctx.to_device(job) if n else None
        dA, dB = ctx.empty(max(n, 1) * w * h * px), ctx.empty(max(n, 1) * w * h * px)
        segs.append((kind, w, h, dP, S, dA, w, dJ, n))
        singles.append((kind, w, h, dB, dJ, n, px))
        keep += [dA, dB] + ([dJ] if n else [])
    ctx.interp_multi_dev(8, 8, segs)
    for (kind, w, h, dB, dJ, n, px), sg in zip(singles, segs):
        if n:
            ctx.interp_dev(kind, 8, 8, w, h, dP, S, dB, w, dJ, n)
            assert np.array_equal(sg[5].download(np.uint8), dB.download(np.uint8)), (kind, w, h, n)
    for b in keep + [dP]:
        b.free()


@pytest.mark.parametrize("depth", [8, 10])
def test_intra_pred_all_modes(ctx, depth):
    R = _ref(depth)
    for log2N in (2, 3, 4, 5):
        N = 1 << log2N
        L = 4 * N + 1
        for fill in ("rand", "min", "max"):
            nbr = _pix(depth, L * 4 + 8, 300 + log2N, fill)
            jobs = []
            for mode in range(35):
                for bf in (0, 1):
                    jobs.append((3 + (mode % 3), len(jobs) * (N * (N + 2)) + 1, mode, bf))
            job = np.zeros(len(jobs), dtype=pkg.INTRA_JOB)
            for i, (so, do, m, bf) in enumerate(jobs):
                job[i] = (so, do, m, bf)
            stride = N + 2
            dst_len = len(jobs) * N * stride + 8
            dt = pdtype(depth)
            poison = np.full(dst_len, 0xcd, dtype=dt)
            dN, dJ, dD = ctx.to_device(nbr), ctx.to_device(job), ctx.to_device(poison)
            ctx.intra_pred_dev(depth, log2N, dN, dD, stride, dJ, len(jobs))
            got = dD.download(dt)
            exp = poison.copy()
            for (so, do, m, bf) in jobs:
                R.ref_intra_pred(log2N - 2, m, vpo(exp, do), ssz(stride), vpo(nbr, so), bf)
            assert np.array_equal(got, exp), (depth, log2N, fill)
            for b in (dN, dJ, dD):
                b.free()


@pytest.mark.parametrize("depth", [8, 10])
def test_intra_filter_and_allangs(ctx, depth):
    R = _ref(depth)
    import ctypes as C
    for log2N in (2, 3, 4, 5):
        N = 1 << log2N
        L = 4 * N + 1
        n = 9
        nbr = _pix(depth, L * n, 400 + log2N)
        dt = pdtype(depth)
        dS, dF = ctx.to_device(nbr), ctx.empty(nbr.nbytes)
        ctx.intra_filter_dev(depth, log2N, dS, dF, n)
        got = dF.download(dt)
        exp = np.zeros_like(nbr)
        for i in range(n):
            R.ref_intra_filter(log2N - 2, vpo(nbr, i * L), vpo(exp, i * L))
        assert np.array_equal(got, exp), (depth, log2N)
        # all-angles: compare with the per-mode reference primitive (the C all_angs is nulled at
        # runtime, primitives.cpp:257; its definition intrapred.cpp:206-234 = per-mode predictions with
        # horizontal modes stored transposed)
        dA = ctx.empty(n * 33 * N * N * nbr.itemsize)
        for bLuma in (0, 1):
            ctx.intra_allangs_dev(depth, log2N, dS, dF, dA, bLuma, n)
            ga = dA.download(dt).reshape(n, 33, N, N)
            thr = {2: 99, 3: 7, 4: 1, 5: 0}[log2N]
            for i in range(n):
                for mode in range(2, 35):
                    dist = min(abs(mode - 26), abs(mode - 10))
                    srcarr = exp if dist > thr else nbr
                    e = np.empty((N, N), dtype=dt)
                    R.ref_intra_pred(log2N - 2, mode, vpo(e, 0), ssz(N), vpo(srcarr, i * L), bLuma)
                    if mode < 18:
                        e = e.T
                    assert np.array_equal(ga[i, mode - 2], e), (depth, log2N, i, mode, bLuma)
        for b in (dS, dF, dA):
            b.free()


@pytest.mark.parametrize("depth", [8, 10])
def test_intra_modes_fused(ctx, depth):
    """x265b200_intra_modes_dev = intra_filter + DC + planar + intra_pred_allangs per block (the prediction half of
    Search::estIntraPredQT, search.cpp:1358-1400) against the reference's per-mode primitives; also from an unaligned
    destination (generic kernel) and for block counts that leave a partial group."""
    R = _ref(depth)
    dt = pdtype(depth)
    for log2N in (2, 3, 4, 5):
        N = 1 << log2N
        L = 4 * N + 1
        for n, misalign in ((13, 0), (5, 4)):      # +4 pixels: 4-byte aligned only -> the generic (one CTA per block) kernel
            nbr = _pix(depth, L * n, 500 + log2N + n)
            filt = np.zeros_like(nbr)
            for i in range(n):
                R.ref_intra_filter(log2N - 2, vpo(nbr, i * L), vpo(filt, i * L))
            bLuma = int(N <= 16)
            dS = ctx.to_device(nbr)
            dD = ctx.empty((n * 35 * N * N + 8) * nbr.itemsize)
            ctx.intra_modes_dev(depth, log2N, dS, dD.ptr + misalign * nbr.itemsize, bLuma, n)
            got = dD.download(dt)[misalign:misalign + n * 35 * N * N].reshape(n, 35, N, N)
            thr = {2: 99, 3: 7, 4: 1, 5: 0}[log2N]
            for i in range(n):
                for mode in range(35):
                    if mode == 0:
                        src, bf = (filt if N >= 8 else nbr), 0
                    elif mode == 1:
                        src, bf = nbr, bLuma
                    else:
                        src, bf = (filt if min(abs(mode - 26), abs(mode - 10)) > thr else nbr), bLuma
                    e = np.empty((N, N), dtype=dt)
                    R.ref_intra_pred(log2N - 2, mode, vpo(e, 0), ssz(N), vpo(src, i * L), bf)
                    if 2 <= mode < 18:
                        e = e.T
                    assert np.array_equal(got[i, mode], e), (depth, log2N, n, i, mode)
            dS.free(); dD.free()
