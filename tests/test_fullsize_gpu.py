"""GPU parity at BASELINE.json's full size (2160p 8-bit, the bench.py frame): size-independent properties and cross-checks
that do not need the CPU oracle to finish a whole frame --
  * SAD pyramid: level consistency (every 16/32/64 SAD is the sum of its four children) + the 8x8 level against numpy;
  * me_frame (TMA-staged frame search) against me_batch (one warp per PU from global memory) on every PU of sampled CTU
    rows -- two independent kernels -- and against the reference's MotionEstimate on one CTU row;
  * fused residual pipeline over the whole frame: numSig == count_nonzero(levels), sse == sum (fenc - recon)^2, recon == pred
    where numSig == 0, and bit-exact agreement with the reference chain on one CTU row;
  * SEA integral planes: box-sum property at random positions of the 2160p plane;
  * fused intra modes over every 32x32 block: DC/planar/angular slots against the reference on a sample of blocks."""
import ctypes
import importlib
import os
import sys

import numpy as np
import pytest

import oracle
from me_util import ref_me
from util import vp, vpo, ssz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (frame generator and geometry of the benchmark)

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu

W, HH, S, PAD, ROWS = bench.W, bench.CTU_ROWS * bench.CTU, bench.STRIDE, bench.PAD, bench.ROWS
ORIGIN = PAD * S + PAD


@pytest.fixture(scope="module")
def frames():
    return bench.synth_frames(3, seed=77)


def test_sad_pyramid_2160p(ctx, frames):
    cur, ref = frames[2], frames[1]
    dC, dR = ctx.to_device(cur), ctx.to_device(ref)
    dPtr = ctx.to_device(np.array([dR.ptr + ORIGIN], dtype=np.uint64))
    n = {s: (W // s) * (HH // s) for s in (8, 16, 32, 64)}
    out = {s: ctx.empty(n[s] * 4) for s in n}
    ctx.sad_pyramid_dev(8, dC.ptr + ORIGIN, S, dPtr, 1, S, bench.CTU_COLS, bench.CTU_ROWS, None, out[8], out[16], out[32], out[64])
    g = {s: out[s].download(np.int32).reshape(HH // s, W // s) for s in n}
    a = cur[PAD:PAD + HH, PAD:PAD + W].astype(np.int32)
    b = ref[PAD:PAD + HH, PAD:PAD + W].astype(np.int32)
    e8 = np.abs(a - b).reshape(HH // 8, 8, W // 8, 8).sum(axis=(1, 3))
    assert np.array_equal(g[8], e8)
    for s in (16, 32, 64):
        child = g[s // 2]
        assert np.array_equal(g[s], child.reshape(HH // s, 2, W // s, 2).sum(axis=(1, 3))), s
    for bfr in [dC, dR, dPtr] + list(out.values()):
        bfr.free()


def test_me_frame_2160p_vs_me_batch_and_reference(ctx, frames):
    cur, refs = frames[2], [frames[1], frames[0]]
    dC = ctx.to_device(cur)
    dR = [ctx.to_device(r) for r in refs]
    lam = pkg.lambda_for_qp(bench.QP, 8)
    per_level = [bench.CTU_COLS * bench.CTU_ROWS * (1 << l) ** 2 for l in range(4)]
    nPU = sum(per_level)
    dOut = ctx.empty(2 * nPU * 12)
    ctx.me_frame_dev(8, dC.ptr + ORIGIN, S, [d.ptr + ORIGIN for d in dR], S, PAD, PAD, ROWS, bench.CTU_COLS, bench.CTU_ROWS, 15, None,
                     pkg.ME_HEX, bench.SUBME, bench.MERANGE, lam, dOut)
    got = dOut.download(np.int32).reshape(2, nPU, 3)
    # every PU of three CTU rows (top, middle, bottom) through the independent one-warp-per-PU kernel
    level_off = np.concatenate([[0], np.cumsum(per_level)])
    for r in range(2):
        jobs, where = [], []
        for level in range(4):
            s, per = 64 >> level, 1 << level
            for cy in (0, bench.CTU_ROWS // 2, bench.CTU_ROWS - 1):
                for gy in range(cy * per, (cy + 1) * per):
                    for gx in range(bench.CTU_COLS * per):
                        jobs.append((gx * s, gy * s, s))
                        where.append(level_off[level] + gy * bench.CTU_COLS * per + gx)
        job = np.zeros(len(jobs), dtype=pkg.ME_JOB)
        for i, (x, y, s) in enumerate(jobs):
            job[i]["puX"], job[i]["puY"], job[i]["w"], job[i]["h"] = x, y, s, s
            job[i]["mvminX"] = job[i]["mvminY"] = -bench.MERANGE
            job[i]["mvmaxX"] = job[i]["mvmaxY"] = bench.MERANGE
        dJ = ctx.to_device(job)
        ctx.me_batch_dev(8, dC.ptr + ORIGIN, S, dR[r].ptr + ORIGIN, S, dJ, len(job), 64, 64, pkg.ME_HEX, bench.SUBME, bench.MERANGE, lam)
        o = dJ.download(pkg.ME_JOB)
        w = np.array(where)
        assert np.array_equal(got[r, w, 0], o["outMvX"]) and np.array_equal(got[r, w, 1], o["outMvY"]) and np.array_equal(got[r, w, 2], o["outCost"]), r
        dJ.free()
        # the reference's own MotionEstimate on the middle CTU row (all four levels)
        mid = [i for i, (x, y, s) in enumerate(jobs) if y // 64 == bench.CTU_ROWS // 2]
        sub = job[mid]
        ex, ey, ec = ref_me(8, cur.ravel(), refs[r].ravel(), S, ORIGIN, sub, pkg.ME_HEX, bench.SUBME, bench.MERANGE, bench.QP, threads=8)
        assert np.array_equal(o["outMvX"][mid], ex) and np.array_equal(o["outMvY"][mid], ey) and np.array_equal(o["outCost"][mid], ec), r
    for bfr in [dC, dOut] + dR:
        bfr.free()


def test_tu_pipeline_2160p_properties(ctx, frames):
    R = oracle.ref(8)
    cur, ref = frames[2], frames[1]
    fenc = np.ascontiguousarray(cur[PAD:PAD + HH, PAD:PAD + W])
    pred = np.ascontiguousarray(ref[PAD:PAD + HH, PAD:PAD + W])
    dF, dP, dR = ctx.to_device(fenc), ctx.to_device(pred), ctx.empty(W * HH)
    dQ = ctx.to_device(np.full(1024, 26214, dtype=np.int32))
    for idx, N in bench.TU_SIZES:
        bx, by = W // N, HH // N
        n = bx * by
        qbits, add = bench.quant_params(N)
        dC, dN, dS = ctx.empty(W * HH * 2), ctx.empty(n * 4), ctx.empty(n * 8)
        ctx.tu_pipeline_dev(idx, 8, 0, dF, W, dP, W, dR, W, bx, by, dQ, qbits, add, None, 40 << 5, 9, dC, dN, dS)
        coef = dC.download(np.int16).reshape(n, N * N)
        ns, sse = dN.download(np.uint32), dS.download(np.uint64)
        rec = dR.download(np.uint8).reshape(HH, W)
        assert np.array_equal(ns, np.count_nonzero(coef, axis=1)), N
        d = fenc.astype(np.int64) - rec.astype(np.int64)
        assert np.array_equal(sse, (d * d).reshape(by, N, bx, N).sum(axis=(1, 3)).ravel().astype(np.uint64)), N
        zero = (ns == 0).reshape(by, bx)
        same = (rec == pred).reshape(by, N, bx, N).all(axis=(1, 3))
        assert same[zero].all(), N
        # one CTU row of TUs against the reference chain
        row0 = (bench.CTU_ROWS // 2) * 64
        rby = 64 // N
        e_rec = np.zeros((64, W), dtype=np.uint8)
        e_coef = np.zeros(rby * bx * N * N, dtype=np.int16)
        e_ns, e_sse = np.zeros(rby * bx, dtype=np.uint32), np.zeros(rby * bx, dtype=np.uint64)
        qt = np.full(1024, 26214, dtype=np.int32)
        R.ref_tu_pipeline(idx, 0, vpo(fenc.ravel(), row0 * W), ssz(W), vpo(pred.ravel(), row0 * W), ssz(W), vp(e_rec), ssz(W), bx, rby,
                          vp(qt), qbits, add, None, 40 << 5, 9, vp(e_coef), vp(e_ns), vp(e_sse), 8)
        first = (row0 // N) * bx
        assert np.array_equal(coef[first:first + rby * bx].ravel(), e_coef), N
        assert np.array_equal(ns[first:first + rby * bx], e_ns) and np.array_equal(sse[first:first + rby * bx], e_sse), N
        assert np.array_equal(rec[row0:row0 + 64], e_rec), N
        for bfr in (dC, dN, dS):
            bfr.free()
    for bfr in (dF, dP, dR, dQ):
        bfr.free()


def test_sea_integral_2160p_box_sums(ctx, frames):
    ref = frames[1]
    dR = ctx.to_device(ref)
    planes = [ctx.empty(S * ROWS * 4) for _ in range(12)]
    ctx.sea_integral_dev(8, dR.ptr + ORIGIN, S, PAD, PAD, ROWS - 2 * PAD, [b.ptr + ORIGIN * 4 for b in planes])
    ii = np.zeros((ROWS + 1, S + 1), dtype=np.int64)
    ii[1:, 1:] = ref.astype(np.int64).cumsum(axis=0).cumsum(axis=1)
    rng = np.random.default_rng(3)
    for k in (0, 3, 5, 9, 11):
        w, h = pkg.SEA_PLANE_W[k], pkg.SEA_PLANE_H[k]
        got = planes[k].download(np.uint32).reshape(ROWS, S)
        rr = rng.integers(1, ROWS - 1 - h + 1, 4000)
        cc = rng.integers(0, S - w, 4000)
        exp = ii[rr + h, cc + w] - ii[rr, cc + w] - ii[rr + h, cc] + ii[rr, cc]
        assert np.array_equal(got[rr, cc].astype(np.int64), exp), k
        assert not got[0].any() and not got[ROWS - h:].any()
    for bfr in planes + [dR]:
        bfr.free()


def test_intra_modes_2160p_sample(ctx, frames):
    R = oracle.ref(8)
    N, log2N = 32, 5
    nbr = bench.neighbour_arrays(frames[0].ravel(), N)
    n, L = len(nbr), 4 * N + 1
    dN, dD = ctx.to_device(nbr), ctx.empty(n * 35 * N * N)
    ctx.intra_modes_dev(8, log2N, dN, dD, 0, n)
    got = dD.download(np.uint8).reshape(n, 35, N, N)
    flat = nbr.ravel()
    rng = np.random.default_rng(11)
    for i in rng.integers(0, n, 24):
        filt = np.zeros(L, dtype=np.uint8)
        R.ref_intra_filter(log2N - 2, vpo(flat, int(i) * L), vp(filt))
        for mode in (0, 1, 2, 9, 10, 11, 18, 26, 27, 34):
            use_filt = (mode == 0) or (mode >= 2 and min(abs(mode - 26), abs(mode - 10)) > 0)
            e = np.empty((N, N), dtype=np.uint8)
            if use_filt:
                R.ref_intra_pred(log2N - 2, mode, vp(e), ssz(N), vp(filt), 0)
            else:
                R.ref_intra_pred(log2N - 2, mode, vp(e), ssz(N), vpo(flat, int(i) * L), 0)
            if 2 <= mode < 18:
                e = e.T
            assert np.array_equal(got[i, mode], e), (int(i), mode)
    dN.free(); dD.free()
