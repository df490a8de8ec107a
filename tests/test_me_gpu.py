"""GPU parity: batched motion estimation vs the reference's own MotionEstimate::motionEstimate
(oracle/_ref links the unmodified source/encoder/motion.cpp).  Every job must return the same
quarter-pel MV and the same cost -- DIA/HEX/UMH/STAR/FULL, subme 0..7 (luma), all inter PU shapes,
8- and 10-bit, candidates, tight ranges (border paths of the star/UMH code) and slice clamping."""
import importlib

import numpy as np
import pytest

from me_util import make_jobs, ref_me, ref_me_chroma, synth_chroma_pair, synth_pair
from util import LUMA_PU_SIZES

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu

INTER_SIZES = [s for s in LUMA_PU_SIZES if s != (4, 4)]


def _run(ctx, depth, method, subme, merange, qp, sizes, seed, n_per_size=10, maxSlices=1, mvp_span=24, motion=(5, -3), W=192, H=128):
    pad = 64 + merange + 16
    cur, ref, S, origin = synth_pair(W, H, pad, depth=depth, seed=seed, motion=motion)
    rng = np.random.default_rng(seed + 1)
    job = make_jobs(pkg, W, H, sizes, merange, rng, n_per_size=n_per_size, mvp_span=mvp_span)
    if maxSlices > 1:
        job["mvminY"] = np.maximum(job["mvminY"], -2)
        job["mvmaxY"] = np.minimum(job["mvmaxY"], 3)
    ex, ey, ec = ref_me(depth, cur, ref, S, origin, job, method, subme, merange, qp, maxSlices)
    item = cur.itemsize
    dC, dR, dJ = ctx.to_device(cur), ctx.to_device(ref), ctx.to_device(job)
    lam = pkg.lambda_for_qp(qp, depth)
    ctx.me_batch_dev(depth, dC.ptr + origin * item, S, dR.ptr + origin * item, S, dJ, len(job), 64, 64, method, subme, merange, lam, maxSlices)
    out = dJ.download(pkg.ME_JOB)
    bad = np.nonzero((out["outMvX"] != ex) | (out["outMvY"] != ey) | (out["outCost"] != ec))[0]
    msg = ""
    if len(bad):
        i = bad[0]
        msg = "job %d %s: got mv (%d,%d) cost %d, reference mv (%d,%d) cost %d; %d/%d differ" % (
            i, job[i], out["outMvX"][i], out["outMvY"][i], out["outCost"][i], ex[i], ey[i], ec[i], len(bad), len(job))
    for b in (dC, dR, dJ):
        b.free()
    assert not len(bad), msg


@pytest.mark.parametrize("method", [pkg.ME_DIA, pkg.ME_HEX, pkg.ME_UMH, pkg.ME_STAR])
@pytest.mark.parametrize("subme", [0, 1, 2])
def test_me_all_shapes_8bit(ctx, method, subme):
    _run(ctx, 8, method, subme, 57, 30, INTER_SIZES, seed=100 + method * 10 + subme, n_per_size=6)


@pytest.mark.parametrize("method", [pkg.ME_HEX, pkg.ME_STAR])
def test_me_10bit(ctx, method):
    _run(ctx, 10, method, 2, 57, 32, INTER_SIZES, seed=300 + method, n_per_size=4)


@pytest.mark.parametrize("subme", [3, 4, 5, 6, 7])
def test_me_subme_levels_luma(ctx, subme):
    _run(ctx, 8, pkg.ME_HEX, subme, 32, 27, [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (32, 24)], seed=400 + subme, n_per_size=8)


def _run_chroma(ctx, depth, csp, method, subme, merange, qp, sizes, seed, n_per_size=6, mv_only=False):
    W, H = 192, 128
    pad = 64 + merange + 16
    cur, ref, S, origin = synth_pair(W, H, pad, depth=depth, seed=seed)
    cCb, cCr, rCb, rCr, Sc, oc = synth_chroma_pair(W, H, pad, csp, depth=depth, seed=seed + 5)
    rng = np.random.default_rng(seed + 1)
    job = make_jobs(pkg, W, H, sizes, merange, rng, n_per_size=n_per_size)
    ex, ey, ec = ref_me_chroma(depth, csp, cur, ref, S, origin, (cCb, cCr), (rCb, rCr), Sc, oc, job, method, subme, merange, qp)
    item = cur.itemsize
    bufs = [ctx.to_device(a) for a in (cur, ref, cCb, cCr, rCb, rCr, job)]
    dC, dR, dCb, dCr, dRb, dRr, dJ = bufs
    lam = pkg.lambda_for_qp(qp, depth)
    ctx.me_batch_chroma_dev(depth, dC.ptr + origin * item, S, dR.ptr + origin * item, S, csp, dCb.ptr + oc * item, dCr.ptr + oc * item, Sc,
                            dRb.ptr + oc * item, dRr.ptr + oc * item, Sc, dJ, len(job), 64, 64, method, subme, merange, lam)
    out = dJ.download(pkg.ME_JOB)
    if mv_only:
        ec = out["outCost"]
    bad = np.nonzero((out["outMvX"] != ex) | (out["outMvY"] != ey) | (out["outCost"] != ec))[0]
    msg = ""
    if len(bad):
        i = bad[0]
        msg = "job %d %s: got mv (%d,%d) cost %d, reference mv (%d,%d) cost %d; %d/%d differ" % (
            i, job[i], out["outMvX"][i], out["outMvY"][i], out["outCost"][i], ex[i], ey[i], ec[i], len(bad), len(job))
    for b in bufs:
        b.free()
    assert not len(bad), msg
    return ec


@pytest.mark.parametrize("subme", [3, 4, 5, 7])
def test_me_chroma_satd_420(ctx, subme):
    """encode-style setSourcePU: subme > 2 adds the Cb/Cr SATD to every subpelCompare (motion.cpp:212,1601-1661);
    shapes whose chroma block is not a multiple of 4x4 (16x12, 12x16, 8x4, 16x4 ...) stay luma-only in the reference too."""
    _run_chroma(ctx, 8, 1, pkg.ME_HEX if subme != 5 else pkg.ME_STAR, subme, 32, 28, INTER_SIZES, seed=900 + subme)


def test_me_chroma_satd_below_subme3_is_luma_only(ctx):
    _run_chroma(ctx, 8, 1, pkg.ME_HEX, 2, 32, 28, [(16, 16), (32, 32)], seed=950)


@pytest.mark.parametrize("csp", [2, 3])
def test_me_chroma_satd_422_444(ctx, csp):
    _run_chroma(ctx, 8, csp, pkg.ME_UMH, 3, 24, 30, [(8, 8), (16, 16), (32, 16), (16, 32), (64, 64), (24, 32), (12, 16)], seed=960 + csp)


def test_me_chroma_satd_10bit(ctx):
    _run_chroma(ctx, 10, 1, pkg.ME_STAR, 3, 32, 32, [(8, 8), (16, 16), (32, 32), (64, 64), (32, 8), (8, 32)], seed=980)


def test_me_refine_mv(ctx):
    """MotionEstimate::refineMV (motion.cpp:606-737): only the MV is returned by the reference."""
    W, H, merange, depth, qp = 192, 128, 16, 8, 30
    pad = 64 + merange + 16
    cur, ref, S, origin = synth_pair(W, H, pad, depth=depth, seed=4242)
    job = make_jobs(pkg, W, H, INTER_SIZES, merange, np.random.default_rng(4243), n_per_size=6, with_cands=False)
    ex, ey, _ = ref_me(depth, cur, ref, S, origin, job, pkg.ME_REFINE, 2, merange, qp)
    dC, dR, dJ = ctx.to_device(cur), ctx.to_device(ref), ctx.to_device(job)
    ctx.me_batch_dev(depth, dC.ptr + origin, S, dR.ptr + origin, S, dJ, len(job), 64, 64, pkg.ME_REFINE, 2, merange, pkg.lambda_for_qp(qp, depth))
    out = dJ.download(pkg.ME_JOB)
    for b in (dC, dR, dJ):
        b.free()
    bad = np.nonzero((out["outMvX"] != ex) | (out["outMvY"] != ey))[0]
    assert not len(bad), "refineMV: %d/%d MVs differ, first job %s got (%d,%d) want (%d,%d)" % (
        len(bad), len(job), job[bad[0]], out["outMvX"][bad[0]], out["outMvY"][bad[0]], ex[bad[0]], ey[bad[0]])
    assert len(set(zip(ex.tolist(), ey.tolist()))) > 8        # the fixture moves the MVs


def test_me_refine_mv_chroma(ctx):
    ec = _run_chroma(ctx, 8, 1, pkg.ME_REFINE, 3, 16, 30, [(8, 8), (16, 16), (32, 32), (64, 64), (32, 16)], seed=4300, mv_only=True)


def test_me_large_motion_star_raster(ctx):
    """large true motion forces the star search past distance 5 -> raster refinement (motion.cpp:1169-1203)."""
    _run(ctx, 8, pkg.ME_STAR, 2, 48, 30, [(16, 16), (32, 32), (8, 8)], seed=500, n_per_size=10, motion=(23, -19), mvp_span=8)
    _run(ctx, 8, pkg.ME_UMH, 2, 48, 30, [(16, 16), (32, 32), (64, 64)], seed=501, n_per_size=10, motion=(-21, 17), mvp_span=8)


def test_me_tight_range_border_paths(ctx):
    """tiny search windows exercise the per-point border checks of STAR / UMH / CROSS."""
    for method in (pkg.ME_DIA, pkg.ME_HEX, pkg.ME_UMH, pkg.ME_STAR):
        _run(ctx, 8, method, 2, 3, 30, [(8, 8), (16, 16), (32, 32)], seed=600 + method, n_per_size=12)
        _run(ctx, 8, method, 1, 9, 22, [(16, 16), (64, 64)], seed=610 + method, n_per_size=8)


def test_me_full_search(ctx):
    _run(ctx, 8, pkg.ME_FULL, 2, 6, 30, [(8, 8), (16, 16), (32, 16)], seed=700, n_per_size=6)


def test_me_slice_clamp(ctx):
    _run(ctx, 8, pkg.ME_HEX, 2, 16, 30, [(16, 16), (32, 32)], seed=800, n_per_size=10, maxSlices=4)
    _run(ctx, 8, pkg.ME_STAR, 3, 16, 30, [(16, 16)], seed=801, n_per_size=10, maxSlices=2)


def test_me_qp_range(ctx):
    for qp in (0, 12, 37, 51):
        _run(ctx, 8, pkg.ME_HEX, 2, 24, qp, [(16, 16), (8, 8)], seed=900 + qp, n_per_size=8)


@pytest.mark.parametrize("depth,method,subme,merange", [(8, pkg.ME_HEX, 2, 57), (8, pkg.ME_STAR, 3, 24), (8, pkg.ME_DIA, 0, 57),
                                                        (8, pkg.ME_UMH, 2, 32), (10, pkg.ME_HEX, 2, 40)])
def test_me_frame_tma_window(ctx, depth, method, subme, merange):
    """frame form (TMA-staged search windows, one CTA per CTU x reference) == the reference's motionEstimate
    for every 2Nx2N PU of every CTU, two references, per-CTU predictors."""
    ctuCols, ctuRows, NREF = 3, 2, 2
    W, H = ctuCols * 64, ctuRows * 64
    mx, my = 144, 128                      # plane margins (>= merange + 16)
    pad = mx
    cur, ref0, S, origin = synth_pair(W, H, pad, depth=depth, seed=42 + depth, motion=(7, -4))
    _, ref1, _, _ = synth_pair(W, H, pad, depth=depth, seed=43 + depth, motion=(-9, 6))
    refs = [ref0, ref1]
    rowsTotal = H + 2 * pad
    rng = np.random.default_rng(5)
    mvp = rng.integers(-40, 41, (NREF, ctuCols * ctuRows, 2)).astype(np.int32)
    mvp[0, 0] = 0
    lam = pkg.lambda_for_qp(30, depth)
    item = cur.itemsize
    dC = ctx.to_device(cur)
    dR = [ctx.to_device(r) for r in refs]
    dMvp = ctx.to_device(mvp)
    per_level = [ctuCols * ctuRows * (1 << l) ** 2 for l in range(4)]
    nPU = sum(per_level)
    dOut = ctx.empty(NREF * nPU * 12)
    ctx.me_frame_dev(depth, dC.ptr + origin * item, S, [d.ptr + origin * item for d in dR], S, pad, pad, rowsTotal, ctuCols, ctuRows, 15,
                     dMvp, method, subme, merange, lam, dOut)
    got = dOut.download(np.int32).reshape(NREF, nPU, 3)
    for r in range(NREF):
        jobs = []
        for level in range(4):
            s = 64 >> level
            per = 1 << level
            for gy in range(ctuRows * per):
                for gx in range(ctuCols * per):
                    ctu = (gy // per) * ctuCols + (gx // per)
                    px, py = int(mvp[r, ctu, 0]), int(mvp[r, ctu, 1])
                    jobs.append((gx * s, gy * s, s, px, py))
        job = np.zeros(len(jobs), dtype=pkg.ME_JOB)
        for i, (x, y, s, px, py) in enumerate(jobs):
            job[i]["puX"], job[i]["puY"], job[i]["w"], job[i]["h"] = x, y, s, s
            job[i]["mvpX"], job[i]["mvpY"] = px, py
            job[i]["mvminX"], job[i]["mvminY"] = (px >> 2) - merange, (py >> 2) - merange
            job[i]["mvmaxX"], job[i]["mvmaxY"] = (px >> 2) + merange, (py >> 2) + merange
        ex, ey, ec = ref_me(depth, cur, refs[r], S, origin, job, method, subme, merange, 30)
        bad = np.nonzero((got[r, :, 0] != ex) | (got[r, :, 1] != ey) | (got[r, :, 2] != ec))[0]
        assert not len(bad), (r, int(bad[0]), jobs[bad[0]], got[r, bad[0]].tolist(), int(ex[bad[0]]), int(ey[bad[0]]), int(ec[bad[0]]), len(bad))
    for b in [dC, dMvp, dOut] + dR:
        b.free()


def test_me_frame_ctu_row_bands_equal_full_frame(ctx):
    """CTU-row sharding: a rank that owns CTU rows [y0, y1) calls the frame search with the plane origins
    advanced by y0 * 64 rows and marginY increased by the same amount; the union of the bands must equal the full-frame call."""
    depth, merange = 8, 24
    ctuCols, ctuRows, NREF = 3, 4, 2
    W, H = ctuCols * 64, ctuRows * 64
    pad = 96
    cur, ref0, S, origin = synth_pair(W, H, pad, depth=depth, seed=91, motion=(3, 5))
    _, ref1, _, _ = synth_pair(W, H, pad, depth=depth, seed=92, motion=(-6, -2))
    rowsTotal = H + 2 * pad
    lam = pkg.lambda_for_qp(30, depth)
    dC, dR = ctx.to_device(cur), [ctx.to_device(ref0), ctx.to_device(ref1)]

    def run(y0, rows):
        per_level = [ctuCols * rows * (1 << l) ** 2 for l in range(4)]
        dOut = ctx.empty(NREF * sum(per_level) * 12)
        sh = y0 * 64 * S
        ctx.me_frame_dev(depth, dC.ptr + origin + sh, S, [d.ptr + origin + sh for d in dR], S, pad, pad + y0 * 64, rowsTotal, ctuCols, rows, 15,
                         None, pkg.ME_HEX, 2, merange, lam, dOut)
        out = dOut.download(np.int32).reshape(NREF, sum(per_level), 3)
        dOut.free()
        return out, per_level

    full, pl_full = run(0, ctuRows)
    for (y0, rows) in ((0, 1), (1, 2), (3, 1)):
        band, pl = run(y0, rows)
        off_f = off_b = 0
        for level in range(4):
            per = 1 << level
            rowlen = ctuCols * per
            fsl = full[:, off_f + y0 * per * rowlen: off_f + (y0 + rows) * per * rowlen]
            assert np.array_equal(band[:, off_b:off_b + pl[level]], fsl), (y0, rows, level)
            off_f += pl_full[level]; off_b += pl[level]
    for b in [dC] + dR:
        b.free()


@pytest.mark.parametrize("method", [pkg.ME_HEX, pkg.ME_UMH, pkg.ME_STAR])
def test_me_frame_predictors_far_from_zero_on_static_content(ctx, method):
    """per-CTU predictors with |mvp >> 2| > merange on STATIC content -- the zero-MV candidate wins, so the search
    continues around (0, clampY(0)), far outside a window centred on the predictor.  Calls with predictors run on the general
    kernel (window test per block, plane fallback) and must still equal the reference's motionEstimate."""
    depth, merange, subme = 8, 16, 2
    ctuCols, ctuRows, NREF = 3, 2, 2
    W, H, pad = ctuCols * 64, ctuRows * 64, 256
    cur, _, S, origin = synth_pair(W, H, pad, depth=depth, seed=1717, motion=(0, 0), noise=0.0)
    refs = [cur.copy(), cur.copy()]                       # static: zero MV is the exact match
    refs[1][::7] ^= 1                                     # second reference: almost static
    rowsTotal = H + 2 * pad
    rng = np.random.default_rng(6)
    mvp = (rng.choice([-1, 1], (NREF, ctuCols * ctuRows, 2)) * rng.integers(4 * (merange + 2), 4 * (merange + 40), (NREF, ctuCols * ctuRows, 2))).astype(np.int32)
    mvp[0, 1] = [4 * (merange + 1), 0]                    # just past the range: UMH / hex overshoot at the window edge
    mvp[1, 2] = [0, -4 * merange]
    lam = pkg.lambda_for_qp(30, depth)
    dC = ctx.to_device(cur); dR = [ctx.to_device(r) for r in refs]; dMvp = ctx.to_device(mvp)
    per_level = [ctuCols * ctuRows * (1 << l) ** 2 for l in range(4)]
    nPU = sum(per_level)
    dOut = ctx.empty(NREF * nPU * 12)
    ctx.me_frame_dev(depth, dC.ptr + origin, S, [d.ptr + origin for d in dR], S, pad, pad, rowsTotal, ctuCols, ctuRows, 15, dMvp, method, subme, merange, lam, dOut)
    got = dOut.download(np.int32).reshape(NREF, nPU, 3)
    for r in range(NREF):
        jobs = []
        for level in range(4):
            s, per = 64 >> level, 1 << level
            for gy in range(ctuRows * per):
                for gx in range(ctuCols * per):
                    ctu = (gy // per) * ctuCols + (gx // per)
                    jobs.append((gx * s, gy * s, s, int(mvp[r, ctu, 0]), int(mvp[r, ctu, 1])))
        job = np.zeros(len(jobs), dtype=pkg.ME_JOB)
        for i, (x, y, s, px, py) in enumerate(jobs):
            job[i]["puX"], job[i]["puY"], job[i]["w"], job[i]["h"] = x, y, s, s
            job[i]["mvpX"], job[i]["mvpY"] = px, py
            job[i]["mvminX"], job[i]["mvminY"] = (px >> 2) - merange, (py >> 2) - merange
            job[i]["mvmaxX"], job[i]["mvmaxY"] = (px >> 2) + merange, (py >> 2) + merange
        ex, ey, ec = ref_me(depth, cur, refs[r], S, origin, job, method, subme, merange, 30)
        bad = np.nonzero((got[r, :, 0] != ex) | (got[r, :, 1] != ey) | (got[r, :, 2] != ec))[0]
        assert not len(bad), (r, int(bad[0]), jobs[bad[0]], got[r, bad[0]].tolist(), int(ex[bad[0]]), int(ey[bad[0]]), int(ec[bad[0]]), len(bad))
        assert (ex == 0).sum() > len(ex) // 4          # the zero MV does win for many PUs
    for b in [dC, dMvp, dOut] + dR:
        b.free()
