"""GPU parity: the lookahead path -- Lowres::init (frameInitLowres + border extension),
lowresIntraEstimate and estimateFrameCost/estimateCUCost -- against the reference's own Lookahead /
Lowres / CostEstimateGroup objects (oracle/_ref links the unmodified slicetype.cpp, lowres.cpp).
Compared bit-exactly: the 4 hpel planes incl. margins, intraCost/intraMode, per-CU MVs and MV costs of
both lists, lowresCosts (cost | listused<<14), rowSatds, costEst, costEstAq, intraMbs."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle
from util import pdtype

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu


def _bind(R):
    R.ref_la_create.restype = ctypes.c_void_p
    R.ref_la_frame_cost.restype = ctypes.c_int64
    R.ref_la_frame_cost_slices.restype = ctypes.c_int64
    R.ref_la_cost_est.restype = ctypes.c_int64
    for f in ("ref_la_mvs", "ref_la_mvcosts", "ref_la_lowres_costs", "ref_la_row_satds", "ref_la_intra_cost", "ref_la_intra_mode",
              "ref_la_fullres_buffer", "ref_la_lowres_buffer", "ref_la_inv_qscale"):
        getattr(R, f).restype = ctypes.c_void_p
    return R


def _arr(ptr, ctype, shape):
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=shape).copy()


def _frames(W, H, n, depth, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H + 96, W + 96)).astype(np.float32)
    k = np.ones(5) / 5
    for ax in (0, 1):
        base = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), ax, base)
    base = (base - base.min()) / (base.max() - base.min()) * 255
    out = []
    x = y = 40
    for i in range(n):
        x += int(rng.integers(-5, 6)); y += int(rng.integers(-4, 5))
        fr = base[y:y + H, x:x + W] + rng.normal(0, 2.5, (H, W))
        if i == n - 1:
            fr[H // 3: H // 3 + 40, W // 4: W // 4 + 60] = rng.integers(0, 256, (40, 60))   # an occlusion: intra-coded CUs
        out.append(np.ascontiguousarray(np.clip(np.rint(fr * (1 << (depth - 8))), 0, (1 << depth) - 1).astype(pdtype(depth))))
    return out


@pytest.mark.parametrize("depth,W,H,aq,slices", [(8, 320, 192, 0, 0), (8, 416, 240, 1, 0), (10, 320, 192, 1, 0),
                                                   # --lookahead-slices (coop slices, slicetype.cpp:3079-3107): 4 slices of 11 rows; 8 asked -> 6 of 10
                                                   # rows (+ a last one of 18); 2 slices of 34 rows (two 32-row bands per slice)
                                                   (8, 320, 704, 0, 4), (8, 256, 1088, 1, 8), (8, 256, 1088, 0, 2), (10, 320, 704, 1, 4)])
def test_lookahead_matches_reference(ctx, depth, W, H, aq, slices):
    R = oracle.ref(depth)
    assert R is not None, "oracle/_ref missing"
    _bind(R)
    NF, BF = 5, 3
    h = ctypes.c_void_p(R.ref_la_create(W, H, BF, aq))
    frames = _frames(W, H, NF, depth, seed=W + depth)
    for f in frames:
        R.ref_la_add_frame(h, ctypes.c_void_p(f.ctypes.data), ctypes.c_ssize_t(W))
    g = (ctypes.c_int64 * 11)()
    R.ref_la_geometry(h, g)
    lw, ll, ls, mx, my, wcu, hcu, fs, fmx, fmy, frows = [int(v) for v in g]
    ncu = wcu * hcu
    dt = pdtype(depth)
    px = np.dtype(dt).itemsize
    ct = ctypes.c_uint8 if depth == 8 else ctypes.c_uint16
    planesize = ls * (ll + 2 * my)
    padoff = ls * my + mx
    rng = np.random.default_rng(99)

    # ---- Lowres::init ---------------------------------------------------------------------------
    dPlanes = []          # device buffers [frame][4]
    plane_ptrs = np.zeros((NF, 4), dtype=np.int64)
    for i in range(NF):
        full = _arr(R.ref_la_fullres_buffer(h, i), ct, (fs * frows,))
        dFull = ctx.to_device(full)
        bufs = [ctx.to_device(np.zeros(planesize, dtype=dt)) for _ in range(4)]
        ctx.lowres_init_dev(depth, dFull.ptr + (fmy * fs + fmx) * px, fs, [b.ptr + padoff * px for b in bufs], ls, lw, ll, mx, my)
        for k in range(4):
            exp = _arr(R.ref_la_lowres_buffer(h, i, k), ct, (planesize,))
            got = bufs[k].download(dt)
            assert np.array_equal(got, exp), ("lowres plane", i, k)
            plane_ptrs[i, k] = bufs[k].ptr + padoff * px
        dPlanes.append(bufs)
        dFull.free()
    dPlanePtrs = ctx.to_device(plane_ptrs)

    # ---- AQ factors (injected on both sides) and lowresIntraEstimate ----------------------------------
    invq_ptrs = np.zeros(NF, dtype=np.int64)
    dInvQ = []
    for i in range(NF):
        if aq:
            q = rng.integers(128, 512, ncu).astype(np.int32)
            ctypes.memmove(R.ref_la_inv_qscale(h, i), q.ctypes.data, q.nbytes)
            dInvQ.append(ctx.to_device(q)); invq_ptrs[i] = dInvQ[-1].ptr
    dInvQPtrs = ctx.to_device(invq_ptrs)
    lam = pkg.lambda_for_qp(12 + 6 * (depth - 8), depth)
    intraPenalty = 5 * int(lam)
    dIntraCost, intra_ptrs = [], np.zeros(NF, dtype=np.int64)
    for i in range(NF):
        R.ref_la_intra(h, i)
        dIC, dIM, dLC, dRS, dSm = ctx.empty(ncu * 4), ctx.empty(ncu), ctx.empty(ncu * 2), ctx.empty(hcu * 4), ctx.empty(8)
        ctx.la_intra_dev(depth, plane_ptrs[i, 0], ls, wcu, hcu, invq_ptrs[i] if aq else None, intraPenalty, dIC, dIM, dLC, dRS, dSm)
        assert np.array_equal(dIC.download(np.int32), _arr(R.ref_la_intra_cost(h, i), ctypes.c_int32, (ncu,))), ("intraCost", i)
        assert np.array_equal(dIM.download(np.uint8), _arr(R.ref_la_intra_mode(h, i), ctypes.c_uint8, (ncu,))), ("intraMode", i)
        assert np.array_equal(dLC.download(np.uint16), _arr(R.ref_la_lowres_costs(h, i, 0, 0), ctypes.c_uint16, (ncu,)))
        assert np.array_equal(dRS.download(np.int32), _arr(R.ref_la_row_satds(h, i, 0, 0), ctypes.c_int32, (hcu,)))
        sums = dSm.download(np.int32)
        assert int(sums[0]) == R.ref_la_cost_est(h, i, 0, 0, 0)
        if aq:
            assert int(sums[1]) == R.ref_la_cost_est(h, i, 0, 0, 1)
        dIntraCost.append(dIC); intra_ptrs[i] = dIC.ptr
        for b in (dIM, dLC, dRS, dSm):
            b.free()
    dIntraPtrs = ctx.to_device(intra_ptrs)

    # ---- estimateFrameCost: waves of triples; MV slots model the lowresMvs[list][dist] cache ---------------
    def slot(b, lst, dist):
        return (b * 2 + lst) * (BF + 2) + dist
    nslots = NF * 2 * (BF + 2)
    dMv, dMvC = ctx.to_device(np.zeros(nslots * ncu * 2, dtype=np.int32)), ctx.to_device(np.zeros(nslots * ncu, dtype=np.int32))
    searched = set()
    waves = [[(0, 4, 2), (0, 4, 1), (0, 4, 3), (0, 4, 4)],      # B, B, B, P(b == p1)
             [(0, 2, 2), (2, 4, 3), (0, 1, 1), (1, 4, 2)]]      # reuse of cached lists mixed with new searches
    for wave in waves:
        tr = np.zeros(len(wave), dtype=pkg.LA_TRIPLE)
        for t, (p0, p1, b) in enumerate(wave):
            tr[t]["b"], tr[t]["p0"], tr[t]["p1"] = b, p0, p1
            for lst, dist in ((0, b - p0), (1, p1 - b)):
                key = (b, lst, dist)
                tr[t]["mvSlot"][lst] = slot(*key)
                need = key not in searched and (lst == 0 or p1 > b)
                tr[t]["doSearch"][lst] = int(need)
                if need:
                    searched.add(key)
            if slices:
                R.ref_la_frame_cost_slices(h, p0, p1, b, slices)      # the reference's processTasks runs every coop slice
            else:
                R.ref_la_frame_cost(h, p0, p1, b, 0)
        dLC, dRS, dSm = ctx.empty(len(wave) * ncu * 2), ctx.empty(len(wave) * hcu * 4), ctx.empty(len(wave) * 16)
        ctx.la_estimate_dev(depth, dPlanePtrs, ls, wcu, hcu, tr, dMv, dMvC, dIntraPtrs, dInvQPtrs if aq else None, dLC, dRS, dSm, lam,
                            lookaheadSlices=slices)
        lc = dLC.download(np.uint16).reshape(len(wave), ncu)
        rs = dRS.download(np.int32).reshape(len(wave), hcu)
        sm = dSm.download(np.int32).reshape(len(wave), 4)
        mv = dMv.download(np.int32).reshape(nslots, ncu, 2)
        mvc = dMvC.download(np.int32).reshape(nslots, ncu)
        for t, (p0, p1, b) in enumerate(wave):
            for lst, dist in ((0, b - p0), (1, p1 - b)):
                if lst == 1 and p1 == b:
                    continue
                emv = _arr(R.ref_la_mvs(h, b, lst, dist), ctypes.c_int32, (ncu, 2))
                emc = _arr(R.ref_la_mvcosts(h, b, lst, dist), ctypes.c_int32, (ncu,))
                bad = np.nonzero((mv[slot(b, lst, dist)] != emv).any(axis=1) | (mvc[slot(b, lst, dist)] != emc))[0]
                assert not len(bad), ("MV/cost", (p0, p1, b), lst, int(bad[0]), mv[slot(b, lst, dist)][bad[0]].tolist(), emv[bad[0]].tolist(),
                                      int(mvc[slot(b, lst, dist)][bad[0]]), int(emc[bad[0]]), len(bad))
            assert np.array_equal(lc[t], _arr(R.ref_la_lowres_costs(h, b, b - p0, p1 - b), ctypes.c_uint16, (ncu,))), ("lowresCosts", wave[t])
            assert np.array_equal(rs[t], _arr(R.ref_la_row_satds(h, b, b - p0, p1 - b), ctypes.c_int32, (hcu,))), ("rowSatds", wave[t])
            score = int(sm[t][0])
            if b != p1:
                score = score * 100 // 130                      # slicetype.cpp:3203-3204, bFrameBias = 0
            assert score == R.ref_la_cost_est(h, b, b - p0, p1 - b, 0), ("costEst", wave[t])
            if aq:
                assert int(sm[t][1]) == R.ref_la_cost_est(h, b, b - p0, p1 - b, 1), ("costEstAq", wave[t])
            if b == p1:
                assert int(sm[t][2]) == R.ref_la_intra_mbs(h, b, b - p0), ("intraMbs", wave[t])
        for bb in (dLC, dRS, dSm):
            bb.free()
    if slices:
        assert R.ref_la_num_coop_slices(h) == {4: 4, 8: 6, 2: 2}[slices]
    R.ref_la_destroy(h)


def _fade_frames(W, H, n, depth, seed, gains, offsets):
    """a moving scene whose brightness follows gains / offsets per frame (what weightp is for)"""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H + 96, W + 96)).astype(np.float32)
    k = np.ones(5) / 5
    for ax in (0, 1):
        base = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), ax, base)
    base = (base - base.min()) / (base.max() - base.min()) * 200 + 20
    out = []
    x = y = 40
    for i in range(n):
        x += int(rng.integers(-4, 5)); y += int(rng.integers(-3, 4))
        fr = base[y:y + H, x:x + W] * gains[i] + offsets[i] + rng.normal(0, 1.5, (H, W))
        out.append(np.ascontiguousarray(np.clip(np.rint(fr * (1 << (depth - 8))), 0, (1 << depth) - 1).astype(pdtype(depth))))
    return out


@pytest.mark.parametrize("depth,W,H,slices", [(8, 320, 192, 0), (10, 320, 192, 0), (8, 320, 704, 4)])
def test_lookahead_weightp_matches_reference(ctx, depth, W, H, slices):
    """weightp in the lookahead (slicetype.cpp:860-961 weightsAnalyse, hook :3136-3138, use :3222): the reference's Lookahead with
    bEnableWeightedPred = 1 runs weightsAnalyse before every list-0 search; ours runs x265b200_la_weights_analyse_dev for the
    wave's (fenc, ref) pairs and x265b200_la_estimate_dev reads the device-side decision.  Frames 0..2 fade (pure gain: the reference takes its means from full-resolution sums over lowres areas, so offsets never survive),
    frames 3..4 do not: weighted pairs, pairs that terminate early and pairs measured but rejected all occur."""
    R = oracle.ref(depth)
    assert R is not None, "oracle/_ref missing"
    _bind(R)
    R.ref_la_weighted_buffer.restype = ctypes.c_void_p
    NF, BF = 5, 3
    h = ctypes.c_void_p(R.ref_la_create(W, H, BF, 0))
    R.ref_la_set_weightp(h, 1)
    frames = _fade_frames(W, H, NF, depth, seed=11 * W + depth, gains=[0.55, 0.7, 0.85, 1.0, 1.0], offsets=[0, 0, 0, 0, 0])
    for f in frames:
        R.ref_la_add_frame(h, ctypes.c_void_p(f.ctypes.data), ctypes.c_ssize_t(W))
    g = (ctypes.c_int64 * 11)()
    R.ref_la_geometry(h, g)
    lw, ll, ls, mx, my, wcu, hcu, fs, fmx, fmy, frows = [int(v) for v in g]
    ncu = wcu * hcu
    dt = pdtype(depth)
    px = np.dtype(dt).itemsize
    ct = ctypes.c_uint8 if depth == 8 else ctypes.c_uint16
    planesize, padoff = ls * (ll + 2 * my), ls * my + mx
    paddedLines = planesize // ls
    # wp_sum[0] / wp_ssd[0] as calcAdaptiveQuantFrame leaves them (slicetype.cpp:672-674), on both sides
    stats = []
    for i, f in enumerate(frames):
        v = f.astype(np.int64)
        sm, ssd = int(v.sum()), int((v * v).sum())
        ssd = ssd - (sm * sm + (W * H) // 2) // (W * H)
        stats.append((sm, ssd))
        R.ref_la_set_wp_stats(h, i, ctypes.c_uint64(sm), ctypes.c_uint64(ssd))

    bufs_all, plane_ptrs = [], np.zeros((NF + 4, 4), dtype=np.int64)
    for i in range(NF):
        full = _arr(R.ref_la_fullres_buffer(h, i), ct, (fs * frows,))
        dFull = ctx.to_device(full)
        bufs = [ctx.to_device(np.zeros(planesize, dtype=dt)) for _ in range(4)]
        ctx.lowres_init_dev(depth, dFull.ptr + (fmy * fs + fmx) * px, fs, [b.ptr + padoff * px for b in bufs], ls, lw, ll, mx, my)
        plane_ptrs[i] = [b.ptr + padoff * px for b in bufs]
        bufs_all.append(bufs)
        dFull.free()
    # one set of wbuffer planes per pair of a wave (the reference reuses ONE per thread, sequentially)
    wbufs = [[ctx.to_device(np.full(planesize, 7, dtype=dt)) for _ in range(4)] for _ in range(4)]
    for j in range(4):
        plane_ptrs[NF + j] = [b.ptr + padoff * px for b in wbufs[j]]
    dPlanePtrs = ctx.to_device(plane_ptrs)

    lam = pkg.lambda_for_qp(12 + 6 * (depth - 8), depth)
    intraPenalty = 5 * int(lam)
    dIntraCost, intra_ptrs = [], np.zeros(NF, dtype=np.int64)
    for i in range(NF):
        R.ref_la_intra(h, i)
        dIC, dIM, dLC, dRS, dSm = ctx.empty(ncu * 4), ctx.empty(ncu), ctx.empty(ncu * 2), ctx.empty(hcu * 4), ctx.empty(8)
        ctx.la_intra_dev(depth, plane_ptrs[i, 0], ls, wcu, hcu, None, intraPenalty, dIC, dIM, dLC, dRS, dSm)
        dIntraCost.append(dIC); intra_ptrs[i] = dIC.ptr
    dIntraPtrs = ctx.to_device(intra_ptrs)

    def slot(b, lst, dist):
        return (b * 2 + lst) * (BF + 2) + dist
    nslots = NF * 2 * (BF + 2)
    dMv, dMvC = ctx.to_device(np.zeros(nslots * ncu * 2, dtype=np.int32)), ctx.to_device(np.zeros(nslots * ncu, dtype=np.int32))
    searched = set()
    waves = [[(0, 4, 2), (0, 4, 4), (3, 4, 4), (1, 3, 2)],      # B with a fading list 0, P across the fade, P without fade, B
             [(0, 2, 2), (2, 4, 3), (0, 1, 1), (1, 4, 2)]]
    seen = {"weighted": 0, "early": 0, "rejected": 0}
    for wave in waves:
        tr = np.zeros(len(wave), dtype=pkg.LA_TRIPLE)
        jobs = np.zeros(len(wave), dtype=pkg.LA_WEIGHT_JOB)
        njobs = 0
        expect = []
        for t, (p0, p1, b) in enumerate(wave):
            tr[t]["b"], tr[t]["p0"], tr[t]["p1"] = b, p0, p1
            for lst, dist in ((0, b - p0), (1, p1 - b)):
                key = (b, lst, dist)
                tr[t]["mvSlot"][lst] = slot(*key)
                need = key not in searched and (lst == 0 or p1 > b)
                tr[t]["doSearch"][lst] = int(need)
                if need:
                    searched.add(key)
            if tr[t]["doSearch"][0]:
                J = jobs[njobs]
                J["fencPlane0"], J["intraCost"] = plane_ptrs[b, 0], intra_ptrs[b]
                J["refBuffer"] = [bb.ptr for bb in bufs_all[p0]]
                J["weighted"] = [bb.ptr for bb in wbufs[njobs]]
                J["fencSum"], J["fencSsd"], J["refSum"], J["refSsd"] = stats[b][0], stats[b][1], stats[p0][0], stats[p0][1]
                tr[t]["weightIdx0"], tr[t]["weightPlanes0"] = njobs + 1, NF + njobs
                njobs += 1
            if slices:
                R.ref_la_frame_cost_slices(h, p0, p1, b, slices)
            else:
                R.ref_la_frame_cost(h, p0, p1, b, 0)
            if tr[t]["doSearch"][0]:
                w = int(R.ref_la_is_weighted(h, b, b - p0))
                planes = [_arr(R.ref_la_weighted_buffer(h, k), ct, (planesize,)) for k in range(4)] if w else None
                expect.append((w, planes))
        dW = ctx.to_device(np.zeros(max(njobs, 1), dtype=pkg.LA_WEIGHT))
        ctx.la_weights_analyse_dev(depth, jobs[:njobs], ls, paddedLines, padoff, lw, ll, dW)
        dLC, dRS, dSm = ctx.empty(len(wave) * ncu * 2), ctx.empty(len(wave) * hcu * 4), ctx.empty(len(wave) * 16)
        ctx.la_estimate_dev(depth, dPlanePtrs, ls, wcu, hcu, tr, dMv, dMvC, dIntraPtrs, None, dLC, dRS, dSm, lam,
                            lookaheadSlices=slices, dWeights=dW)
        got = dW.download(pkg.LA_WEIGHT)
        for j, (w, planes) in enumerate(expect):
            assert int(got[j]["isWeighted"]) == w, ("isWeighted", wave, j, got[j])
            if w:
                seen["weighted"] += 1
                for k in range(4):
                    assert np.array_equal(wbufs[j][k].download(dt), planes[k]), ("weighted plane", wave, j, k)
            elif int(got[j]["origscore"]) == 0:
                seen["early"] += 1
            else:
                seen["rejected"] += 1
        lc = dLC.download(np.uint16).reshape(len(wave), ncu)
        rs = dRS.download(np.int32).reshape(len(wave), hcu)
        sm = dSm.download(np.int32).reshape(len(wave), 4)
        mv = dMv.download(np.int32).reshape(nslots, ncu, 2)
        mvc = dMvC.download(np.int32).reshape(nslots, ncu)
        for t, (p0, p1, b) in enumerate(wave):
            for lst, dist in ((0, b - p0), (1, p1 - b)):
                if lst == 1 and p1 == b:
                    continue
                emv = _arr(R.ref_la_mvs(h, b, lst, dist), ctypes.c_int32, (ncu, 2))
                emc = _arr(R.ref_la_mvcosts(h, b, lst, dist), ctypes.c_int32, (ncu,))
                bad = np.nonzero((mv[slot(b, lst, dist)] != emv).any(axis=1) | (mvc[slot(b, lst, dist)] != emc))[0]
                assert not len(bad), ("MV/cost", (p0, p1, b), lst, int(bad[0]), len(bad))
            assert np.array_equal(lc[t], _arr(R.ref_la_lowres_costs(h, b, b - p0, p1 - b), ctypes.c_uint16, (ncu,))), ("lowresCosts", wave[t])
            assert np.array_equal(rs[t], _arr(R.ref_la_row_satds(h, b, b - p0, p1 - b), ctypes.c_int32, (hcu,))), ("rowSatds", wave[t])
            score = int(sm[t][0])
            if b != p1:
                score = score * 100 // 130
            assert score == R.ref_la_cost_est(h, b, b - p0, p1 - b, 0), ("costEst", wave[t])
        for bb in (dLC, dRS, dSm, dW):
            bb.free()
    assert seen["weighted"] >= 2 and seen["early"] >= 1, seen
    R.ref_la_destroy(h)


@pytest.mark.parametrize("depth,W,H", [(8, 320, 192), (8, 424, 240), (10, 320, 192)])
def test_lookahead_hme_matches_reference(ctx, depth, W, H):
    """--hme: quarter-resolution planes (frameInitLowerRes + borders), the level-0 search (lowerResMvs / lowerResMvCosts) and the
    8x8 level with hmeSearchMethod[1] (UMH) / hmeRange[1] and the extra doubled candidate -- all against the reference's own
    Lookahead with bEnableHME = 1.  424 wide makes m_4x4Width differ from m_8x8Width / 2 (the candidate index pitch quirk)."""
    R = oracle.ref(depth)
    assert R is not None, "oracle/_ref missing"
    _bind(R)
    R.ref_la_create_hme.restype = ctypes.c_void_p
    for f in ("ref_la_lower_buffer", "ref_la_lower_mvs", "ref_la_lower_mvcosts"):
        getattr(R, f).restype = ctypes.c_void_p
    NF, BF = 4, 2
    h = ctypes.c_void_p(R.ref_la_create_hme(W, H, BF, 0))
    frames = _frames(W, H, NF, depth, seed=7 * W + depth)
    for f in frames:
        R.ref_la_add_frame(h, ctypes.c_void_p(f.ctypes.data), ctypes.c_ssize_t(W))
    g = (ctypes.c_int64 * 11)()
    R.ref_la_geometry(h, g)
    lw, ll, ls, mx, my, wcu, hcu, fs, fmx, fmy, frows = [int(v) for v in g]
    hg = (ctypes.c_int64 * 8)()
    R.ref_la_hme_geometry(h, hg)
    w4, h4, m0, m1, r0, r1, lowOff, lowPlane = [int(v) for v in hg]
    ncu, n4 = wcu * hcu, w4 * h4
    dt = pdtype(depth)
    px = np.dtype(dt).itemsize
    ct = ctypes.c_uint8 if depth == 8 else ctypes.c_uint16
    planesize, padoff = ls * (ll + 2 * my), ls * my + mx
    ls2, lw2, ll2, mx2, my2 = ls // 2, lw // 2, ll // 2, mx // 2, my // 2
    rows2 = ll2 + 2 * my2 + 8
    plane2size, pad2 = ls2 * rows2, ls2 * my2 + mx2

    plane_ptrs = np.zeros((NF, 4), dtype=np.int64)
    lower_ptrs = np.zeros((NF, 4), dtype=np.int64)
    keep = []
    for i in range(NF):
        full = _arr(R.ref_la_fullres_buffer(h, i), ct, (fs * frows,))
        dFull = ctx.to_device(full)
        bufs = [ctx.to_device(np.zeros(planesize, dtype=dt)) for _ in range(4)]
        ctx.lowres_init_dev(depth, dFull.ptr + (fmy * fs + fmx) * px, fs, [b.ptr + padoff * px for b in bufs], ls, lw, ll, mx, my)
        low = [ctx.to_device(np.zeros(plane2size, dtype=dt)) for _ in range(4)]
        # Lowres::init, lowres.cpp:304-313: frameInitLowerRes(lowresPlane[0], ..., lumaStride, lumaStride/2, width/2, lines/2) + extendPicBorder
        ctx.lowres_init_dev(depth, bufs[0].ptr + padoff * px, ls, [b.ptr + pad2 * px for b in low], ls2, lw2, ll2, mx2, my2)
        ref_low = _arr(R.ref_la_lower_buffer(h, i), ct, (4 * lowPlane,))
        for k in range(4):
            plane_ptrs[i, k] = bufs[k].ptr + padoff * px
            lower_ptrs[i, k] = low[k].ptr + pad2 * px
            got = low[k].download(dt).reshape(rows2, ls2)
            for r in (-my2, -1, 0, ll2 // 2, ll2 - 1, ll2 + my2 - 1):
                e = ref_low[k * lowPlane + lowOff + r * ls2 - mx2: k * lowPlane + lowOff + r * ls2 + lw2 + mx2]
                assert np.array_equal(got[my2 + r, :lw2 + 2 * mx2], e), ("lower-res plane", i, k, r)
        keep += bufs + low
        dFull.free()
    dPlanePtrs, dLowerPtrs = ctx.to_device(plane_ptrs), ctx.to_device(lower_ptrs)

    lam = pkg.lambda_for_qp(12 + 6 * (depth - 8), depth)
    intraPenalty = 5 * int(lam)
    intra_ptrs = np.zeros(NF, dtype=np.int64)
    for i in range(NF):
        R.ref_la_intra(h, i)
        dIC, dIM, dLC, dRS, dSm = ctx.empty(ncu * 4), ctx.empty(ncu), ctx.empty(ncu * 2), ctx.empty(hcu * 4), ctx.empty(8)
        ctx.la_intra_dev(depth, plane_ptrs[i, 0], ls, wcu, hcu, None, intraPenalty, dIC, dIM, dLC, dRS, dSm)
        intra_ptrs[i] = dIC.ptr
        keep += [dIC, dIM, dLC, dRS, dSm]
    dIntraPtrs = ctx.to_device(intra_ptrs)

    def slot(b, lst, dist):
        return (b * 2 + lst) * (BF + 2) + dist
    nslots = NF * 2 * (BF + 2)
    dMv, dMvC = ctx.to_device(np.zeros(nslots * ncu * 2, dtype=np.int32)), ctx.to_device(np.zeros(nslots * ncu, dtype=np.int32))
    dMv4, dMvC4 = ctx.to_device(np.zeros(nslots * n4 * 2, dtype=np.int32)), ctx.to_device(np.zeros(nslots * n4, dtype=np.int32))
    hme = pkg.LA_HME(dLowerPtrs.ptr, ls2, w4, h4, dMv4.ptr, dMvC4.ptr, (ctypes.c_int32 * 2)(m0, m1), (ctypes.c_int32 * 2)(r0, r1))
    assert (m0, m1, r0, r1) == (pkg.ME_HEX, pkg.ME_UMH, 16, 32)
    searched = set()
    for wave in ([(0, 3, 1), (0, 3, 2), (0, 3, 3)], [(0, 2, 1), (1, 3, 2), (2, 3, 3)]):
        tr = np.zeros(len(wave), dtype=pkg.LA_TRIPLE)
        for t, (p0, p1, b) in enumerate(wave):
            tr[t]["b"], tr[t]["p0"], tr[t]["p1"] = b, p0, p1
            for lst, dist in ((0, b - p0), (1, p1 - b)):
                key = (b, lst, dist)
                tr[t]["mvSlot"][lst] = slot(*key)
                need = key not in searched and (lst == 0 or p1 > b)
                tr[t]["doSearch"][lst] = int(need)
                if need:
                    searched.add(key)
            R.ref_la_frame_cost(h, p0, p1, b, 0)
        dLC, dRS, dSm = ctx.empty(len(wave) * ncu * 2), ctx.empty(len(wave) * hcu * 4), ctx.empty(len(wave) * 16)
        ctx.la_estimate_hme_dev(depth, dPlanePtrs, ls, wcu, hcu, hme, tr, dMv, dMvC, dIntraPtrs, None, dLC, dRS, dSm, lam)
        lc = dLC.download(np.uint16).reshape(len(wave), ncu)
        sm = dSm.download(np.int32).reshape(len(wave), 4)
        mv, mvc = dMv.download(np.int32).reshape(nslots, ncu, 2), dMvC.download(np.int32).reshape(nslots, ncu)
        mv4, mvc4 = dMv4.download(np.int32).reshape(nslots, n4, 2), dMvC4.download(np.int32).reshape(nslots, n4)
        for t, (p0, p1, b) in enumerate(wave):
            for lst, dist in ((0, b - p0), (1, p1 - b)):
                if lst == 1 and p1 == b:
                    continue
                s_ = slot(b, lst, dist)
                e4 = _arr(R.ref_la_lower_mvs(h, b, lst, dist), ctypes.c_int32, (n4, 2))
                c4 = _arr(R.ref_la_lower_mvcosts(h, b, lst, dist), ctypes.c_int32, (n4,))
                bad = np.nonzero((mv4[s_] != e4).any(axis=1) | (mvc4[s_] != c4))[0]
                assert not len(bad), ("level-0 MV/cost", wave[t], lst, int(bad[0]), mv4[s_][bad[0]].tolist(), e4[bad[0]].tolist(), len(bad))
                emv = _arr(R.ref_la_mvs(h, b, lst, dist), ctypes.c_int32, (ncu, 2))
                emc = _arr(R.ref_la_mvcosts(h, b, lst, dist), ctypes.c_int32, (ncu,))
                bad = np.nonzero((mv[s_] != emv).any(axis=1) | (mvc[s_] != emc))[0]
                assert not len(bad), ("MV/cost", wave[t], lst, int(bad[0]), mv[s_][bad[0]].tolist(), emv[bad[0]].tolist(), len(bad))
            assert np.array_equal(lc[t], _arr(R.ref_la_lowres_costs(h, b, b - p0, p1 - b), ctypes.c_uint16, (ncu,))), ("lowresCosts", wave[t])
            score = int(sm[t][0])
            if b != p1:
                score = score * 100 // 130
            assert score == R.ref_la_cost_est(h, b, b - p0, p1 - b, 0), ("costEst", wave[t])
        for bb in (dLC, dRS, dSm):
            bb.free()
    for b in keep + [dPlanePtrs, dLowerPtrs, dIntraPtrs, dMv, dMvC, dMv4, dMvC4]:
        b.free()
    R.ref_la_destroy(h)


@pytest.mark.parametrize("depth,W,H", [(8, 416, 240), (10, 320, 192)])
def test_cutree_propagate_matches_reference(ctx, depth, W, H):
    """Lookahead::estimateCUPropagate (slicetype.cpp:2641-2747) for a B triple and a P triple, referenced and not: the reference's own
    function runs on its Lowres objects, the device driver on copies of the same arrays; the saturating scatter must agree entry for
    entry (incl. CUs whose MVs land partly outside the frame and entries that saturate at 65535)."""
    R = oracle.ref(depth)
    assert R is not None, "oracle/_ref missing"
    _bind(R)
    R.ref_la_propagate_cost.restype = ctypes.c_void_p
    R.ref_la_fps_factor.restype = ctypes.c_double
    NF, BF = 5, 3
    h = ctypes.c_void_p(R.ref_la_create(W, H, BF, 1))
    for f in _frames(W, H, NF, depth, seed=3 * W + depth):
        R.ref_la_add_frame(h, ctypes.c_void_p(f.ctypes.data), ctypes.c_ssize_t(W))
    g = (ctypes.c_int64 * 11)()
    R.ref_la_geometry(h, g)
    wcu, hcu = int(g[5]), int(g[6])
    ncu = wcu * hcu
    rng = np.random.default_rng(5)
    for i in range(NF):
        q = rng.integers(128, 512, ncu).astype(np.int32)
        ctypes.memmove(R.ref_la_inv_qscale(h, i), q.ctypes.data, q.nbytes)
        R.ref_la_intra(h, i)
    avg_dur = 1 / 25.0
    for (p0, p1, b, referenced) in ((0, 4, 2, 1), (0, 4, 4, 1), (2, 4, 3, 0), (0, 2, 1, 1)):
        R.ref_la_frame_cost(h, p0, p1, b, 0)
        # random, partly near-saturated source / target costs on the reference's arrays
        for i in {p0, p1, b}:
            pc = rng.integers(0, 65536 if i != b else 3000, ncu).astype(np.uint16)
            ctypes.memmove(R.ref_la_propagate_cost(h, i), pc.ctypes.data, pc.nbytes)
        before = {i: _arr(R.ref_la_propagate_cost(h, i), ctypes.c_uint16, (ncu,)) for i in {p0, p1, b}}
        intra = _arr(R.ref_la_intra_cost(h, b), ctypes.c_int32, (ncu,))
        lc = _arr(R.ref_la_lowres_costs(h, b, b - p0, p1 - b), ctypes.c_uint16, (ncu,))
        invq = _arr(R.ref_la_inv_qscale(h, b), ctypes.c_int32, (ncu,))
        mv0 = _arr(R.ref_la_mvs(h, b, 0, b - p0), ctypes.c_int32, (ncu, 2))
        mv1 = _arr(R.ref_la_mvs(h, b, 1, p1 - b), ctypes.c_int32, (ncu, 2)) if p1 > b else None
        fps = R.ref_la_fps_factor(h, ctypes.c_double(avg_dur))
        R.ref_la_cutree_propagate(h, p0, p1, b, referenced, ctypes.c_double(avg_dur))
        want0 = _arr(R.ref_la_propagate_cost(h, p0), ctypes.c_uint16, (ncu,))
        want1 = _arr(R.ref_la_propagate_cost(h, p1), ctypes.c_uint16, (ncu,))
        d = {k: ctx.to_device(v) for k, v in dict(pb=before[b], intra=intra, lc=lc, invq=invq, mv0=mv0, r0=before[p0], r1=before[p1]).items()}
        dmv1 = ctx.to_device(mv1) if mv1 is not None else None
        ctx.cutree_propagate_dev(wcu, hcu, d["pb"] if referenced else None, d["intra"], d["lc"], d["invq"], d["mv0"], dmv1, d["r0"],
                                 d["r1"] if p1 != b else None, 32, fps)
        got0 = d["r0"].download(np.uint16)
        assert np.array_equal(got0, want0), ((p0, p1, b), int(np.count_nonzero(got0 != want0)))
        if p1 != b:
            got1 = d["r1"].download(np.uint16)
            assert np.array_equal(got1, want1), ((p0, p1, b), "list 1", int(np.count_nonzero(got1 != want1)))
        assert (want0 != before[p0]).any()
        for v in list(d.values()) + ([dmv1] if dmv1 is not None else []):
            v.free()
    R.ref_la_destroy(h)
