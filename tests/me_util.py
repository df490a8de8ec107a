"""Helpers for the motion-estimation parity tests: synthetic frame pairs and job builders."""
import ctypes

import numpy as np

import oracle
from util import pdtype, vp

REF_ME_JOB = np.dtype([("puX", np.int32), ("puY", np.int32), ("w", np.int32), ("h", np.int32),
                       ("mvminX", np.int32), ("mvminY", np.int32), ("mvmaxX", np.int32), ("mvmaxY", np.int32),
                       ("mvpX", np.int32), ("mvpY", np.int32), ("numCand", np.int32), ("mvc", np.int32, (8, 2)),
                       ("outMvX", np.int32), ("outMvY", np.int32), ("outCost", np.int32)])


def synth_pair(W, H, pad, depth=8, seed=1234, motion=(5, -3), noise=3.0):
    """Two padded luma planes: band-limited noise, the second a translated + noisy copy of the first
    (BASELINE.md 2.3 generator shape).  Returns (cur, ref, stride, origin_offset)."""
    rng = np.random.default_rng(seed)
    S, Hh = W + 2 * pad, H + 2 * pad
    big = rng.uniform(0, 255, (Hh + 64, S + 64))
    k = np.ones(5) / 5.0
    for ax in (0, 1):
        big = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), ax, big)
    big = (big - big.min()) / (big.max() - big.min()) * 255.0
    dx, dy = motion
    ref = big[32:32 + Hh, 32:32 + S]
    cur = big[32 + dy:32 + dy + Hh, 32 + dx:32 + dx + S] + rng.normal(0, noise, (Hh, S))
    scale = 1 << (depth - 8)
    dt = pdtype(depth)
    pmax = (1 << depth) - 1
    cur = np.clip(np.rint(cur * scale), 0, pmax).astype(dt).ravel()
    ref = np.clip(np.rint(ref * scale), 0, pmax).astype(dt).ravel()
    return cur, ref, S, pad * S + pad


def make_jobs(pkg, W, H, sizes, merange, rng, n_per_size=24, with_cands=True, mvp_span=24):
    jobs = []
    for (w, h) in sizes:
        for _ in range(n_per_size):
            x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
            y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
            mvp = rng.integers(-mvp_span, mvp_span + 1, 2)
            # Search::setSearchRange (search.cpp:2724-2769) shape: +-merange around the full-pel mvp
            cx, cy = (int(mvp[0]) >> 2), (int(mvp[1]) >> 2)
            nc = int(rng.integers(0, 4)) if with_cands else 0
            mvc = np.zeros((8, 2), dtype=np.int32)
            mvc[:nc] = rng.integers(-40, 41, (nc, 2))
            if nc and rng.integers(0, 3) == 0:
                mvc[0] = mvp          # duplicate of the predictor: exercises the "already measured" test
            jobs.append((x, y, w, h, cx - merange, cy - merange, cx + merange, cy + merange, int(mvp[0]), int(mvp[1]), nc, mvc))
    job = np.zeros(len(jobs), dtype=pkg.ME_JOB)
    for i, (x, y, w, h, a, b, c, d, px, py, nc, mvc) in enumerate(jobs):
        job[i]["puX"], job[i]["puY"], job[i]["w"], job[i]["h"] = x, y, w, h
        job[i]["mvminX"], job[i]["mvminY"], job[i]["mvmaxX"], job[i]["mvmaxY"] = a, b, c, d
        job[i]["mvpX"], job[i]["mvpY"], job[i]["numCand"] = px, py, nc
        job[i]["mvc"] = mvc
    return job


def ref_me(depth, cur, ref, stride, origin, job, method, subme, merange, qp, maxSlices=1, threads=4):
    """Run the reference's own MotionEstimate (oracle/_ref) over the jobs; returns (mvx, mvy, cost)."""
    R = oracle.ref(depth)
    assert R is not None, "oracle/_ref missing"
    rj = np.zeros(len(job), dtype=REF_ME_JOB)
    for f in ("puX", "puY", "w", "h", "mvminX", "mvminY", "mvmaxX", "mvmaxY", "mvpX", "mvpY", "numCand", "mvc"):
        rj[f] = job[f]
    item = cur.itemsize
    R.ref_me_batch(ctypes.c_void_p(cur.ctypes.data + origin * item), ctypes.c_ssize_t(stride),
                   ctypes.c_void_p(ref.ctypes.data + origin * item), ctypes.c_ssize_t(stride),
                   vp(rj), ctypes.c_int64(len(rj)), int(method), int(subme), int(merange), int(qp), int(maxSlices), int(threads))
    return rj["outMvX"].copy(), rj["outMvY"].copy(), rj["outCost"].copy()


def synth_chroma_pair(W, H, pad, csp, depth=8, seed=77, motion=(5, -3), noise=2.0):
    """Cb/Cr plane pairs for a (cur, ref) picture pair: chroma-resolution band-limited noise moved by the luma motion
    scaled to the chroma grid.  Returns (curCb, curCr, refCb, refCr, strideC, originC, padC)."""
    hs, vs = int(csp != 3), int(csp == 1)
    Wc, Hc, pc = W >> hs, H >> vs, pad
    out = []
    for k in range(2):
        c, r, Sc, oc = synth_pair(Wc, Hc, pc, depth=depth, seed=seed + 13 * k, motion=(motion[0] >> hs, motion[1] >> vs), noise=noise)
        out.append((c, r))
    return out[0][0], out[1][0], out[0][1], out[1][1], Sc, oc


def ref_me_chroma(depth, csp, cur, ref, stride, origin, curC, refC, strideC, originC, job, method, subme, merange, qp, maxSlices=1, threads=4):
    """The reference's encode-style setSourcePU + motionEstimate (bChromaSATD decided by the reference)."""
    R = oracle.ref(depth)
    assert R is not None, "oracle/_ref missing"
    rj = np.zeros(len(job), dtype=REF_ME_JOB)
    for f in ("puX", "puY", "w", "h", "mvminX", "mvminY", "mvmaxX", "mvmaxY", "mvpX", "mvpY", "numCand", "mvc"):
        rj[f] = job[f]
    item = cur.itemsize
    P = lambda a, o: ctypes.c_void_p(a.ctypes.data + o * item)
    rc = R.ref_me_batch_chroma(P(cur, origin), P(curC[0], originC), P(curC[1], originC), ctypes.c_ssize_t(stride), ctypes.c_ssize_t(strideC),
                               P(ref, origin), P(refC[0], originC), P(refC[1], originC), ctypes.c_ssize_t(stride), ctypes.c_ssize_t(strideC),
                               int(csp), vp(rj), ctypes.c_int64(len(rj)), int(method), int(subme), int(merange), int(qp), int(maxSlices), int(threads))
    assert rc == 0
    return rj["outMvX"].copy(), rj["outMvY"].copy(), rj["outCost"].copy()


# ---- the general frame search (x265b200_me_frame_ex_dev / csrc/me_ctu_kernels.cu) ------------------------------------------
def ctu_layout(C, minCu, rect, amp):
    """The PUs of one CTU in the order the C ABI documents (x265b200_me_frame_layout), with the CU that owns each:
    rows of (x, y, w, h, cuX, cuY, cuSize)."""
    out = []
    S = C
    while S >= minCu:
        for cy in range(0, C, S):
            for cx in range(0, C, S):
                pus = [(0, 0, S, S)]
                if rect:
                    pus += [(0, 0, S, S // 2), (0, S // 2, S, S // 2), (0, 0, S // 2, S), (S // 2, 0, S // 2, S)]
                if amp and S >= 16:
                    q = S // 4
                    pus += [(0, 0, S, q), (0, q, S, 3 * q), (0, 0, S, 3 * q), (0, 3 * q, S, q),
                            (0, 0, q, S), (q, 0, 3 * q, S), (0, 0, 3 * q, S), (3 * q, 0, q, S)]
                out += [(cx + x, cy + y, w, h, cx, cy, S) for (x, y, w, h) in pus]
        S //= 2
    return np.array(out, dtype=np.int32)


def search_range(picW, picH, C, cuPelX, cuPelY, mvp, merange, slice_bounds=None, ref_lag=None):
    """Search::setSearchRange (search.cpp:2724-2769) with CUData::clipMv (cudata.cpp:1915-1928); returns full-pel (mvmin, mvmax)."""
    mn = [int(mvp[0]) - (merange << 2), int(mvp[1]) - (merange << 2)]
    mx = [int(mvp[0]) + (merange << 2), int(mvp[1]) + (merange << 2)]
    xmax, xmin = (picW + 8 - cuPelX - 1) << 2, -((C + 8 + cuPelX - 1) << 2)
    ymax, ymin = (picH + 8 - cuPelY - 1) << 2, -((C + 8 + cuPelY - 1) << 2)
    for v in (mn, mx):
        v[0] = min(xmax, max(xmin, v[0])); v[1] = min(ymax, max(ymin, v[1]))
    if slice_bounds is not None:
        mn[1] = max(mn[1], int(slice_bounds[0])); mx[1] = min(mx[1], int(slice_bounds[1]))
    L = (1 << 15) - 1
    mn = [max(mn[0], -L) >> 2, max(mn[1], -L) >> 2]
    mx = [min(mx[0], L) >> 2, min(mx[1], L) >> 2]
    if ref_lag is not None:
        mn[1] = min(mn[1], ref_lag); mx[1] = min(mx[1], ref_lag)
    mx[1] = max(mx[1], mn[1])
    return mn, mx


def ctu_jobs(pkg, layout, C, ctuX, ctuY, picW, picH, mvp, merange, ncand=None, mvc=None, slice_bounds=None, ref_lag=None):
    """ME_JOB records (for the reference's MotionEstimate) of the PUs of one CTU: mvp [nPU][2], optional candidates.
    Returns (jobs, searched) -- searched[i] False for PUs of CUs that leave the picture."""
    n = len(layout)
    job = np.zeros(n, dtype=pkg.ME_JOB)
    searched = np.ones(n, dtype=bool)
    for i, (x, y, w, h, cuX, cuY, S) in enumerate(layout.tolist()):
        cuPelX, cuPelY = ctuX * C + cuX, ctuY * C + cuY
        if cuPelX + S > picW or cuPelY + S > picH:
            searched[i] = False
            continue
        mn, mx = search_range(picW, picH, C, cuPelX, cuPelY, mvp[i], merange, slice_bounds, ref_lag)
        job[i]["puX"], job[i]["puY"], job[i]["w"], job[i]["h"] = ctuX * C + x, ctuY * C + y, w, h
        job[i]["mvminX"], job[i]["mvminY"], job[i]["mvmaxX"], job[i]["mvmaxY"] = mn[0], mn[1], mx[0], mx[1]
        job[i]["mvpX"], job[i]["mvpY"] = int(mvp[i][0]), int(mvp[i][1])
        if ncand is not None:
            k = int(ncand[i])
            job[i]["numCand"] = k
            job[i]["mvc"][:k] = mvc[i][:k]
    return job, searched


def box_blur(a, k=5):
    """separable k-tap box filter by cumulative sums (fast band-limiting for full-size frames)"""
    for ax in (0, 1):
        c = np.cumsum(a, axis=ax, dtype=np.float64)
        pad = [(0, 0), (0, 0)]; pad[ax] = (k, 0)
        c = np.pad(c, pad)
        n = a.shape[ax]
        a = (np.take(c, np.arange(k, n + k), axis=ax) - np.take(c, np.arange(0, n), axis=ax)) / k
    return a


def synth_sequence(W, H, padX, padY, depth, nframes, seed, max_motion=12, noise=2.5):
    """nframes padded planes (stride W + 2*padX, H + 2*padY rows): band-limited noise under per-frame global motion + noise.
    Returns (list of flat arrays, stride, rows, origin offset in elements)."""
    rng = np.random.default_rng(seed)
    S, R = W + 2 * padX, H + 2 * padY
    m = max_motion + 8
    big = box_blur(rng.uniform(0, 255, (R + 2 * m, S + 2 * m)))
    big = (big - big.min()) / (big.max() - big.min()) * 255.0
    scale, pmax = 1 << (depth - 8), (1 << depth) - 1
    out = []
    for f in range(nframes):
        dx, dy = (0, 0) if f == 0 else rng.integers(-max_motion, max_motion + 1, 2)
        fr = big[m + dy:m + dy + R, m + dx:m + dx + S] + rng.normal(0, noise, (R, S))
        out.append(np.clip(np.rint(fr * scale), 0, pmax).astype(pdtype(depth)).ravel())
    return out, S, R, padY * S + padX
