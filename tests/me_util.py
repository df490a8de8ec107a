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
