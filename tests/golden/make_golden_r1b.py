#!/usr/bin/env python
"""Generate tests/golden/x265_golden_r1b.npz from the UNMODIFIED reference compiled under oracle/_ref (run where /root/reference
exists: python tests/golden/make_golden_r1b.py).  Families added late in round 1 -- inputs AND the reference's outputs:
pu[].ads, the 12 SEA integral planes, SAO offset application and statistics, the deblocking line filters, propagateCost, and the
residual chain of one TU (levels, numSig, recon, SSE).  8- and 10-bit where the entry depends on the pixel type."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from util import LUMA_PU_SIZES, pdtype, vp, vpo, ssz  # noqa: E402

out = {}
rng = np.random.default_rng(20261017)

# ---- ads (pixel-type independent) ----
R = oracle.ref(8)
stride, width = 128, 60
sums = rng.integers(0, 30000, stride * 12).astype(np.uint32)
cost = rng.integers(0, 200, width).astype(np.uint16)
cases, res = [], []
for (w, h), kind in (((8, 8), 1), ((16, 8), 2), ((16, 16), 4), ((32, 24), 4)):
    enc = rng.integers(0, 30000, 4).astype(np.int32)
    delta, thresh, off = (8 * stride if kind != 1 else 8), int(rng.integers(5000, 40000)), int(rng.integers(0, stride * 2))
    mv = np.zeros(width + 8, dtype=np.int16)
    n = R.ref_ads(LUMA_PU_SIZES.index((w, h)), vp(enc), vpo(sums, off), delta, vp(cost), vp(mv), width, thresh)
    cases.append([kind, w >> 1, delta, thresh, off, n] + enc.tolist()); res.append(mv[:width].copy())
out["ads_sums"], out["ads_cost"], out["ads_cases"], out["ads_mvs"] = sums, cost, np.array(cases, dtype=np.int64), np.array(res)

buf = rng.integers(-1, 2, 160).astype(np.int8); offs = rng.integers(-7, 8, 32).astype(np.int8)
diff = rng.integers(-50, 51, 64 * 64).astype(np.int16)
out["sao_buf"], out["sao_offs"], out["sao_diff"] = buf, offs, diff
for depth in (8, 10):
    R = oracle.ref(depth)
    dt = pdtype(depth)
    # ---- SEA integral planes of a small padded plane ----
    W, H, padX, padY = 48, 32, 40, 36
    S, rows = W + 2 * padX, H + 2 * padY
    plane = rng.integers(0, 1 << depth, S * rows).astype(dt)
    origin = padY * S + padX
    pl = [np.zeros(S * rows, dtype=np.uint32) for _ in range(12)]
    arr = (ctypes.c_void_p * 12)(*[p.ctypes.data + origin * 4 for p in pl])
    R.ref_sea_integrals(vpo(plane, origin), ssz(S), padX, padY, H, arr)
    out["sea_plane%d" % depth], out["sea_geom"] = plane, np.array([W, H, padX, padY])
    # two checksums per plane over the region the reference finalises (box origins in rows 1 .. rows-1-h, columns < S - w):
    # sum and position-weighted sum mod 2^64 -- the planes themselves would be 1.3 MB of fixture
    PW_, PH_ = [32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4], [32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4]
    chk = []
    for k in range(12):
        reg = pl[k].reshape(rows, S)[1:rows - PH_[k], :S - PW_[k]].astype(np.uint64)
        wts = (np.arange(reg.size, dtype=np.uint64).reshape(reg.shape) * np.uint64(2654435761) + np.uint64(1))
        chk.append([int(reg.sum(dtype=np.uint64)), int((reg * wts).sum(dtype=np.uint64))])
    out["sea_chk%d" % depth] = np.array(chk, dtype=np.uint64)
    # ---- SAO / deblock on a smooth-ish image ----
    IS, IR = 96, 48
    base = rng.integers(0, 1 << depth, (IR // 4 + 1, IS // 4 + 1))
    img = (np.kron(base, np.ones((4, 4), dtype=np.int64))[:IR, :IS] + rng.integers(-2, 3, (IR, IS))).clip(0, (1 << depth) - 1).astype(dt).ravel()
    out["sao_img%d" % depth] = img
    for kind in range(6):
        r, b = img.copy(), buf.copy()
        R.ref_sao_apply(kind, vpo(r, IS * 2 + 3), ssz(IS), vpo(b, 1), vpo(b, 81), vp(offs), 40, 6, 1)
        out["sao_apply%d_%d" % (kind, depth)] = r; out["sao_applybuf%d_%d" % (kind, depth)] = b
    for kind in (5, 0, 1, 3, 4):
        b = buf.copy(); st = np.arange(32, dtype=np.int32); ct = np.arange(32, dtype=np.int32) * 2
        R.ref_sao_stats(kind, vp(diff), vpo(img, IS * 2 + 3), ssz(IS), vpo(b, 2), vpo(b, 82), 38, 30, vp(st), vp(ct))
        out["sao_stats%d_%d" % (kind, depth)] = np.stack([st, ct]); out["sao_statsbuf%d_%d" % (kind, depth)] = b
    for chroma in (0, 1):
        p = img.copy()
        for i in range(6):
            step, off = ((IS, 1) if i & 1 else (1, IS))
            R.ref_deblock(chroma, vpo(p, IS * 16 + 12 + 8 * i), ssz(step), ssz(off), 3 + i, (5 if not chroma else -1), -1 * (i & 1))
        out["deblock%d_%d" % (chroma, depth)] = p
    # ---- the residual chain of single TUs ----
    for sizeIdx in range(4):
        N = 4 << sizeIdx
        fenc = rng.integers(0, 1 << depth, N * N).astype(dt)
        pred = (fenc.astype(np.int64) + rng.integers(-9, 10, N * N) * (1 << (depth - 8))).clip(0, (1 << depth) - 1).astype(dt)
        ts = 15 - depth - (sizeIdx + 2)
        qbits = 14 + 4 + (depth - 8) + ts
        qc = np.full(N * N, 16384, dtype=np.int32)
        rec, coef = np.zeros_like(fenc), np.zeros(N * N, dtype=np.int16)
        ns, sse = np.zeros(1, dtype=np.uint32), np.zeros(1, dtype=np.uint64)
        R.ref_tu_pipeline(sizeIdx, 0, vp(fenc), ssz(N), vp(pred), ssz(N), vp(rec), ssz(N), 1, 1, vp(qc), qbits, 171 << (qbits - 9), None,
                          64 << (4 + depth - 8), 6 - ts, vp(coef), vp(ns), vp(sse), 1)
        out["tu_fenc%d_%d" % (sizeIdx, depth)], out["tu_pred%d_%d" % (sizeIdx, depth)] = fenc, pred
        out["tu_rec%d_%d" % (sizeIdx, depth)], out["tu_coef%d_%d" % (sizeIdx, depth)] = rec, coef
        out["tu_meta%d_%d" % (sizeIdx, depth)] = np.array([qbits, 171 << (qbits - 9), 64 << (4 + depth - 8), 6 - ts, int(ns[0]), int(sse[0])], dtype=np.int64)

# ---- propagateCost ----
n = 500
pin = rng.integers(0, 65536, n).astype(np.uint16); intra = rng.integers(0, 50000, n).astype(np.int32); intra[::41] = 0
inter = rng.integers(0, 65536, n).astype(np.uint16); invq = rng.integers(1, 70000, n).astype(np.int32)
dst = np.zeros(n, dtype=np.int32)
oracle.ref(8).ref_propagate_cost(vp(dst), vp(pin), vp(intra), vp(inter), vp(invq), ctypes.c_double(192.0), n)
out["prop_in"], out["prop_intra"], out["prop_inter"], out["prop_invq"], out["prop_out"] = pin, intra, inter, invq, dst

path = os.path.join(ROOT, "tests", "golden", "x265_golden_r1b.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")
