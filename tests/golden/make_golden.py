#!/usr/bin/env python
"""Generate tests/golden/x265_golden.npz from the UNMODIFIED reference compiled under oracle/_ref
(run in the authoring container where /root/reference exists:  python tests/golden/make_golden.py).
The fixture pins oracle + CUDA path on boxes that have no reference at all: inputs AND the reference's
outputs are stored.  Families: block compares (all 25 PU sizes / 5 CU sizes, 8- and 10-bit), DCT/IDCT/DST,
quant, luma interpolation (hv), intra (35 modes), motionEstimate jobs (HEX/STAR), lookahead frame costs."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib  # noqa: E402

import oracle  # noqa: E402
from util import CU_SIZES, LUMA_PU_SIZES, STRIDE, block_offsets, pixel_buffers, ref_cmp, short_buffers, vpo, ssz  # noqa: E402
from me_util import make_jobs, ref_me, synth_pair  # noqa: E402

pkg = importlib.import_module("x265-yuuki-asuna_b200")
out = {}
for depth in (8, 10):
    bufs = pixel_buffers(depth, seed=2024, size=64 * 80)
    sb = short_buffers(depth, seed=2025, size=64 * 80)
    out["px%d" % depth] = bufs[0]; out["sh%d" % depth] = sb[0]
    offs = block_offsets(6); offb = offs[::-1].copy() + 5
    out["offs"] = offs; out["offb"] = offb
    for kind in ("sad", "satd"):
        out["%s%d" % (kind, depth)] = np.array([ref_cmp(kind, depth, w, h, bufs[0], STRIDE, bufs[0], 61, offs, offb) for (w, h) in LUMA_PU_SIZES], dtype=np.int64)
    for kind in ("sa8d", "sse_pp"):
        out["%s%d" % (kind, depth)] = np.array([ref_cmp(kind, depth, s, s, bufs[0], STRIDE, bufs[0], 61, offs, offb) for s in CU_SIZES], dtype=np.uint64)
    for kind in ("sse_ss", "ssd_s"):
        out["%s%d" % (kind, depth)] = np.array([ref_cmp(kind, depth, s, s, sb[0], STRIDE, sb[0], 61, offs, offb) for s in CU_SIZES], dtype=np.uint64)
    R = oracle.ref(depth)
    # transforms
    for idx, N in ((0, 4), (1, 8), (2, 16), (3, 32), (4, 4)):
        src = sb[0][:N * N * 3]
        d = np.empty(N * N * 3, dtype=np.int16); e = np.empty(N * N * 3, dtype=np.int16)
        for i in range(3):
            R.ref_dct(idx, vpo(src, i * N * N), vpo(d, i * N * N), ssz(N))
            R.ref_idct(idx, vpo(d, i * N * N), vpo(e, i * N * N), ssz(N))
        out["dct%d_%d" % (idx, depth)] = d; out["idct%d_%d" % (idx, depth)] = e
    # luma hv interpolation 16x16, all 9 fractional combos
    plane = bufs[0]
    res = []
    for cx in (1, 2, 3):
        for cy in (1, 2, 3):
            e = np.zeros(256, dtype=plane.dtype)
            R.ref_interp(6, -1, 2, vpo(plane, 10 * STRIDE + 10), ssz(STRIDE), vpo(e, 0), ssz(16), cx, cy)
            res.append(e)
    out["hvpp16_%d" % depth] = np.array(res)
    # intra 8x8 all modes, unfiltered
    nb = bufs[0][100:133].copy()
    res = []
    for m in range(35):
        e = np.zeros(64, dtype=plane.dtype)
        R.ref_intra_pred(1, m, vpo(e, 0), ssz(8), vpo(nb, 0), 1)
        res.append(e)
    out["intra8_nb_%d" % depth] = nb; out["intra8_%d" % depth] = np.array(res)

# motion estimation jobs (8-bit)
W, H, merange = 128, 96, 24
cur, ref, S, origin = synth_pair(W, H, 64 + merange + 16, depth=8, seed=77)
rng = np.random.default_rng(78)
job = make_jobs(pkg, W, H, [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (32, 24)], merange, rng, n_per_size=4)
out["me_cur"] = cur; out["me_ref"] = ref; out["me_geom"] = np.array([W, H, S, origin, merange]); out["me_jobs"] = job.view(np.int32).reshape(len(job), -1)
for name, m, sub in (("hex2", 1, 2), ("star3", 3, 3), ("dia0", 0, 0), ("umh2", 2, 2)):
    x, y, c = ref_me(8, cur, ref, S, origin, job, m, sub, merange, 30)
    out["me_" + name] = np.stack([x, y, c], axis=1)

path = os.path.join(ROOT, "tests", "golden", "x265_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")
