"""CPU: the host half of x265b200_la_weights_analyse_dev (x265b200_la_weight_guess: the float scale / offset guess of
LookaheadTLD::weightsAnalyse, slicetype.cpp:886-933) against the reference's own weightsAnalyse (oracle/_ref) over many fades: whenever
the reference decides to weight, the weight our guess would apply must reproduce the reference's weighted plane (wbuffer[0]) bit for
bit; whenever our guess terminates early the reference must not weight; and the reference never weights with what we call the identity."""
import ctypes
import importlib

import numpy as np
import pytest

from util import oracle, pdtype
from test_lookahead_gpu import _arr, _bind, _fade_frames

pkg = importlib.import_module("x265-yuuki-asuna_b200")
needs_ref = pytest.mark.skipif(not oracle.have_ref(8), reason="oracle/_ref not built (no /root/reference here)")


def _weight_pp(plane, depth, scale, denom, offset):
    """weight_pp_c (pixel.cpp:518-543) with the operands weightsAnalyse passes (:947-956)"""
    corr = 14 - depth
    rnd = ((1 << (denom - 1)) if denom else 0) << corr
    v = (plane.astype(np.int64) << corr).astype(np.int16).astype(np.int64)
    out = ((scale * v + rnd) >> (denom + corr)) + offset * (1 << (depth - 8))
    return np.clip(out, 0, (1 << depth) - 1).astype(plane.dtype)


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
def test_weight_guess_reproduces_the_reference_planes(depth):
    if not oracle.have_ref(depth):
        pytest.skip("oracle/_ref for this depth not built")
    R = _bind(oracle.ref(depth))
    R.ref_la_weighted_buffer.restype = ctypes.c_void_p
    W, H, NF = 320, 192, 6
    rng = np.random.default_rng(40 + depth)
    dt = pdtype(depth)
    ct = ctypes.c_uint8 if depth == 8 else ctypes.c_uint16
    seen = {"weighted": 0, "early": 0, "rejected": 0}
    for trial in range(8):
        gains = np.concatenate([[1.0], rng.uniform(0.35, 1.0, NF - 1)])
        if trial % 3 == 0:
            gains[2] = gains[1]                                         # an equal pair: early termination
        h = ctypes.c_void_p(R.ref_la_create(W, H, 3, 0))
        R.ref_la_set_weightp(h, 1)
        frames = _fade_frames(W, H, NF, depth, seed=100 * trial + depth, gains=list(gains), offsets=[0] * NF)
        stats = []
        for i, f in enumerate(frames):
            R.ref_la_add_frame(h, ctypes.c_void_p(f.ctypes.data), ctypes.c_ssize_t(W))
            v = f.astype(np.int64)
            sm, ssd = int(v.sum()), int((v * v).sum())
            ssd -= (sm * sm + (W * H) // 2) // (W * H)
            stats.append((sm, ssd))
            R.ref_la_set_wp_stats(h, i, ctypes.c_uint64(sm), ctypes.c_uint64(ssd))
        g = (ctypes.c_int64 * 11)()
        R.ref_la_geometry(h, g)
        lw, ll, ls, mx, my = [int(v) for v in g][:5]
        planesize = ls * (ll + 2 * my)
        for i in range(NF):
            R.ref_la_intra(h, i)
        out = (ctypes.c_int32 * 2)()
        for b in range(1, NF):
            for p0 in range(max(0, b - 3), b):
                R.ref_la_weights_analyse(h, b, p0, out)
                weighted = int(out[0])
                gss = pkg.la_weight_guess(depth, lw, ll, stats[b][0], stats[b][1], stats[p0][0], stats[p0][1])
                if not gss["measure"]:
                    assert not weighted, (trial, b, p0, gss)
                    seen["early"] += 1
                    continue
                if weighted:
                    assert not gss["identity"], (trial, b, p0, gss)
                    ref0 = _arr(R.ref_la_lowres_buffer(h, p0, 0), ct, (planesize,))
                    wb = _arr(R.ref_la_weighted_buffer(h, 0), ct, (planesize,))
                    mine = _weight_pp(ref0, depth, gss["finScale"], gss["finDenom"], gss["curOffset"])
                    assert np.array_equal(mine, wb), (trial, b, p0, gss)
                    seen["weighted"] += 1
                else:
                    seen["rejected"] += 1
        R.ref_la_destroy(h)
    assert seen["weighted"] >= 10 and seen["early"] >= 1, seen
