"""CPU: the lookahead list search (csrc/la_search_thread.cu = estimateCUCost phase 1) with the DEVICE source of the search
(csrc/me_device.cuh, lowres thread-only build) compiled for the host (tests/host_emu/) and compared with the per-CU MVs and MV
costs the reference's own CostEstimateGroup / Lowres objects produce (oracle/_ref, unmodified slicetype.cpp) -- P and B
frame-triples, both lists, 8- and 10-bit; the shipped build (two lanes per CU, packed-word SATD), the scalar-SATD form and the
one-lane-per-CU form."""
import ctypes
import importlib
import os
import subprocess

import numpy as np
import pytest

from util import oracle, vp
from test_lookahead_gpu import _arr, _bind, _frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("x265-yuuki-asuna_b200")
needs_ref = pytest.mark.skipif(not oracle.have_ref(8), reason="oracle/_ref not built (no /root/reference here)")
VARIANTS = {"default": [], "scalar_satd": ["-DEMU_LA_PACKED_SATD_OFF=1"], "one_lane_per_cu": ["-DEMU_LA_LANES=1"]}


@pytest.fixture(scope="module", params=sorted(VARIANTS))
def emu(request, tmp_path_factory):
    so = str(tmp_path_factory.mktemp("laemu") / ("la_search_emu_%s.so" % request.param))
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas"] + VARIANTS[request.param] +
                   ["-I", os.path.join(ROOT, "tests", "host_emu"), "-I", os.path.join(ROOT, "x265-yuuki-asuna_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), "-o", so, os.path.join(ROOT, "tests", "host_emu", "la_search_emu.cpp")], check=True)
    return ctypes.CDLL(so)


@needs_ref
@pytest.mark.parametrize("depth,W,H", [(8, 320, 192), (8, 416, 240), (10, 320, 192)])
def test_lookahead_search_source_on_host_equals_reference(emu, depth, W, H):
    if not oracle.have_ref(depth):
        pytest.skip("oracle/_ref for this depth not built")
    R = _bind(oracle.ref(depth))
    NF, BF = 5, 3
    h = ctypes.c_void_p(R.ref_la_create(W, H, BF, 0))
    frames = _frames(W, H, NF, depth, seed=W + depth)
    for f in frames:
        R.ref_la_add_frame(h, ctypes.c_void_p(f.ctypes.data), ctypes.c_ssize_t(W))
    g = (ctypes.c_int64 * 11)()
    R.ref_la_geometry(h, g)
    lw, ll, ls, mx, my, wcu, hcu = [int(v) for v in g][:7]
    ncu = wcu * hcu
    px = 2 if depth > 8 else 1
    padoff = ls * my + mx
    for i in range(NF):
        R.ref_la_intra(h, i)
    planes = lambda i: (ctypes.c_void_p * 4)(*[R.ref_la_lowres_buffer(h, i, k) + padoff * px for k in range(4)])
    cost = pkg.bitcost_table(pkg.lambda_for_qp(12 + 6 * (depth - 8), depth))
    for (p0, p1, b) in ((0, 4, 4), (0, 4, 2), (1, 3, 2)):
        R.ref_la_frame_cost(h, p0, p1, b, 0)
        bidir = int(b < p1)
        for lst, ref, dist in ((0, p0, b - p0), (1, p1, p1 - b)):
            if lst == 1 and not bidir:
                continue
            want_mv = _arr(R.ref_la_mvs(h, b, lst, dist), ctypes.c_int32, (ncu, 2))
            want_c = _arr(R.ref_la_mvcosts(h, b, lst, dist), ctypes.c_int32, (ncu,))
            mvs, mvc = np.zeros((ncu, 2), dtype=np.int32), np.zeros(ncu, dtype=np.int32)
            rc = emu.emu_la_search_field(depth, planes(b), planes(ref), ctypes.c_int64(ls), wcu, hcu, bidir, 16, vp(cost), vp(mvs), vp(mvc))
            assert rc == 0
            bad = np.nonzero((mvs != want_mv).any(axis=1) | (mvc != want_c))[0]
            assert not len(bad), ((p0, p1, b), lst, int(bad[0]), mvs[bad[0]].tolist(), want_mv[bad[0]].tolist(), int(mvc[bad[0]]), int(want_c[bad[0]]), len(bad))
