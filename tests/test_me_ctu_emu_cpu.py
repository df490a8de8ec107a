"""CPU: the DEVICE source of the general frame search (csrc/me_ctu_device.cuh + me_device.cuh + me_ctu_layout.h, exactly what
csrc/me_ctu_kernels.cu compiles) run on the host through tests/host_emu/: every PU of whole CTUs -- 2Nx2N / rect / AMP,
CTU 64 / 32 / 16, per-PU predictors and candidates, predictors far outside the staged window, the chroma SATD term, 8/10-bit --
must return what the reference's own MotionEstimate returns for the search range Search::setSearchRange gives that PU."""
import ctypes
import importlib
import os
import subprocess

import numpy as np
import pytest

from me_util import ctu_jobs, ctu_layout, ref_me, ref_me_chroma, synth_chroma_pair, synth_pair
from util import oracle, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("x265-yuuki-asuna_b200")
needs_ref = pytest.mark.skipif(not oracle.have_ref(8), reason="oracle/_ref not built (no /root/reference here)")

EMU_PARAMS = np.dtype([(k, np.int32) for k in ("depth", "ctuSize", "minCuSize", "rect", "amp", "picWidth", "picHeight", "ctuCols", "ctuRows", "marginX",
                                              "marginY", "rowsTotal", "numRefs", "searchMethod", "subpelRefine", "merange", "csp", "maxCand",
                                              "maxSlices", "refLagPixels")])


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "me_ctu_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas",
                    "-I", os.path.join(ROOT, "tests", "host_emu"), "-I", os.path.join(ROOT, "x265-yuuki-asuna_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), "-o", so, os.path.join(ROOT, "tests", "host_emu", "me_ctu_emu.cpp")], check=True)
    return ctypes.CDLL(so)


def run_case(emu, depth, C, minCu, rect, amp, method, subme, merange, csp, seed, ctus, maxCand=3, far=False, picW=None, picH=None,
             ctuCols=2, ctuRows=2, slice_bounds=None):
    W, H = ctuCols * C, ctuRows * C
    picW, picH = picW or W, picH or H
    pad = 64 + merange + 48 + (120 if far else 0)
    cur, ref, S, origin = synth_pair(W, H, pad, depth=depth, seed=seed, motion=(7, -4))
    item = cur.itemsize
    hs, vs = int(csp in (1, 2)), int(csp == 1)
    if csp:
        cCb, cCr, rCb, rCr, Sc, oc = synth_chroma_pair(W, H, pad, csp, depth=depth, seed=seed + 5)
    cost = pkg.bitcost_table(pkg.lambda_for_qp(30, depth))
    layout = ctu_layout(C, minCu, rect, amp)
    n = len(layout)
    lay = np.zeros((n, 4), dtype=np.int32)
    assert emu.emu_me_ctu_layout(C, minCu, int(rect), int(amp), vp(lay), n) == n and np.array_equal(lay, layout[:, :4])
    rng = np.random.default_rng(seed)
    P = np.zeros(1, dtype=EMU_PARAMS)
    for k, v in dict(depth=depth, ctuSize=C, minCuSize=minCu, rect=int(rect), amp=int(amp), picWidth=picW, picHeight=picH, ctuCols=ctuCols, ctuRows=ctuRows,
                     marginX=pad, marginY=pad, rowsTotal=H + 2 * pad, numRefs=1, searchMethod=int(method), subpelRefine=subme, merange=merange, csp=csp,
                     maxCand=maxCand, maxSlices=1, refLagPixels=0).items():
        P[k] = v
    ptr = lambda a, o: ctypes.c_void_p(a.ctypes.data + o * item)
    for (cx, cy) in ctus:
        mvpCtu = rng.integers(-30, 31, 2).astype(np.int32) if (cx, cy) != (0, 0) else np.zeros(2, dtype=np.int32)
        mvp = (mvpCtu[None, :] + rng.integers(-12, 13, (n, 2))).astype(np.int32)
        if far:          # some predictors far outside the CTU's staged window, in both directions
            idx = rng.choice(n, max(2, n // 6), replace=False)
            mvp[idx] += (rng.choice([-1, 1], (len(idx), 2)) * rng.integers(4 * (merange + 20), 4 * (merange + 60), (len(idx), 2))).astype(np.int32)
        ncand = rng.integers(0, maxCand + 1, n).astype(np.uint8) if maxCand else np.zeros(n, dtype=np.uint8)
        mvc = (mvpCtu[None, None, :] + rng.integers(-40, 41, (n, max(maxCand, 1), 2))).astype(np.int32)
        out = np.full((n, 3), -777, dtype=np.int32)
        npu = ctypes.c_int32(0)
        sb = None if slice_bounds is None else np.asarray(slice_bounds, dtype=np.int32)
        rc = emu.emu_me_ctu(vp(P), ptr(cur, origin), ptr(cCb, oc) if csp else None, ptr(cCr, oc) if csp else None, ctypes.c_int64(S), ctypes.c_int64(Sc if csp else 0),
                            ptr(ref, origin), ptr(rCb, oc) if csp else None, ptr(rCr, oc) if csp else None, ctypes.c_int64(S), ctypes.c_int64(Sc if csp else 0),
                            cx, cy, vp(mvpCtu), vp(mvp), vp(ncand), vp(mvc), vp(sb) if sb is not None else None, vp(cost), vp(out), ctypes.byref(npu))
        assert rc == 0 and npu.value == n, (rc, npu.value, n)
        job, searched = ctu_jobs(pkg, layout, C, cx, cy, picW, picH, mvp, merange, ncand, mvc,
                                 slice_bounds=None if sb is None else sb[2 * cy: 2 * cy + 2])
        js = job[searched]
        if csp:
            ex, ey, ec = ref_me_chroma(depth, csp, cur, ref, S, origin, (cCb, cCr), (rCb, rCr), Sc, oc, js, method, subme, merange, 30)
        else:
            ex, ey, ec = ref_me(depth, cur, ref, S, origin, js, method, subme, merange, 30)
        got = out[searched]
        bad = np.nonzero((got[:, 0] != ex) | (got[:, 1] != ey) | (got[:, 2] != ec))[0]
        assert not len(bad), ((cx, cy), int(bad[0]), js[bad[0]], got[bad[0]].tolist(), int(ex[bad[0]]), int(ey[bad[0]]), int(ec[bad[0]]), len(bad), len(js))
        assert (out[~searched] == [0, 0, -1]).all()


@needs_ref
@pytest.mark.parametrize("depth,C,minCu,rect,amp,method,subme,merange,csp", [
    (8, 64, 8, False, False, 1, 2, 57, 0),      # config 3 shape: 2Nx2N, HEX subme 2
    (8, 32, 16, False, False, 0, 0, 57, 0),     # config 2: CTU 32 / minCU 16, DIA subme 0
    (8, 64, 8, True, False, 3, 3, 24, 1),       # rect + STAR + chroma SATD 4:2:0
    (8, 64, 8, True, True, 2, 5, 16, 1),        # rect + AMP + UMH subme 5, chroma
    (8, 16, 8, True, True, 1, 2, 16, 0),        # CTU 16
])
def test_ctu_search_on_host_equals_reference(emu, depth, C, minCu, rect, amp, method, subme, merange, csp):
    run_case(emu, depth, C, minCu, rect, amp, method, subme, merange, csp, seed=4000 + 13 * method + subme + C, ctus=[(0, 0), (1, 1)])


@needs_ref
def test_ctu_search_10bit_config4_shape(emu):
    if not oracle.have_ref(10):
        pytest.skip("10-bit oracle/_ref not built")
    run_case(emu, 10, 64, 8, True, False, 3, 3, 24, 1, seed=4100, ctus=[(1, 0)])


@needs_ref
def test_ctu_search_predictors_outside_the_window(emu):
    """predictors whose search range lies (partly or wholly) outside the CTU's staged window, and zero-MV
    winners from there: every read falls back to the global plane with the same arithmetic."""
    run_case(emu, 8, 64, 16, True, False, 1, 2, 16, 0, seed=4200, ctus=[(1, 0)], far=True)
    run_case(emu, 8, 64, 16, False, False, 2, 3, 16, 1, seed=4201, ctus=[(0, 1)], far=True)


@needs_ref
def test_ctu_search_picture_edge_and_slices(emu):
    """CUs that leave the picture are not searched; clipMv at the picture edge; slice-row bounds on the range."""
    run_case(emu, 8, 64, 8, False, False, 1, 2, 24, 0, seed=4300, ctus=[(1, 1)], picW=2 * 64 - 24, picH=2 * 64 - 40)
    run_case(emu, 8, 64, 16, False, False, 3, 2, 24, 0, seed=4301, ctus=[(0, 0), (1, 1)], slice_bounds=[12, 240, -244, -16])
