"""CPU: the per-cell prediction + cost functions of la_intra_kernel (csrc/lookahead_kernels.cu: the closed-form DC / planar cells and the
mode-as-data angular cells with their projected reference line) compiled for the host and compared, for every mode 0..34 of an 8x8
lowres CU, with the reference's own intra_pred[mode] + intra_filter + pu[LUMA_8x8].satd chained as lowresIntraEstimate chains them
(slicetype.cpp:731-771: planar from the filtered neighbours, DC with the edge filter, angular modes from neighbours[g_intraFilterFlags &
8]); 8 and 10 bit."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from util import LUMA_PU_SIZES, oracle, pdtype, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not oracle.have_ref(8), reason="oracle/_ref not built (no /root/reference here)")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    d = tmp_path_factory.mktemp("laintra")
    src = open(os.path.join(ROOT, "x265-yuuki-asuna_b200", "csrc", "lookahead_kernels.cu")).read()
    a = src.index("// [host-testable:"); a = src.index("\n", a) + 1
    b = src.index("// [host-testable end]")
    body = src[a:b]
    c0 = body.index("struct LAIntraArgs"); c1 = body.index("};", c0) + 2          # the kernel's argument block is not needed on the host
    open(os.path.join(d, "la_intra_cell_src.inc"), "w").write(body[:c0] + body[c1:])
    so = str(d / "la_intra_cell_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-I", str(d), "-o", so,
                    os.path.join(ROOT, "tests", "host_emu", "la_intra_cell_emu.cpp")], check=True)
    return ctypes.CDLL(so)


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
def test_la_intra_cells_equal_reference_prediction_and_satd(emu, depth):
    if not oracle.have_ref(depth):
        pytest.skip("oracle/_ref for this depth not built")
    R = oracle.ref(depth)
    dt = pdtype(depth)
    rng = np.random.default_rng(3 + depth)
    mx = (1 << depth) - 1
    part = LUMA_PU_SIZES.index((8, 8))
    for it in range(60):
        if it % 5 == 0:
            nb = np.where(np.arange(33) & 1, mx, 0).astype(dt)              # extreme neighbours: clipping of the edge filters
        else:
            nb = rng.integers(0, mx + 1, 33).astype(dt)
        src = rng.integers(0, mx + 1, 64).astype(dt)
        filt = np.zeros(33, dtype=dt)
        R.ref_intra_filter(1, vp(nb), vp(filt))                              # cu[BLOCK_8x8].intra_filter
        smp, flt, fenc = nb.astype(np.int32), filt.astype(np.int32), src.astype(np.int32)
        dc = (8 + int(nb[1:9].astype(np.int64).sum()) + int(nb[17:25].astype(np.int64).sum())) // 16
        for mode in range(35):
            pred = np.zeros(64, dtype=dt)
            if mode == 0:
                R.ref_intra_pred(1, 0, vp(pred), ctypes.c_ssize_t(8), vp(filt), 0)          # planar: filtered neighbours (cuSize >= 8)
            elif mode == 1:
                R.ref_intra_pred(1, 1, vp(pred), ctypes.c_ssize_t(8), vp(nb), 1)            # DC: raw neighbours, edge filter (cuSize <= 16)
            else:
                use_f = min(abs(mode - 26), abs(mode - 10)) > 7                                  # g_intraFilterFlags[mode] & 8
                R.ref_intra_pred(1, mode, vp(pred), ctypes.c_ssize_t(8), vp(filt if use_f else nb), 1)
            want = R.ref_pixelcmp(1, part, vp(src), ctypes.c_ssize_t(8), vp(pred), ctypes.c_ssize_t(8))
            got = emu.emu_la_intra_mode_cost(vp(smp), vp(flt), vp(fenc), mode, dc, depth)
            assert got == want, (depth, it, mode, got, want)
