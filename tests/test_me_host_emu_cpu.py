"""CPU: the DEVICE source of the frame search (csrc/me_device.cuh, the thread-only build me_frame_kernels.cu compiles, with
the same variant switches) is compiled for the host through tests/host_emu/ and run, lane for lane, on the CTUs of a small
frame: every 2Nx2N PU 64..8 must return the reference's motionEstimate result (oracle/_ref, unmodified motion.cpp).
The lanes of a PU are host threads that meet at a barrier for every warp shuffle, so the sub-block decomposition, the lane
reductions and the data-dependent control flow of the kernel are all exercised without a GPU.  (The GPU tests remain the
parity tests proper; this one lets changes to the search code be checked on the CPU first.)"""
import ctypes
import importlib
import os
import subprocess

import numpy as np
import pytest

from me_util import ref_me, synth_pair
from util import oracle, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("x265-yuuki-asuna_b200")
needs_ref = pytest.mark.skipif(not oracle.have_ref(8), reason="oracle/_ref not built (no /root/reference here)")


# the build the kernel ships, the same without the shared vertical cells, and the generic-pointer form kept for A/B runs
VARIANTS = {"default": [], "no_vcell_reuse": ["-DEMU_VCELL_REUSE_OFF=1"], "window_slow": ["-DEMU_WINDOW_SLOW=1"]}


@pytest.fixture(scope="module", params=sorted(VARIANTS))
def emu(request, tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / ("me_frame_emu_%s.so" % request.param))
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas"] + VARIANTS[request.param] +
                   ["-I", os.path.join(ROOT, "tests", "host_emu"), "-I", os.path.join(ROOT, "x265-yuuki-asuna_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), "-o", so, os.path.join(ROOT, "tests", "host_emu", "me_frame_emu.cpp")], check=True)
    return ctypes.CDLL(so)


def _check(emu, depth, method, subme, merange, seed, ctus):
    ctuCols, ctuRows = 2, 2
    W, H, pad = ctuCols * 64, ctuRows * 64, 144
    cur, ref, S, origin = synth_pair(W, H, pad, depth=depth, seed=seed, motion=(7, -4))
    item = cur.itemsize
    cost = pkg.bitcost_table(pkg.lambda_for_qp(30, depth))
    rng = np.random.default_rng(seed)
    for (cx, cy) in ctus:
        mvp = rng.integers(-40, 41, 2).astype(np.int32) if (cx, cy) != (0, 0) else np.zeros(2, dtype=np.int32)
        out = np.zeros((85, 3), dtype=np.int32)
        rc = emu.emu_me_frame_ctu(depth, ctypes.c_void_p(cur.ctypes.data + origin * item), ctypes.c_int64(S),
                                  ctypes.c_void_p(ref.ctypes.data + origin * item), ctypes.c_int64(S), pad, cx, cy, int(mvp[0]), int(mvp[1]),
                                  int(method), subme, merange, vp(cost), vp(out))
        assert rc == 0, rc          # -2: the lanes of a PU disagreed
        jobs = []
        for level in range(4):
            s, per = 64 >> level, 1 << level
            for py in range(per):
                for px in range(per):
                    jobs.append((cx * 64 + px * s, cy * 64 + py * s, s))
        job = np.zeros(len(jobs), dtype=pkg.ME_JOB)
        for i, (x, y, s) in enumerate(jobs):
            job[i]["puX"], job[i]["puY"], job[i]["w"], job[i]["h"] = x, y, s, s
            job[i]["mvpX"], job[i]["mvpY"] = int(mvp[0]), int(mvp[1])
            job[i]["mvminX"], job[i]["mvminY"] = (int(mvp[0]) >> 2) - merange, (int(mvp[1]) >> 2) - merange
            job[i]["mvmaxX"], job[i]["mvmaxY"] = (int(mvp[0]) >> 2) + merange, (int(mvp[1]) >> 2) + merange
        ex, ey, ec = ref_me(depth, cur, ref, S, origin, job, method, subme, merange, 30)
        bad = np.nonzero((out[:, 0] != ex) | (out[:, 1] != ey) | (out[:, 2] != ec))[0]
        assert not len(bad), (depth, method, subme, merange, (cx, cy), int(bad[0]), jobs[bad[0]], out[bad[0]].tolist(), int(ex[bad[0]]), int(ey[bad[0]]), int(ec[bad[0]]), len(bad))


@needs_ref
@pytest.mark.parametrize("depth,method,subme,merange", [(8, 1, 2, 57), (8, 3, 3, 24), (8, 0, 0, 16), (8, 2, 5, 32), (8, 1, 7, 16), (10, 1, 2, 40), (10, 3, 4, 16)])
def test_device_search_source_on_host_equals_reference(emu, depth, method, subme, merange):
    if depth > 8 and not oracle.have_ref(10):
        pytest.skip("10-bit oracle/_ref not built")
    _check(emu, depth, method, subme, merange, seed=300 + depth + 7 * method + subme, ctus=[(0, 0), (1, 1)])
