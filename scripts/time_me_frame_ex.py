#!/usr/bin/env python
"""Times the general frame search (x265b200_me_frame_ex_dev) at the four BASELINE geometries and, at the config-3 shape, next
to the 2Nx2N-only entry (x265b200_me_frame_dev) with equal results.  Host clock around stream syncs, median of `reps`.
   python scripts/time_me_frame_ex.py [--reps N] [--configs 2,3,4,5]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("x265-yuuki-asuna_b200")
from me_util import synth_sequence  # noqa: E402

CONFIGS = {
    2: dict(W=1920, H=1080, depth=8, C=32, minCu=16, rect=0, amp=0, method=pkg.ME_DIA, subme=0, merange=57, nref=1, csp=0, padX=64, padY=72),
    3: dict(W=3840, H=2176, depth=8, C=64, minCu=8, rect=0, amp=0, method=pkg.ME_HEX, subme=2, merange=57, nref=3, csp=0, padX=128, padY=128),
    4: dict(W=3840, H=2160, depth=10, C=64, minCu=8, rect=1, amp=0, method=pkg.ME_STAR, subme=3, merange=57, nref=4, csp=1, padX=96, padY=80),
    5: dict(W=7680, H=4320, depth=8, C=64, minCu=8, rect=1, amp=1, method=pkg.ME_STAR, subme=5, merange=128, nref=5, csp=1, padX=96, padY=80),
    # reduced frames of configs 4 / 5 for profiling runs (same per-CTU work)
    40: dict(W=1280, H=704, depth=10, C=64, minCu=8, rect=1, amp=0, method=pkg.ME_STAR, subme=3, merange=57, nref=2, csp=1, padX=96, padY=80),
    50: dict(W=1280, H=704, depth=8, C=64, minCu=8, rect=1, amp=1, method=pkg.ME_STAR, subme=5, merange=128, nref=2, csp=1, padX=224, padY=208),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--configs", default="2,3,4,5")
    args = ap.parse_args()
    ctx = pkg.Ctx(0)
    for k in [int(x) for x in args.configs.split(",")]:
        c = CONFIGS[k]
        item = 2 if c["depth"] > 8 else 1
        y, S, R, origin = synth_sequence(c["W"], c["H"], c["padX"], c["padY"], c["depth"], c["nref"] + 1, 100 + k, max_motion=8)
        dY = [ctx.to_device(a) for a in y]
        kw = {}
        if c["csp"]:
            cb, Sc, Rc, oc = synth_sequence(c["W"] // 2, c["H"] // 2, c["padX"] // 2, c["padY"] // 2, c["depth"], c["nref"] + 1, 200 + k, max_motion=4)
            cr, _, _, _ = synth_sequence(c["W"] // 2, c["H"] // 2, c["padX"] // 2, c["padY"] // 2, c["depth"], c["nref"] + 1, 300 + k, max_motion=4)
            dCb, dCr = [ctx.to_device(a) for a in cb], [ctx.to_device(a) for a in cr]
            kw = dict(curC=(dCb[0].ptr + oc * item, dCr[0].ptr + oc * item), curStrideC=Sc, refCb=[b.ptr + oc * item for b in dCb[1:]],
                      refCr=[b.ptr + oc * item for b in dCr[1:]], refStrideC=Sc)
        C = c["C"]
        cols, rows = (c["W"] + C - 1) // C, (c["H"] + C - 1) // C
        n = len(pkg.me_frame_layout(C, c["minCu"], c["rect"], c["amp"]))
        params = dict(depth=c["depth"], ctuSize=C, minCuSize=c["minCu"], rect=c["rect"], amp=c["amp"], picWidth=c["W"], picHeight=c["H"], ctuCols=cols, ctuRows=rows,
                      marginX=c["padX"], marginY=c["padY"], rowsTotal=R, searchMethod=int(c["method"]), subpelRefine=c["subme"], merange=c["merange"], csp=c["csp"],
                      maxCand=0, maxSlices=1)
        params["lambda"] = pkg.lambda_for_qp(30, c["depth"])
        dOut = ctx.empty(c["nref"] * cols * rows * n * 12)
        o = origin * item
        go = lambda: ctx.me_frame_ex_dev(params, dY[0].ptr + o, S, [b.ptr + o for b in dY[1:]], S, dOut, **kw)
        go(); ctx.sync()
        ts = []
        for _ in range(args.reps):
            ctx.sync(); t0 = time.perf_counter(); go(); ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
        out = dOut.download(np.int32).reshape(c["nref"], cols * rows, n, 3)
        searched = int((out[..., 2] >= 0).sum())
        line = dict(config=k, searches=searched, pus_per_ctu=n, ms=round(float(np.median(ts)), 3), ms_min=round(min(ts), 3),
                    Msearches_per_s=round(searched / np.median(ts) / 1e3, 2), checksum=int(out.astype(np.int64).sum()))
        if k == 3:
            per_level = [cols * rows * (1 << l) ** 2 for l in range(4)]
            dOld = ctx.empty(c["nref"] * sum(per_level) * 12)
            old = lambda: ctx.me_frame_dev(8, dY[0].ptr + o, S, [b.ptr + o for b in dY[1:]], S, c["padX"], c["padY"], R, cols, rows, 15, None, pkg.ME_HEX, 2, 57,
                                           params["lambda"], dOld)
            old(); ctx.sync()
            to = []
            for _ in range(args.reps):
                ctx.sync(); t0 = time.perf_counter(); old(); ctx.sync(); to.append((time.perf_counter() - t0) * 1e3)
            line["old_entry_ms"] = round(float(np.median(to)), 3)
            line["old_checksum"] = int(dOld.download(np.int32).astype(np.int64).sum())
            dOld.free()
        print(json.dumps(line), flush=True)
        dOut.free()
        for b in dY + (dCb + dCr if c["csp"] else []):
            b.free()
    ctx.close()


if __name__ == "__main__":
    main()
