#!/usr/bin/env python
"""A/B of the two forms of the 8-bit luma interpolation behind x265b200_interp_dev, in ONE process, no torch:
the shipped one-pixel-per-thread kernel and the staged cell form (csrc/interp_cell.cuh, X265B200_INTERP_FAST=1 -- the
switch is read on every call).  Workload = the bench's interpolation stage: one 8-tap interpolation per PU and level of a
2160p frame, all 15 fractions (173 400 blocks in 12 launches).  Checks that the prediction planes are identical, then times
both (host clock around stream syncs).  One JSON line per stage to stdout and gpurun_out/ab_interp.log."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
pkg = importlib.import_module("x265-yuuki-asuna_b200")
import bench                      # job builders only (no torch at import)
from ab_me_frame import frames


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "ab_interp.log"), "a")

    def say(**kw):
        line = json.dumps(kw)
        print(line, flush=True)
        log.write(line + "\n"); log.flush()

    ctx = pkg.Ctx(0)
    W, Hp, PAD, S = bench.W, bench.CTU_ROWS * bench.CTU, bench.PAD, bench.STRIDE
    fr, S2, R = frames(W, Hp, PAD, 8, 0, 4321)
    assert S2 == S
    dRef = ctx.to_device(fr[0])
    origin = PAD * S + PAD
    jobs = {s: {k: ctx.to_device(a) for k, a in bench.interp_jobs(pkg, s).items() if len(a)} for s in bench.LEVELS}
    counts = {s: {k: len(a) for k, a in bench.interp_jobs(pkg, s).items() if len(a)} for s in bench.LEVELS}
    pred = {s: ctx.empty(W * Hp) for s in bench.LEVELS}

    def run():
        for s in bench.LEVELS:
            for kind, dj in jobs[s].items():
                ctx.interp_dev(kind, 8, 8, s, s, dRef.ptr + origin, S, pred[s], W, dj, counts[s][kind], 0)

    out = {}
    for mode in ("0", "1"):
        os.environ["X265B200_INTERP_FAST"] = mode
        for p in pred.values():
            p.upload(np.zeros(W * Hp, dtype=np.uint8))
        run(); ctx.sync()
        out[mode] = {s: pred[s].download(np.uint8) for s in bench.LEVELS}
    say(stage="parity", equal={str(s): bool(np.array_equal(out["0"][s], out["1"][s])) for s in bench.LEVELS},
        blocks=int(sum(sum(c.values()) for c in counts.values())))
    t = {"0": [], "1": []}
    for rep in range(12):
        for mode in ("0", "1"):
            os.environ["X265B200_INTERP_FAST"] = mode
            ctx.sync(); t0 = time.perf_counter(); run(); ctx.sync()
            t[mode].append((time.perf_counter() - t0) * 1e3)
    say(stage="timing", base_ms=round(float(np.median(t["0"][2:])), 4), cell_ms=round(float(np.median(t["1"][2:])), 4),
        speedup=round(float(np.median(t["0"][2:]) / np.median(t["1"][2:])), 3))
    ctx.close()


if __name__ == "__main__":
    main()
