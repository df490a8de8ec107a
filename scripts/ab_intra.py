#!/usr/bin/env python
"""A/B of the two forms of the 8-bit all-modes intra prediction behind x265b200_intra_modes_dev, in ONE process, no torch:
the shipped persistent shared-memory kernel and the staged cell form (csrc/intra_cell.cuh, X265B200_INTRA_FAST=1 -- the switch
is read on every call).  Workload = the bench's intra stage: all 35 modes of every 8x8 / 16x16 / 32x32 block of a 2160p frame.
Checks that the prediction arrays are identical, then times both (host clock around stream syncs)."""
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
    log = open(os.path.join(ROOT, "gpurun_out", "ab_intra.log"), "a")

    def say(**kw):
        line = json.dumps(kw)
        print(line, flush=True)
        log.write(line + "\n"); log.flush()

    ctx = pkg.Ctx(0)
    fr, S, R = frames(bench.W, bench.CTU_ROWS * bench.CTU, bench.PAD, 8, 0, 777)
    frame = np.zeros(bench.ROWS * bench.STRIDE, dtype=np.uint8)
    frame[:min(len(frame), len(fr[0]))] = fr[0][:len(frame)]
    for sizeIdx, N, log2N in bench.INTRA_SIZES:
        nbr = np.ascontiguousarray(bench.neighbour_arrays(frame, N))
        n = len(nbr)
        dN, dD = ctx.to_device(nbr), ctx.empty(n * 35 * N * N)
        bLuma = int(N <= 16)
        out = {}
        for mode in ("0", "1"):
            os.environ["X265B200_INTRA_FAST"] = mode
            ctx.intra_modes_dev(8, log2N, dN, dD, bLuma, n); ctx.sync()
            out[mode] = dD.download(np.uint8)
        equal = bool(np.array_equal(out["0"], out["1"]))
        t = {"0": [], "1": []}
        for rep in range(8):
            for mode in ("0", "1"):
                os.environ["X265B200_INTRA_FAST"] = mode
                ctx.sync(); t0 = time.perf_counter()
                ctx.intra_modes_dev(8, log2N, dN, dD, bLuma, n)
                ctx.sync(); t[mode].append((time.perf_counter() - t0) * 1e3)
        say(stage="intra", N=N, blocks=n, equal=equal, differing=int(np.count_nonzero(out["0"] != out["1"])),
            base_ms=round(float(np.median(t["0"][2:])), 4), cell_ms=round(float(np.median(t["1"][2:])), 4),
            speedup=round(float(np.median(t["0"][2:]) / np.median(t["1"][2:])), 3))
        dN.free(); dD.free()
    ctx.close()


if __name__ == "__main__":
    main()
