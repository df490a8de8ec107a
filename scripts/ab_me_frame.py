#!/usr/bin/env python
"""A/B of builds of libx265b200.so on the frame-search kernel (x265b200_me_frame_dev), in ONE process, no torch:
   make -C x265-yuuki-asuna_b200/csrc exp EXPNAME=x EXPFLAGS=-D...      # -> libx265b200_x.so (only me_frame differs)
   python scripts/ab_me_frame.py [--exp PATH ...] [--skip-sweep] [--reps N]
Every --exp library is compared with the default one in turn:
1. parity sweep: both libraries search the same small frames (3x2 CTUs, 2 references, random per-CTU predictors) for
   depth 8/10 x DIA/HEX/UMH/STAR x subme 0..7 x merange 16/57 (+ FULL at merange 8); every {mv, cost} must be equal.
   The default library is the one the parity tests pin to the reference, so equality here carries that parity over.
2. 2160p, 3 references, HEX / subme 2 / merange 57 (the bench workload shape): outputs equal, then interleaved timing
   (host clock around stream syncs, `reps` launches each).
Writes one JSON line per stage to stdout and to gpurun_out/ab_me_frame.log (flushed as it goes)."""
import argparse
import ctypes
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("x265-yuuki-asuna_b200")

LOG = None


def say(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    if LOG:
        LOG.write(line + "\n"); LOG.flush(); os.fsync(LOG.fileno())


def open_ctx(path):
    """a pkg.Ctx bound to the library at `path` (the class normally binds the default build)"""
    L = ctypes.CDLL(path)
    L.x265b200_last_error.restype = ctypes.c_char_p
    L.x265b200_stream.restype = ctypes.c_void_p
    L.x265b200_launch_count.restype = ctypes.c_uint64
    c = pkg.Ctx.__new__(pkg.Ctx)
    c.L = L
    h = ctypes.c_void_p()
    if L.x265b200_create(0, ctypes.c_void_p(0), ctypes.byref(h)) != 0:
        raise RuntimeError(L.x265b200_last_error().decode())
    c.h = h
    return c


def box_blur(a, k=5):
    """separable k-tap box filter by cumulative sums (fast stand-in for the bench's band-limited noise)"""
    for ax in (0, 1):
        c = np.cumsum(a, axis=ax, dtype=np.float64)
        pad = [(0, 0), (0, 0)]; pad[ax] = (k, 0)
        c = np.pad(c, pad)
        n = a.shape[ax]
        hi = np.take(c, np.arange(k, n + k), axis=ax); lo = np.take(c, np.arange(0, n), axis=ax)
        a = (hi - lo) / k
    return a


_FRAMES = {}


def frames(W, H, pad, depth, nref, seed):
    key = (W, H, pad, depth, nref, seed)
    if key not in _FRAMES:
        _FRAMES[key] = _frames(W, H, pad, depth, nref, seed)
    return _FRAMES[key]


def _frames(W, H, pad, depth, nref, seed):
    rng = np.random.default_rng(seed)
    S, R = W + 2 * pad, H + 2 * pad
    big = box_blur(rng.uniform(0, 255, (R + 64, S + 64)))
    big = (big - big.min()) / (big.max() - big.min()) * 255.0
    scale, pmax = 1 << (depth - 8), (1 << depth) - 1
    dt = np.uint16 if depth > 8 else np.uint8
    out = []
    for f in range(nref + 1):
        dx, dy = (0, 0) if f == 0 else rng.integers(-12, 13, 2)
        fr = big[32 + dy:32 + dy + R, 32 + dx:32 + dx + S] + rng.normal(0, 2.5, (R, S))
        out.append(np.clip(np.rint(fr * scale), 0, pmax).astype(dt).ravel())
    return out, S, R


def run(ctx, depth, bufs, S, R, pad, ctuCols, ctuRows, mvp, method, subme, merange, dOut):
    item = 2 if depth > 8 else 1
    origin = (pad * S + pad) * item
    ctx.me_frame_dev(depth, bufs[0].ptr + origin, S, [b.ptr + origin for b in bufs[1:]], S, pad, pad, R, ctuCols, ctuRows, 15,
                     mvp, method, subme, merange, pkg.lambda_for_qp(30, depth), dOut)


def compare(A, path, args, t00):
    B = open_ctx(path)
    name = os.path.basename(path)
    say(stage="init", exp=name, s=round(time.time() - t00, 2))

    if not args.skip_sweep:
        ctuCols, ctuRows, NREF, pad = 3, 2, 2, 144
        W, H = ctuCols * 64, ctuRows * 64
        nPU = ctuCols * ctuRows * 85
        bad, ncase = [], 0
        for depth in (8, 10):
            fr, S, R = frames(W, H, pad, depth, NREF, 100 + depth)
            bufsA = [A.to_device(f) for f in fr]; bufsB = [B.to_device(f) for f in fr]
            outA, outB = A.empty(NREF * nPU * 12), B.empty(NREF * nPU * 12)
            rng = np.random.default_rng(9 + depth)
            for method, meranges in ((pkg.ME_DIA, (16, 57)), (pkg.ME_HEX, (16, 57)), (pkg.ME_UMH, (16, 57)), (pkg.ME_STAR, (16, 57)), (pkg.ME_FULL, (8,))):
                for merange in meranges:
                    for subme in range(8):
                        mvp = rng.integers(-40, 41, (NREF, ctuCols * ctuRows, 2)).astype(np.int32)
                        dA, dB = A.to_device(mvp), B.to_device(mvp)
                        run(A, depth, bufsA, S, R, pad, ctuCols, ctuRows, dA, method, subme, merange, outA)
                        run(B, depth, bufsB, S, R, pad, ctuCols, ctuRows, dB, method, subme, merange, outB)
                        a, b = outA.download(np.int32), outB.download(np.int32)
                        ncase += 1
                        if not np.array_equal(a, b):
                            bad.append([depth, int(method), merange, subme, int(np.count_nonzero(a != b))])
                        dA.free(); dB.free()
            for x in bufsA + bufsB + [outA, outB]:
                x.free()
        say(stage="sweep", exp=name, cases=ncase, mismatching_cases=len(bad), first=bad[:8], s=round(time.time() - t00, 2))

    # ---- 2160p ------------------------------------------------------------------------------------------------
    W, H, pad, NREF = 3840, 2176, 128, 3
    ctuCols, ctuRows = W // 64, H // 64
    fr, S, R = frames(W, H, pad, 8, NREF, 1234)
    nPU = ctuCols * ctuRows * 85
    bufsA = [A.to_device(f) for f in fr]; bufsB = [B.to_device(f) for f in fr]
    outA, outB = A.empty(NREF * nPU * 12), B.empty(NREF * nPU * 12)
    go = lambda c, bufs, out: run(c, 8, bufs, S, R, pad, ctuCols, ctuRows, None, pkg.ME_HEX, 2, 57, out)
    go(A, bufsA, outA); go(B, bufsB, outB)
    a, b = outA.download(np.int32), outB.download(np.int32)
    say(stage="2160p_parity", exp=name, searches=NREF * nPU, equal=bool(np.array_equal(a, b)), differing=int(np.count_nonzero((a != b).reshape(-1, 3).any(axis=1))),
        checksum=int(a.astype(np.int64).sum()), s=round(time.time() - t00, 2))
    tA, tB = [], []
    for rep in range(args.reps + 2):
        for c, bufs, out, acc in ((A, bufsA, outA, tA), (B, bufsB, outB, tB)):
            c.sync(); t0 = time.perf_counter()
            go(c, bufs, out)
            c.sync(); acc.append((time.perf_counter() - t0) * 1e3)
    tA, tB = tA[2:], tB[2:]
    say(stage="2160p_timing", exp=name, base_ms=round(float(np.median(tA)), 4), exp_ms=round(float(np.median(tB)), 4), base_min=round(min(tA), 4), exp_min=round(min(tB), 4),
        speedup=round(float(np.median(tA) / np.median(tB)), 4), reps=args.reps, s=round(time.time() - t00, 2))
    for x in bufsA + bufsB + [outA, outB]:
        x.free()
    B.close()


def main():
    global LOG
    ap = argparse.ArgumentParser()
    ap.add_argument("--exp", action="append", default=None, help="variant library (repeatable); default libx265b200_exp.so")
    ap.add_argument("--skip-sweep", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    exps = args.exp or [os.path.join(ROOT, "x265-yuuki-asuna_b200", "libx265b200_exp.so")]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    LOG = open(os.path.join(ROOT, "gpurun_out", "ab_me_frame.log"), "a")
    t00 = time.time()
    A = open_ctx(pkg.LIB_PATH)
    say(stage="init", base=pkg.LIB_PATH, s=round(time.time() - t00, 2))
    for path in exps:
        compare(A, path, args, t00)
    A.close()


if __name__ == "__main__":
    main()
