"""lookahead path timing at 2160p (lowres 1920x1080, 240x135 CUs): Lowres::init, lowresIntraEstimate, estimateFrameCost."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("x265-yuuki-asuna_b200")
ctx = pkg.Ctx(0)
NF, BF = 14, 4
frames = bench.synth_frames(NF)
d = [ctx.to_device(f) for f in frames]
origin = bench.PAD * bench.STRIDE + bench.PAD
W, H = bench.W, bench.CTU_ROWS * bench.CTU            # coded size
mx, my = 64 + 32, 64 + 16                              # PicYuv luma margins (picyuv.cpp:87-88)
lw, ll = W // 2, H // 2
wcu, hcu = (lw + 7) // 8, (ll + 7) // 8
lw, ll = wcu * 8, hcu * 8
ls = (W // 2 + 2 * mx + 31) & ~31
planesize, padoff = ls * (ll + 2 * my), ls * my + mx
ncu = wcu * hcu
planes = [[ctx.to_device(np.zeros(planesize, dtype=np.uint8)) for _ in range(4)] for _ in range(NF)]
ptrs = np.array([[b.ptr + padoff for b in fr] for fr in planes], dtype=np.int64)

def timeit(fn, reps=3):
    fn(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    ctx.sync()
    return (time.perf_counter() - t0) / reps * 1e3

print("lowres geometry: %dx%d stride %d, %dx%d CUs" % (lw, ll, ls, wcu, hcu))
print("lowres_init  %.3f ms/frame" % timeit(lambda: ctx.lowres_init_dev(8, d[0].ptr + origin, bench.STRIDE, ptrs[0], ls, lw, ll, mx, my)))
for i in range(NF):
    ctx.lowres_init_dev(8, d[i].ptr + origin, bench.STRIDE, ptrs[i], ls, lw, ll, mx, my)
lam = pkg.lambda_for_qp(12, 8)
dIC = [ctx.empty(ncu * 4) for _ in range(NF)]
dIM, dLC0, dRS0, dSm0 = ctx.empty(ncu), ctx.empty(ncu * 2), ctx.empty(hcu * 4), ctx.empty(8)
print("la_intra     %.3f ms/frame" % timeit(lambda: ctx.la_intra_dev(8, ptrs[0, 0], ls, wcu, hcu, None, 5 * int(lam), dIC[0], dIM, dLC0, dRS0, dSm0)))
for i in range(NF):
    ctx.la_intra_dev(8, ptrs[i, 0], ls, wcu, hcu, None, 5 * int(lam), dIC[i], dIM, dLC0, dRS0, dSm0)
dPlanePtrs = ctx.to_device(ptrs)
dIntraPtrs = ctx.to_device(np.array([b.ptr for b in dIC], dtype=np.int64))
nslots = NF * 2 * (BF + 2)
dMv, dMvC = ctx.to_device(np.zeros(nslots * ncu * 2, dtype=np.int32)), ctx.to_device(np.zeros(nslots * ncu, dtype=np.int32))
def slot(b, lst, dist): return (b * 2 + lst) * (BF + 2) + dist
# every (b, list, dist <= BF+1) search that exists among NF frames, issued as B-triples (p0, p1, b) with 2 new searches each
allt = [(b - dd, b + dd, b) for dd in range(1, BF + 2) for b in range(NF) if b - dd >= 0 and b + dd < NF]
import sys
SLICES = int(sys.argv[1]) if len(sys.argv) > 1 else 0          # --lookahead-slices (0 = non-cooperative path)
print("lookahead slices:", SLICES)
ONLY_FULL = len(sys.argv) > 2 and sys.argv[2] == "full"     # one size only (for an ncu capture)
for k in ((len(allt),) if ONLY_FULL else (1, 4, 8, 16, 24, 32, len(allt))):
    wave = allt[:k]
    tr = np.zeros(len(wave), dtype=pkg.LA_TRIPLE)
    for t, (p0, p1, b) in enumerate(wave):
        tr[t]["b"], tr[t]["p0"], tr[t]["p1"] = b, p0, p1
        for lst, dist in ((0, b - p0), (1, p1 - b)):
            tr[t]["mvSlot"][lst] = slot(b, lst, dist)
            tr[t]["doSearch"][lst] = 1
    dLC, dRS, dSm = ctx.empty(len(wave) * ncu * 2), ctx.empty(len(wave) * hcu * 4), ctx.empty(len(wave) * 16)
    ms = timeit(lambda: ctx.la_estimate_dev(8, dPlanePtrs, ls, wcu, hcu, tr, dMv, dMvC, dIntraPtrs, None, dLC, dRS, dSm, lam, lookaheadSlices=SLICES), reps=2)
    ns = 2 * len(wave)
    print("la_estimate  %3d triples = %3d list searches: %8.3f ms  (%.3f ms per search of %d CUs)" % (len(wave), ns, ms, ms / ns, ncu))
