"""One launch of each kernel added or rewritten late in round 1, at the 2160p sizes of bench.py, for
`ncu --set full -k regex:...` (profiles/r01_new_kernels.txt): the fused residual pipeline (4 TU sizes), the persistent
all-angles intra kernel (3 sizes), the SEA integral planes, the motion-compensation driver."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("x265-yuuki-asuna_b200")
ctx = pkg.Ctx(0)
W, HH, S, PAD = bench.W, bench.CTU_ROWS * bench.CTU, bench.STRIDE, bench.PAD
origin = PAD * S + PAD
frames = bench.synth_frames(2)
dCur, dRef = ctx.to_device(frames[1]), ctx.to_device(frames[0])
dPred = ctx.to_device(frames[0].reshape(bench.ROWS, S)[PAD:PAD + HH, PAD:PAD + W].copy())
dRecon = ctx.empty(W * HH)
dQt = ctx.to_device(np.full(1024, 26214, dtype=np.int32))
dCoef, dNs, dSse = ctx.empty(W * HH * 2), ctx.empty((W // 4) * (HH // 4) * 4), ctx.empty((W // 4) * (HH // 4) * 8)
for rep in range(2):
    for idx, N in bench.TU_SIZES:
        qbits, add = bench.quant_params(N)
        ctx.tu_pipeline_dev(idx, 8, 0, dCur.ptr + origin, S, dPred, W, dRecon, W, W // N, HH // N, dQt, qbits, add, None, 40 << 5, 9, dCoef, dNs, dSse)
dOut = ctx.empty(33 * W * HH)
for _, N, log2N in bench.INTRA_SIZES:
    nb = bench.neighbour_arrays(frames[0].ravel(), N)
    dN, dF = ctx.to_device(nb), ctx.empty(nb.nbytes)
    for rep in range(2):
        ctx.intra_filter_dev(8, log2N, dN, dF, len(nb))
        ctx.intra_allangs_dev(8, log2N, dN, dF, dOut, 1, len(nb))
# SEA integral planes of one 2160p reference (12 x uint32 planes = 452 MB)
planes = [ctx.empty(S * bench.ROWS * 4) for _ in range(12)]
for rep in range(2):
    ctx.sea_integral_dev(8, dRef.ptr + origin, S, PAD, PAD, bench.ROWS - 2 * PAD, [b.ptr + origin * 4 for b in planes])
# MC: every 16x16 PU of the frame, bi-prediction from two references at fractional MVs
jobs = np.zeros((W // 16) * (HH // 16), dtype=pkg.MC_JOB)
ys, xs = np.meshgrid(np.arange(0, HH, 16), np.arange(0, W, 16), indexing="ij")
jobs["puX"], jobs["puY"], jobs["w"], jobs["h"] = xs.ravel(), ys.ravel(), 16, 16
jobs["cuX"], jobs["cuY"] = jobs["puX"] & ~63, jobs["puY"] & ~63
jobs["refIdx"] = [0, 1]
rng = np.random.default_rng(1)
jobs["mv"] = rng.integers(-40, 41, (len(jobs), 2, 2))
refs = np.zeros(2 * 2 * 3, dtype=np.uint64)
refs[0 * 3] = dRef.ptr + origin; refs[1 * 3] = dCur.ptr + origin; refs[2 * 3] = dCur.ptr + origin; refs[3 * 3] = dRef.ptr + origin
dRefs, dJ = ctx.to_device(refs), ctx.to_device(jobs)
desc = pkg.MC_DESC(0, 0, 0, 0, W, HH, 64, 2, dRefs.ptr, S, S, dRecon.ptr, None, None, W, W, None)
for rep in range(2):
    ctx.mc_dev(8, desc, dJ, len(jobs), 1, 0)
ctx.sync()
print("launches", ctx.launches)
