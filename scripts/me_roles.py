"""per-role busy cycles of the me_frame kernel on a 2160p frame (profiling aid)."""
import ctypes, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("x265-yuuki-asuna_b200")
ctx = pkg.Ctx(0)
frames = bench.synth_frames(4)
d = [ctx.to_device(f) for f in frames]
origin = bench.PAD * bench.STRIDE + bench.PAD
n = sum(bench.CTU_COLS * bench.CTU_ROWS * (1 << l) ** 2 for l in range(4))
dOut = ctx.empty(3 * n * 12)
import time
NW = 4   # MF_WARPS in csrc/me_frame_kernels.cu
def counters():
    out = (ctypes.c_uint64 * (NW + 1))()
    ctx.L.x265b200_debug_me_frame_cycles(ctx.h, out)
    return np.array(list(out), dtype=np.float64)
def launch(mask):
    ctx.me_frame_dev(8, d[3].ptr + origin, bench.STRIDE, [d[i].ptr + origin for i in range(3)], bench.STRIDE, bench.PAD, bench.PAD, bench.ROWS,
                     bench.CTU_COLS, bench.CTU_ROWS, mask, None, pkg.ME_HEX, 2, 57, pkg.lambda_for_qp(30, 8), dOut)
launch(15); ctx.sync()
for mask in [int(a) for a in sys.argv[1:]] or [15, 1, 2, 4, 8, 3, 12]:
    c0 = counters(); ctx.sync(); t0 = time.perf_counter()
    launch(mask); ctx.sync(); ms = (time.perf_counter() - t0) * 1e3
    c = counters() - c0
    print(f"mask {mask:2d}: {ms:6.2f} ms  CTAs {int(c[NW])}  avg kcycles per role:", [round(v / max(c[NW], 1) / 1e3, 1) for v in c[:NW]])
