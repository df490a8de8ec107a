"""debug driver: one small me_frame call (used under compute-sanitizer)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("x265-yuuki-asuna_b200")
from me_util import synth_pair
mask = int(os.environ.get("MF_MASK", "15"))
ctx = pkg.Ctx(0)
ctuCols, ctuRows = 2, 1
W, H, pad = ctuCols * 64, ctuRows * 64, 144
cur, ref, S, origin = synth_pair(W, H, pad, depth=8, seed=1)
dC, dR = ctx.to_device(cur), ctx.to_device(ref)
n = sum(ctuCols * ctuRows * (1 << l) ** 2 for l in range(4) if mask & (1 << l))
dOut = ctx.empty(max(n, 1) * 12)
ctx.me_frame_dev(8, dC.ptr + origin, S, [dR.ptr + origin], S, pad, pad, H + 2 * pad, ctuCols, ctuRows, mask, None, pkg.ME_HEX, 2, 57, pkg.lambda_for_qp(30, 8), dOut)
ctx.sync()
print("mask", mask, "ok", dOut.download(np.int32)[:12])
