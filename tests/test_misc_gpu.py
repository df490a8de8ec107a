"""GPU parity (batched form, n > 1) for the small table entries next to the hot path: ssim_4x4x2_core / ssim_end_4 (float,
BIT-exact: every operation individually rounded in the reference's order), planeClipAndMax (10-bit), propagateCost incl. the
degenerate intraCost == 0 inputs (x86 conversion semantics)."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle
from util import pdtype, vp, vpo, ssz

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("depth", [8, 10])
def test_ssim_primitives_bit_exact(ctx, depth):
    R = oracle.ref(depth)
    R.ref_ssim_end4.restype = ctypes.c_float
    rng = np.random.default_rng(depth)
    n = 300
    hi = 16 * 2 * ((1 << depth) - 1)
    s0 = np.zeros((n, 5, 4), dtype=np.int32); s1 = np.zeros((n, 5, 4), dtype=np.int32)
    for arr in (s0, s1):
        arr[:, :, 0] = rng.integers(0, hi, (n, 5)); arr[:, :, 1] = rng.integers(0, hi, (n, 5))
        arr[:, :, 2] = rng.integers(0, hi * ((1 << depth) - 1) // 2, (n, 5)); arr[:, :, 3] = rng.integers(0, hi * ((1 << depth) - 1) // 4, (n, 5))
    widths = rng.integers(1, 5, n).astype(np.int32)
    d0, d1, dW, dO = ctx.to_device(s0), ctx.to_device(s1), ctx.to_device(widths), ctx.empty(n * 4)
    ctx.ssim_end4_dev(depth, d0, d1, dW, n, dO)
    got = dO.download(np.float32)
    exp = np.array([R.ref_ssim_end4(vp(s0[i]), vp(s1[i]), int(widths[i])) for i in range(n)], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    # ssim_4x4x2_core over a list of block pairs
    S = 80
    a = rng.integers(0, 1 << depth, S * 40).astype(pdtype(depth)); b = rng.integers(0, 1 << depth, S * 40).astype(pdtype(depth))
    off = (rng.integers(0, 30, n) * S + rng.integers(0, 60, n)).astype(np.int64)
    dA, dB, dOff, dS = ctx.to_device(a), ctx.to_device(b), ctx.to_device(off), ctx.empty(n * 32)
    ctx.ssim_4x4x2_dev(depth, dA, S, dB, S, dOff, dOff, n, dS)
    gs = dS.download(np.int32).reshape(n, 2, 4)
    for i in range(n):
        e = np.zeros((2, 4), dtype=np.int32)
        R.ref_ssim_core(vpo(a, int(off[i])), ssz(S), vpo(b, int(off[i])), ssz(S), vp(e))
        assert np.array_equal(gs[i], e), i
    for bfr in (d0, d1, dW, dO, dA, dB, dOff, dS):
        bfr.free()


def test_plane_clip_and_max_10bit(ctx):
    R = oracle.ref(10)
    rng = np.random.default_rng(4)
    S, Wd, Hh = 208, 200, 37
    src = rng.integers(0, 1024, S * Hh).astype(np.uint16)
    e = src.copy(); esum = ctypes.c_uint64(0)
    emax = R.ref_plane_clip_max(vp(e), ssz(S), Wd, Hh, ctypes.byref(esum), 64, 940)
    assert emax >= 0
    dS, dO = ctx.to_device(src), ctx.empty(16)
    ctx.plane_clip_max_dev(10, dS, S, Wd, Hh, 64, 940, dO.ptr, dO.ptr + 8)
    out = dO.download(np.uint64)
    assert np.array_equal(dS.download(np.uint16), e) and int(out[0]) == esum.value and int(out[1] & 0xffffffff) == emax
    dS.free(); dO.free()


def test_propagate_cost_batch(ctx):
    R = oracle.ref(8)
    rng = np.random.default_rng(8)
    n = 5000
    pin = rng.integers(0, 65536, n).astype(np.uint16)
    intra = rng.integers(0, 40000, n).astype(np.int32)
    intra[::97] = 0                                               # division by zero -> x86 "integer indefinite"
    inter = rng.integers(0, 65536, n).astype(np.uint16)
    invq = rng.integers(1, 70000, n).astype(np.int32)
    exp = np.zeros(n, dtype=np.int32)
    R.ref_propagate_cost(vp(exp), vp(pin), vp(intra), vp(inter), vp(invq), ctypes.c_double(0.75 * 256), n)
    bufs = [ctx.to_device(x) for x in (pin, intra, inter, invq)]
    dD = ctx.empty(n * 4)
    ctx.propagate_cost_dev(dD, bufs[0], bufs[1], bufs[2], bufs[3], 0.75 * 256, n)
    assert np.array_equal(dD.download(np.int32), exp)
    for bfr in bufs + [dD]:
        bfr.free()


# ---- lookahead pre-analysis drivers and the weighted reference plane ---------------------------------------------
def _ref_pic(R, W, H, csp, depth, seed):
    from util import pdtype
    rng = np.random.default_rng(seed)
    dt = pdtype(depth)
    hs, vs = int(csp in (1, 2)), int(csp == 1)
    y = rng.integers(0, 1 << depth, (H, W)).astype(dt)
    y[: H // 2] = (y[: H // 2] >> 3) + (60 << (depth - 8))          # a flat-ish half: small energies next to large ones
    cb = rng.integers(0, 1 << depth, (H >> vs, W >> hs)).astype(dt) if csp else None
    cr = rng.integers(0, 1 << depth, (H >> vs, W >> hs)).astype(dt) if csp else None
    R.ref_pic_create.restype = ctypes.c_void_p
    R.ref_pic_buffer.restype = ctypes.c_void_p
    vp = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
    h = ctypes.c_void_p(R.ref_pic_create(W, H, csp, vp(y), vp(cb), vp(cr), ctypes.c_ssize_t(W), ctypes.c_ssize_t(W >> hs)))
    g = (ctypes.c_int64 * 8)()
    R.ref_pic_geometry(h, g)
    return h, [int(v) for v in g]


@pytest.mark.parametrize("depth,csp,qg,W,H", [(8, 1, 16, 200, 120), (8, 1, 8, 136, 72), (8, 3, 16, 96, 64), (8, 0, 16, 128, 96), (10, 1, 16, 200, 120), (10, 2, 8, 96, 64)])
def test_aq_energy_matches_reference(ctx, depth, csp, qg, W, H):
    """acEnergyCu over the block loop of calcAdaptiveQuantFrame (slicetype.cpp:49-86,252-275,519-523) vs the reference's own
    LookaheadTLD::acEnergyCu on its own padded PicYuv; picture sizes that are not multiples of the block size read the padding."""
    import oracle
    from util import pdtype
    R = oracle.ref(depth)
    assert R is not None
    h, (S, Sc, mx, my, cmx, cmy, rows, rowsC) = _ref_pic(R, W, H, csp, depth, 31 + qg)
    ct = ctypes.c_uint8 if depth == 8 else ctypes.c_uint16
    item = 2 if depth > 8 else 1
    grab = lambda plane, n: np.ctypeslib.as_array(ctypes.cast(R.ref_pic_buffer(h, plane), ctypes.POINTER(ct)), shape=(n,)).copy()
    dY = ctx.to_device(grab(0, S * rows))
    dCb = ctx.to_device(grab(1, Sc * rowsC)) if csp else None
    dCr = ctx.to_device(grab(2, Sc * rowsC)) if csp else None
    bx, by = (W + qg - 1) // qg, (H + qg - 1) // qg
    want = np.zeros(bx * by, dtype=np.uint32); wp = np.zeros(6, dtype=np.uint64)
    R.ref_aq_energy(h, qg, ctypes.c_void_p(want.ctypes.data), ctypes.c_void_p(wp.ctypes.data))
    dE, dW = ctx.empty(bx * by * 4), ctx.empty(48)
    ctx.aq_energy_dev(depth, csp, qg, dY.ptr + (my * S + mx) * item, S, dCb.ptr + (cmy * Sc + cmx) * item if csp else None,
                      dCr.ptr + (cmy * Sc + cmx) * item if csp else None, Sc, W, H, dE, dW)
    assert np.array_equal(dE.download(np.uint32), want)
    assert np.array_equal(dW.download(np.uint64), wp)
    assert len(set(want.tolist())) > bx * by // 2
    R.ref_pic_destroy(h)
    for b in (dY, dCb, dCr, dE, dW):
        if b is not None:
            b.free()


@pytest.mark.parametrize("depth,W,H,wt,off,den", [(8, 200, 120, 70, -3, 6), (8, 192, 128, 61, 5, 6), (8, 130, 70, 128, 0, 7), (10, 200, 120, 70, -3, 6)])
def test_apply_weight_matches_reference(ctx, depth, W, H, wt, off, den):
    """MotionReference::init + applyWeight over all rows (reference.cpp:51-185): the whole padded weighted luma plane, margins
    included, equals the reference's weightBuffer (a height that is not a multiple of the CTU exercises the partial last row)."""
    import oracle
    R = oracle.ref(depth)
    assert R is not None
    h, (S, Sc, mx, my, cmx, cmy, rows, rowsC) = _ref_pic(R, W, H, 0, depth, 77)
    ct = ctypes.c_uint8 if depth == 8 else ctypes.c_uint16
    item = 2 if depth > 8 else 1
    src = np.ctypeslib.as_array(ctypes.cast(R.ref_pic_buffer(h, 0), ctypes.POINTER(ct)), shape=(S * rows,)).copy()
    want = np.zeros_like(src)
    assert R.ref_apply_weight(h, wt, off, den, ctypes.c_void_p(want.ctypes.data)) == 0
    dS, dD = ctx.to_device(src), ctx.to_device(np.zeros_like(src))
    o = (my * S + mx) * item
    ctx.apply_weight_dev(depth, dS.ptr + o, dD.ptr + o, S, W, H, mx, my, wt, off, den)
    got = dD.download(src.dtype).reshape(rows, S)
    want = want.reshape(rows, S)
    # the reference allocates numCUinHeight * maxCU + 2 * marginY rows but fills picHeight + 2 * marginY of them
    # ... and of each row the columns [0, marginX + width + marginX): what lies between the right margin and the row pitch is never written
    filled, cols = slice(0, H + 2 * my), slice(0, W + 2 * mx)
    assert np.array_equal(got[filled, cols], want[filled, cols]), int(np.count_nonzero(got[filled, cols] != want[filled, cols]))
    R.ref_pic_destroy(h); dS.free(); dD.free()


def test_extend_row_border(ctx):
    """primitives.extendRowBorder = extendCURowColBorder (ipfilter.cpp:59-77): left / right margins of every row replicated, nothing else touched."""
    rng = np.random.default_rng(3)
    W, H, M = 70, 9, 24
    S = W + 2 * M + 8
    buf = rng.integers(0, 256, (H + 2) * S).astype(np.uint8)
    want = buf.copy().reshape(H + 2, S)
    for y in range(1, H + 1):
        want[y, :M] = want[y, M]; want[y, M + W:M + W + M] = want[y, M + W - 1]
    d = ctx.to_device(buf)
    ctx.L.x265b200_extend_border_dev(ctx.h, 8, ctypes.c_void_p(d.ptr + S + M), ctypes.c_int64(S), W, H, M, 0)
    assert np.array_equal(d.download(np.uint8).reshape(H + 2, S), want)
    d.free()
