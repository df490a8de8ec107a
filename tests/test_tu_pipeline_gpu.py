"""GPU parity for the fused residual pipeline: sub_ps -> DCT/DST -> quant ->
dequant -> (zero | DC fill | IDCT/IDST) -> add_ps -> sse_pp per TU must equal the reference's own table entries chained the way
Quant::transformNxN / invtransformNxN chain them (oracle/ref_capi.cpp ref_tu_pipeline, quant.cpp:397-480, :543-605)."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle
from util import pdtype, vp

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu

QUANT_SCALES = [26214, 23302, 20560, 18396, 16384, 14564]          # scalinglist.cpp s_quantScales
INV_QUANT_SCALES = [40, 45, 51, 57, 64, 72]                        # s_invQuantScales


def make_planes(depth, N, bx, by, rng, stride_pad=16):
    """fenc / pred planes whose TUs cover the interesting cases: identical (numSig 0), constant offset (DC only),
    small and large noise."""
    W, H = bx * N, by * N
    S = W + stride_pad
    pmax = (1 << depth) - 1
    fenc = rng.integers(0, pmax + 1, (H, S)).astype(np.int64)
    k = np.ones(3) / 3.0
    sm = fenc.astype(np.float64)
    for ax in (0, 1):
        sm = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), ax, sm)
    fenc = np.clip(np.rint(sm), 0, pmax).astype(np.int64)
    pred = fenc.copy()
    for ty in range(by):
        for tx in range(bx):
            kind = int(rng.integers(0, 5))
            blk = pred[ty * N:(ty + 1) * N, tx * N:(tx + 1) * N]
            if kind == 1:
                blk += int(rng.integers(-12, 13)) << (depth - 8)
            elif kind == 2:
                blk += rng.integers(-3, 4, blk.shape) << (depth - 8)
            elif kind == 3:
                blk += rng.integers(-60, 61, blk.shape) << (depth - 8)
            elif kind == 4:
                blk[:] = rng.integers(0, pmax + 1, blk.shape)
    pred = np.clip(pred, 0, pmax)
    dt = pdtype(depth)
    return fenc.astype(dt).ravel(), pred.astype(dt).ravel(), S


def run_case(ctx, depth, sizeIdx, qp, useDST, scaling, intra_slice, seed, bx=9, by=5):
    R = oracle.ref(depth)
    N = 4 << sizeIdx
    log2N = sizeIdx + 2
    rng = np.random.default_rng(seed)
    fenc, pred, S = make_planes(depth, N, bx, by, rng)
    n = bx * by
    qpb = qp + 6 * (depth - 8)
    per, rem = qpb // 6, qpb % 6
    ts = 15 - depth - log2N
    qbits = 14 + per + ts
    add = (171 if intra_slice else 85) << (qbits - 9)
    dq_shift = 20 - 14 - ts
    if scaling:
        qc = (QUANT_SCALES[rem] * 16 // rng.integers(8, 48, N * N)).astype(np.int32)
        dqc = (INV_QUANT_SCALES[rem] * rng.integers(8, 48, N * N)).astype(np.int32)
        scale_or_per = per
    else:
        qc = np.full(N * N, QUANT_SCALES[rem], dtype=np.int32)
        dqc = None
        scale_or_per = INV_QUANT_SCALES[rem] << per
    # reference
    e_rec = np.zeros_like(fenc)
    e_coef = np.zeros(n * N * N, dtype=np.int16)
    e_ns = np.zeros(n, dtype=np.uint32)
    e_sse = np.zeros(n, dtype=np.uint64)
    R.ref_tu_pipeline(sizeIdx, int(useDST), vp(fenc), ctypes.c_ssize_t(S), vp(pred), ctypes.c_ssize_t(S), vp(e_rec), ctypes.c_ssize_t(S), bx, by,
                      vp(qc), qbits, add, vp(dqc) if dqc is not None else None, scale_or_per, dq_shift, vp(e_coef), vp(e_ns), vp(e_sse), 4)
    # backend
    dF, dP, dR = ctx.to_device(fenc), ctx.to_device(pred), ctx.to_device(np.zeros_like(fenc))
    dQ = ctx.to_device(qc)
    dD = ctx.to_device(dqc) if dqc is not None else None
    dC, dN, dS = ctx.empty(e_coef.nbytes), ctx.empty(n * 4), ctx.empty(n * 8)
    ctx.tu_pipeline_dev(sizeIdx, depth, int(useDST), dF, S, dP, S, dR, S, bx, by, dQ, qbits, add, dD, scale_or_per, dq_shift, dC, dN, dS)
    g_coef, g_ns, g_sse, g_rec = dC.download(np.int16), dN.download(np.uint32), dS.download(np.uint64), dR.download(fenc.dtype)
    for b in (dF, dP, dR, dQ, dC, dN, dS) + ((dD,) if dD else ()):
        b.free()
    tag = "depth %d N %d qp %d dst %d scaling %d" % (depth, N, qp, useDST, scaling)
    assert np.array_equal(g_ns, e_ns), tag
    assert np.array_equal(g_coef, e_coef), tag
    W = bx * N
    assert np.array_equal(g_rec.reshape(-1, S)[:, :W], e_rec.reshape(-1, S)[:, :W]), tag
    assert np.array_equal(g_sse, e_sse), tag
    return e_ns


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("sizeIdx", [0, 1, 2, 3])
def test_tu_pipeline_matches_reference_chain(ctx, depth, sizeIdx):
    seen = set()
    for qp, intra in ((22, True), (30, False), (38, False)):
        ns = run_case(ctx, depth, sizeIdx, qp, False, False, intra, seed=50 + qp + sizeIdx)
        seen |= {0 if v == 0 else (1 if v == 1 else 2) for v in ns}
    assert seen == {0, 1, 2}          # zero-residual, DC-only and full-inverse TUs were all exercised


@pytest.mark.parametrize("depth", [8, 10])
def test_tu_pipeline_dst_and_scaling_lists(ctx, depth):
    run_case(ctx, depth, 0, 27, True, False, True, seed=90)         # 4x4 luma intra: DST, no DC shortcut
    for sizeIdx in range(4):
        run_case(ctx, depth, sizeIdx, 33, False, True, False, seed=95 + sizeIdx)
        run_case(ctx, depth, sizeIdx, 8, False, True, False, seed=99 + sizeIdx)      # small per: the shift <= per branch of dequant_scaling


def test_tu_pipeline_odd_grid_and_unaligned(ctx):
    # odd TU counts (the second 8x8 of the last warp unit is absent) and a byte-aligned recon plane
    run_case(ctx, 8, 1, 30, False, False, False, seed=120, bx=7, by=3)
    run_case(ctx, 8, 3, 30, False, False, False, seed=121, bx=1, by=1)
    with pytest.raises(pkg.X265B200Error):
        ctx.tu_pipeline_dev(1, 8, 1, 1, 64, 1, 64, 1, 64, 1, 1, 1, 14, 0, None, 40, 4, 1, 1, 1)      # DST only for 4x4
