"""GPU parity: block-compare kernels (through the C ABI) vs the oracle on TestBench-shaped
fixtures -- all 25 PU sizes x {sad, satd}, all CU sizes x {sa8d, sse_pp, sse_ss, ssd_s}, 8- and
10-bit, random / all-min / all-max buffers, unaligned offsets and strides.  Bit-exact."""
import importlib

import numpy as np
import pytest

from util import (CU_SIZES, LUMA_PU_SIZES, STRIDE, block_offsets, orc_cmp, pixel_buffers, ref_cmp, short_buffers)

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu

KIND = {"sad": pkg.CMP_SAD, "satd": pkg.CMP_SATD, "sa8d": pkg.CMP_SA8D, "sa8d8": pkg.CMP_SA8D8,
        "sse_pp": pkg.CMP_SSE_PP, "sse_ss": pkg.CMP_SSE_SS, "ssd_s": pkg.CMP_SSD_S}


def _check(ctx, kind, depth, w, h, A, sa, B, sb, offA, offB):
    got = ctx.pixelcmp_host(KIND[kind], depth, w, h, A, sa, B, sb, offA, offB)
    exp = orc_cmp(kind, depth, w, h, A, sa, B, sb, offA, offB)
    assert [int(v) for v in got] == exp, (kind, depth, w, h)
    r = ref_cmp(kind, depth, w, h, A, sa, B, sb, offA, offB)
    if r is not None:
        assert [int(v) for v in got] == r, ("vs reference", kind, depth, w, h)


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind", ["sad", "satd"])
def test_pu_compare_all_sizes(ctx, depth, kind):
    bufs = pixel_buffers(depth)
    offs = block_offsets()
    for (w, h) in LUMA_PU_SIZES:
        for ia, ib in [(0, 0), (0, 1), (0, 2), (1, 2)]:
            _check(ctx, kind, depth, w, h, bufs[ia], STRIDE, bufs[ib], STRIDE, offs, offs[::-1].copy())


@pytest.mark.parametrize("depth", [8, 10])
def test_pu_compare_unaligned_ref(ctx, depth):
    """reference block at odd offsets and stride 59 (pixelharness.cpp:150 uses FENC_STRIDE - 5)."""
    bufs = pixel_buffers(depth, seed=99)
    offa = block_offsets()
    offb = offa + np.arange(len(offa)) % 7 + 1
    for (w, h) in LUMA_PU_SIZES:
        for kind in ("sad", "satd"):
            _check(ctx, kind, depth, w, h, bufs[0], STRIDE, bufs[0][::-1].copy(), 59, offa, offb)


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind", ["sa8d", "sse_pp"])
def test_cu_compare(ctx, depth, kind):
    bufs = pixel_buffers(depth, seed=7)
    offs = block_offsets()
    for s in CU_SIZES:
        for ia, ib in [(0, 0), (0, 1), (0, 2), (1, 2)]:
            _check(ctx, kind, depth, s, s, bufs[ia], STRIDE, bufs[ib], STRIDE, offs, offs[::-1].copy() + 3)


@pytest.mark.parametrize("depth", [8, 10])
def test_sa8d8_chroma_aliases(ctx, depth):
    bufs = pixel_buffers(depth, seed=8)
    offs = block_offsets()
    for (w, h) in [(8, 8), (8, 16), (16, 8), (32, 32)]:
        _check(ctx, "sa8d8", depth, w, h, bufs[0], STRIDE, bufs[0], 61, offs, offs + 5)
    for (w, h) in [(16, 32), (32, 64), (32, 16)]:       # sa8d16<16,32> etc. (pixel.cpp:1324-1325)
        _check(ctx, "sa8d", depth, w, h, bufs[0], STRIDE, bufs[0], 61, offs, offs + 5)


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind", ["sse_ss", "ssd_s"])
def test_short_compare(ctx, depth, kind):
    for full in (False, True):
        bufs = short_buffers(depth, full_range=full)
        offs = block_offsets()
        for s in CU_SIZES:
            for ia, ib in [(0, 0), (0, 1), (1, 2)]:
                _check(ctx, kind, depth, s, s, bufs[ia], STRIDE, bufs[ib], STRIDE, offs, offs[::-1].copy() + 1)


@pytest.mark.parametrize("depth", [8, 10])
def test_sad_x3_x4(ctx, depth):
    """pu[].sad_x3 / sad_x4 (pixel.cpp:74-119): cached 64-stride fenc vs K refs at stride 59."""
    import oracle
    import ctypes
    from util import vpo, ssz
    bufs = pixel_buffers(depth, seed=21)
    fenc_src, ref = bufs[0], bufs[0][::-1].copy()
    n = 12
    L = oracle.orc()
    dRef = ctx.to_device(ref)
    for (w, h) in LUMA_PU_SIZES:
        for K in (3, 4):
            fenc = np.zeros(n * 64 * 64, dtype=fenc_src.dtype)
            for i in range(n):
                blk = fenc_src[32 * i: 32 * i + 64 * 64]
                fenc[i * 4096:(i + 1) * 4096] = blk
            roff = (np.arange(n * K, dtype=np.int64) * 13 + 3) % 2000
            dF, dO = ctx.to_device(fenc), ctx.to_device(roff)
            dR = ctx.empty(n * K * 4)
            ctx.sad_xn_dev(depth, K, w, h, dF, 4096, dRef, 59, dO, n, dR)
            got = dR.download(np.int32).reshape(n, K)
            for i in range(n):
                ptrs = (ctypes.c_void_p * K)(*[vpo(ref, roff[i * K + k]).value for k in range(K)])
                res = (ctypes.c_int32 * K)()
                L.orc_sad_xn(depth, K, w, h, vpo(fenc, i * 4096), ptrs, ssz(59), res)
                assert list(res) == list(got[i]), (w, h, K, i)
            for b in (dF, dO, dR):
                b.free()
    dRef.free()


def test_grid_mode_with_mv_field(ctx):
    """grid mode = cost-at-predictor for every PU of a frame (motion.cpp:771-796 first step)."""
    rng = np.random.default_rng(5)
    W, H, S, pad = 256, 128, 320, 32
    cur = rng.integers(0, 256, (H + 2 * pad) * S, dtype=np.int64).astype(np.uint8)
    ref = rng.integers(0, 256, (H + 2 * pad) * S, dtype=np.int64).astype(np.uint8)
    base = pad * S + pad
    for (w, h) in [(8, 8), (16, 16), (32, 32), (64, 64), (32, 16)]:
        cols, rows = W // w, H // h
        n = cols * rows
        mv = rng.integers(-16, 17, (n, 2), dtype=np.int64).astype(np.int16)
        dC, dR, dMv, dOut = ctx.to_device(cur[base:]), ctx.to_device(ref), ctx.to_device(mv), ctx.empty(n * 4)
        # A is passed pre-offset; B via a device pointer offset
        ctx.pixelcmp_dev(pkg.CMP_SAD, 8, w, h, dC, S, dR.ptr + base, S, None, None, n, dOut, dMv=dMv, grid_cols=cols)
        got = dOut.download(np.int32)
        offA = np.array([base + (i // cols) * h * S + (i % cols) * w for i in range(n)], dtype=np.int64)
        offB = offA + mv[:, 0].astype(np.int64) + mv[:, 1].astype(np.int64) * S
        exp = orc_cmp("sad", 8, w, h, cur, S, ref, S, offA, offB)
        assert list(map(int, got)) == exp, (w, h)
        for b in (dC, dR, dMv, dOut):
            b.free()


@pytest.mark.parametrize("with_mv,depth", [(False, 8), (True, 8), (False, 10), (True, 10), (True, 12)])
def test_sad_pyramid_matches_per_level_sad(ctx, with_mv, depth):
    """one streaming pass = sad<8,8>..sad<64,64> of every 2Nx2N PU (incl. an odd CTU count and unaligned MVs), 8- and 16-bit samples."""
    from util import pdtype
    rng = np.random.default_rng(11)
    ctuCols, ctuRows, pad = 5, 2, 48
    W, H = ctuCols * 64, ctuRows * 64
    S = W + 2 * pad + 16 - (W + 2 * pad) % 16
    px = 2 if depth > 8 else 1
    cur = rng.integers(0, 1 << depth, (H + 2 * pad) * S, dtype=np.int64).astype(pdtype(depth))
    NREF = 2
    refs = [rng.integers(0, 1 << depth, (H + 2 * pad) * S, dtype=np.int64).astype(pdtype(depth)) for _ in range(NREF)]
    base = pad * S + 48
    mv = rng.integers(-20, 21, (NREF, ctuRows * ctuCols, 2), dtype=np.int64).astype(np.int16) if with_mv else None
    dC = ctx.to_device(cur)
    dR = [ctx.to_device(r) for r in refs]
    dPtrs = ctx.to_device(np.array([d.ptr + base * px for d in dR], dtype=np.int64))
    dMv = ctx.to_device(mv) if with_mv else None
    outs = {s: ctx.empty(NREF * (W // s) * (H // s) * 4) for s in (8, 16, 32, 64)}
    ctx.sad_pyramid_dev(depth, dC.ptr + base * px, S, dPtrs, NREF, S, ctuCols, ctuRows, dMv, outs[8], outs[16], outs[32], outs[64])
    for s in (8, 16, 32, 64):
        got = outs[s].download(np.int32).reshape(NREF, -1)
        cols = W // s
        for r in range(NREF):
            offA, offB = [], []
            for i in range(cols * (H // s)):
                bx, by = i % cols, i // cols
                ctu = (by * s // 64) * ctuCols + (bx * s // 64)
                mx, my = (int(mv[r][ctu][0]), int(mv[r][ctu][1])) if with_mv else (0, 0)
                offA.append(base + by * s * S + bx * s)
                offB.append(base + (by * s + my) * S + bx * s + mx)
            exp = orc_cmp("sad", depth, s, s, cur, S, refs[r], S, np.array(offA), np.array(offB))
            assert list(map(int, got[r])) == exp, (s, with_mv, depth, r)


@pytest.mark.parametrize("depth,ctuCols,ctuRows,nref,ngroups", [(8, 5, 2, 2, 3), (8, 9, 3, 3, 4), (10, 5, 2, 2, 3), (10, 3, 3, 1, 2)])
def test_sad_stream_pool_matches_per_level_sad(ctx, depth, ctuCols, ctuRows, nref, ngroups):
    """the streaming form over a frame pool (TMA ring, many groups per launch) = pu[8x8..64x64].sad at zero displacement of every
    2Nx2N PU, per group and reference; odd CTU counts leave the last 256-byte tile partly outside the picture."""
    from util import pdtype
    rng = np.random.default_rng(13 + depth)
    padX, padY = 96, 80
    W, H = ctuCols * 64, ctuRows * 64
    S, R = W + 2 * padX, H + 2 * padY
    item = 2 if depth > 8 else 1
    NF = 6
    pitch = S * R + 64                                            # frames need not be back to back
    pool = rng.integers(0, 1 << depth, NF * pitch, dtype=np.int64).astype(pdtype(depth))
    groups = np.zeros(ngroups, dtype=pkg.SAD_GROUP)
    for g in range(ngroups):
        idx = rng.permutation(NF)
        groups[g]["cur"] = idx[0]
        groups[g]["ref"][:nref] = idx[1:1 + nref]
    origin = padY * S + padX
    dP = ctx.to_device(pool)
    outs = {s: ctx.empty(ngroups * nref * (W // s) * (H // s) * 4) for s in (8, 16, 32, 64)}
    ctx.sad_stream_dev(depth, dP.ptr + origin * item, pitch, S, padX, padY, R, NF, ctuCols, ctuRows, groups, nref, outs[8], outs[16], outs[32], outs[64])
    for s in (8, 16, 32, 64):
        got = outs[s].download(np.int32).reshape(ngroups, nref, -1)
        cols = W // s
        off = np.array([origin + (i // cols) * s * S + (i % cols) * s for i in range(cols * (H // s))], dtype=np.int64)
        for g in range(ngroups):
            cur = pool[int(groups[g]["cur"]) * pitch:][:S * R]
            for r in range(nref):
                ref = pool[int(groups[g]["ref"][r]) * pitch:][:S * R]
                exp = orc_cmp("sad", depth, s, s, cur, S, ref, S, off, off)
                assert list(map(int, got[g, r])) == exp, (s, g, r)
    dP.free()
    for b in outs.values():
        b.free()
