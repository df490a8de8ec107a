"""CPU: pin the restatement (oracle/x265_oracle.c) against the reference itself
(oracle/_ref, compiled from /root/reference) on TestBench-shaped fixtures.  Runs without a GPU."""
import numpy as np
import pytest

import oracle
from util import (LUMA_PU_SIZES, CU_SIZES, STRIDE, block_offsets, orc_cmp, pixel_buffers, ref_cmp, short_buffers)

needs_ref = pytest.mark.skipif(not oracle.have_ref(8), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind", ["sad", "satd"])
def test_pu_cmp_matches_reference(depth, kind):
    bufs = pixel_buffers(depth)
    offs = block_offsets()
    for (w, h) in LUMA_PU_SIZES:
        for ia, ib in [(0, 0), (0, 1), (0, 2), (1, 2), (2, 1)]:
            offb = offs[::-1].copy()
            got = orc_cmp(kind, depth, w, h, bufs[ia], STRIDE, bufs[ib], STRIDE, offs, offb)
            exp = ref_cmp(kind, depth, w, h, bufs[ia], STRIDE, bufs[ib], STRIDE, offs, offb)
            assert got == exp, (kind, w, h, ia, ib)


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind", ["sa8d", "sse_pp", "sa8d8"])
def test_cu_cmp_matches_reference(depth, kind):
    bufs = pixel_buffers(depth)
    offs = block_offsets()
    sizes = [(s, s) for s in CU_SIZES] if kind != "sa8d8" else [(8, 8), (8, 16)]
    for (w, h) in sizes:
        for ia, ib in [(0, 0), (0, 1), (0, 2), (1, 2)]:
            offb = offs[::-1].copy()
            got = orc_cmp(kind, depth, w, h, bufs[ia], STRIDE, bufs[ib], STRIDE, offs, offb)
            exp = ref_cmp(kind, depth, w, h, bufs[ia], STRIDE, bufs[ib], STRIDE, offs, offb)
            assert got == exp, (kind, w, h, ia, ib)


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind", ["sse_ss", "ssd_s"])
def test_short_cmp_matches_reference(depth, kind):
    for full in (False, True):
        bufs = short_buffers(depth, full_range=full)
        offs = block_offsets()
        for s in CU_SIZES:
            for ia, ib in [(0, 0), (0, 1), (1, 2), (0, 2)]:
                offb = offs[::-1].copy()
                got = orc_cmp(kind, depth, s, s, bufs[ia], STRIDE, bufs[ib], STRIDE, offs, offb)
                exp = ref_cmp(kind, depth, s, s, bufs[ia], STRIDE, bufs[ib], STRIDE, offs, offb)
                assert got == exp, (kind, s, ia, ib, full)


def test_known_answers_ramp():
    """KATs recorded from the reference C table in SURVEY.md 8c (ramp fixture, stride 64)."""
    i = np.arange(64 * 64)
    a = ((7 * i + 3) & 255).astype(np.uint8)
    b = ((13 * i + 5) & 255).astype(np.uint8)
    z = np.zeros(1, dtype=np.int64)
    assert orc_cmp("sad", 8, 16, 16, a, 64, b, 64, z, z) == [22800]
    assert orc_cmp("satd", 8, 16, 16, a, 64, b, 64, z, z) == [37184]
    assert orc_cmp("sa8d", 8, 16, 16, a, 64, b, 64, z, z) == [25984]
    assert orc_cmp("sse_pp", 8, 16, 16, a, 64, b, 64, z, z) == [2960896]


# ---- transforms / interpolation / intra restatements vs the compiled reference ------------------------------
import ctypes  # noqa: E402
from util import pdtype, vpo, ssz  # noqa: E402


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
def test_transform_restatement_matches_reference(depth):
    R, O = oracle.ref(depth), oracle.orc()
    rng = np.random.default_rng(5)
    for full in (False, True):
        for idx, N in ((0, 4), (1, 8), (2, 16), (3, 32), (4, 4)):
            lim = 32767 if full else (1 << depth) - 1
            src = rng.integers(-lim - (1 if full else 0), lim + 1, N * (N + 3), dtype=np.int64).astype(np.int16)
            a, b = np.empty(N * N, dtype=np.int16), np.empty(N * N, dtype=np.int16)
            R.ref_dct(idx, vpo(src, 0), vpo(a, 0), ssz(N + 3)); O.orc_dct(depth, idx, vpo(src, 0), vpo(b, 0), ssz(N + 3))
            assert np.array_equal(a, b), ("dct", idx, full)
            c, d = np.zeros(N * (N + 2), dtype=np.int16), np.zeros(N * (N + 2), dtype=np.int16)
            R.ref_idct(idx, vpo(src, 0), vpo(c, 0), ssz(N + 2)); O.orc_idct(depth, idx, vpo(src, 0), vpo(d, 0), ssz(N + 2))
            assert np.array_equal(c, d), ("idct", idx, full)
    for numCoeff in (16, 64, 256, 1024):
        coef = rng.integers(-32768, 32768, numCoeff, dtype=np.int64).astype(np.int16)
        qc = rng.integers(1, 1 << 15, numCoeff, dtype=np.int64).astype(np.int32)
        qBits, add = 19, 171 << 10
        q1, q2 = np.empty(numCoeff, dtype=np.int16), np.empty(numCoeff, dtype=np.int16)
        u1, u2 = np.empty(numCoeff, dtype=np.int32), np.empty(numCoeff, dtype=np.int32)
        assert R.ref_quant(vpo(coef, 0), vpo(qc, 0), vpo(u1, 0), vpo(q1, 0), qBits, add, numCoeff) == O.orc_quant(vpo(coef, 0), vpo(qc, 0), vpo(u2, 0), vpo(q2, 0), qBits, add, numCoeff)
        assert np.array_equal(q1, q2) and np.array_equal(u1, u2)
        assert R.ref_nquant(vpo(coef, 0), vpo(qc, 0), vpo(q1, 0), qBits, add, numCoeff) == O.orc_nquant(vpo(coef, 0), vpo(qc, 0), vpo(q2, 0), qBits, add, numCoeff)
        assert np.array_equal(q1, q2)
        R.ref_dequant_normal(vpo(coef, 0), vpo(q1, 0), numCoeff, 40 << 3, 6); O.orc_dequant_normal(vpo(coef, 0), vpo(q2, 0), numCoeff, 40 << 3, 6)
        assert np.array_equal(q1, q2)
        for per, shift in ((0, 3), (9, 2)):
            R.ref_dequant_scaling(vpo(coef, 0), vpo(qc, 0), vpo(q1, 0), numCoeff, per, shift); O.orc_dequant_scaling(vpo(coef, 0), vpo(qc, 0), vpo(q2, 0), numCoeff, per, shift)
            assert np.array_equal(q1, q2), (per, shift)


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
def test_interp_restatement_matches_reference(depth):
    R, O = oracle.ref(depth), oracle.orc()
    rng = np.random.default_rng(6)
    dt = pdtype(depth)
    for part, (w, h) in enumerate(LUMA_PU_SIZES):
        S = w + 19
        pix = rng.integers(0, 1 << depth, (h + 16) * S + 32, dtype=np.int64).astype(dt)
        sht = rng.integers(-9000, 9000, (h + 16) * S + 32, dtype=np.int64).astype(np.int16)
        off = 8 * S + 8
        for kind in range(8):
            src = sht if kind in (4, 5) else pix
            dsh = kind in (1, 3, 5, 7)
            rows = h + (7 if kind == 1 else 0)
            for cx, cy in ((1, 2), (2, 3), (3, 1)):
                a = np.zeros((rows + 1) * (w + 2), dtype=np.int16 if dsh else dt)
                b = a.copy()
                R.ref_interp(kind, -1, part, vpo(src, off), ssz(S), vpo(a, 0), ssz(w + 2), cx, cy if kind == 6 else int(kind == 1))
                O.orc_interp(depth, kind, 8, w, h, vpo(src, off), ssz(S), vpo(b, 0), ssz(w + 2), cx, cy, int(kind == 1))
                assert np.array_equal(a, b), (kind, w, h, cx, cy)
    # chroma 4-tap (4:2:0 shapes)
    for part, (lw, lh) in enumerate(LUMA_PU_SIZES):
        w, h = lw // 2, lh // 2
        S = w + 11
        pix = rng.integers(0, 1 << depth, (h + 8) * S + 16, dtype=np.int64).astype(dt)
        sht = rng.integers(-9000, 9000, (h + 8) * S + 16, dtype=np.int64).astype(np.int16)
        off = 4 * S + 4
        for kind in (0, 1, 2, 3, 4, 5):
            if not R.ref_interp_available(kind, 1, part):
                continue
            src = sht if kind in (4, 5) else pix
            dsh = kind in (1, 3, 5)
            rows = h + (3 if kind == 1 else 0)
            for cx in (1, 4, 7):
                a = np.zeros((rows + 1) * (w + 2), dtype=np.int16 if dsh else dt)
                b = a.copy()
                R.ref_interp(kind, 1, part, vpo(src, off), ssz(S), vpo(a, 0), ssz(w + 2), cx, int(kind == 1))
                O.orc_interp(depth, kind, 4, w, h, vpo(src, off), ssz(S), vpo(b, 0), ssz(w + 2), cx, 0, int(kind == 1))
                assert np.array_equal(a, b), ("chroma", kind, w, h, cx)


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
def test_intra_restatement_matches_reference(depth):
    R, O = oracle.ref(depth), oracle.orc()
    rng = np.random.default_rng(7)
    dt = pdtype(depth)
    for log2N in (2, 3, 4, 5):
        N = 1 << log2N
        for trial in range(3):
            nb = rng.integers(0, 1 << depth, 4 * N + 1, dtype=np.int64).astype(dt) if trial < 2 else np.full(4 * N + 1, (1 << depth) - 1, dtype=dt)
            f1, f2 = np.zeros_like(nb), np.zeros_like(nb)
            R.ref_intra_filter(log2N - 2, vpo(nb, 0), vpo(f1, 0)); O.orc_intra_filter(depth, log2N, vpo(nb, 0), vpo(f2, 0))
            assert np.array_equal(f1, f2)
            for mode in range(35):
                for bf in (0, 1):
                    a = np.zeros(N * (N + 1), dtype=dt); b = a.copy()
                    R.ref_intra_pred(log2N - 2, mode, vpo(a, 0), ssz(N + 1), vpo(nb, 0), bf)
                    O.orc_intra_pred(depth, log2N, mode, bf, vpo(nb, 0), vpo(b, 0), ssz(N + 1))
                    assert np.array_equal(a, b), (log2N, mode, bf)


# ---- families added late in round 1: SEA support, in-loop filters, cuTree, the residual chain ------------------------------------
import ctypes  # noqa: E402

from util import pdtype, vp, vpo, ssz  # noqa: E402


@needs_ref
def test_ads_restatement():
    R, O = oracle.ref(8), oracle.orc()
    rng = np.random.default_rng(1)
    stride, width = 256, 100
    sums = rng.integers(0, 30000, stride * 24).astype(np.uint32)
    cost = rng.integers(0, 200, width).astype(np.uint16)
    for (w, h), kind in (((8, 8), 1), ((16, 4), 1), ((16, 8), 2), ((32, 64), 2), ((16, 16), 4), ((64, 64), 4), ((32, 24), 4)):
        part = LUMA_PU_SIZES.index((w, h))
        for trial in range(6):
            enc = rng.integers(0, 30000, 4).astype(np.int32)
            delta = int(rng.choice([8, 12, 8 * stride]))
            thresh = int(rng.integers(1000, 50000))
            off = int(rng.integers(0, stride * 10))
            m1, m2 = np.zeros(width + 8, dtype=np.int16), np.zeros(width + 8, dtype=np.int16)
            n1 = R.ref_ads(part, vp(enc), vpo(sums, off), delta, vp(cost), vp(m1), width, thresh)
            n2 = O.orc_ads(kind, w >> 1, vp(enc), vpo(sums, off), ssz(delta), vp(cost), vp(m2), width, thresh)
            assert n1 == n2 and np.array_equal(m1[:n1], m2[:n2]), (w, h, trial)


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
def test_sea_integral_restatement(depth):
    R, O = oracle.ref(depth), oracle.orc()
    W, H, padX, padY = 128, 64, 96, 80
    stride, rows = W + 2 * padX, H + 2 * padY
    rng = np.random.default_rng(2)
    plane = rng.integers(0, 1 << depth, stride * rows).astype(pdtype(depth))
    origin = padY * stride + padX
    a = [np.full(stride * rows, 0xAAAAAAAA, dtype=np.uint32) for _ in range(12)]
    b = [x.copy() for x in a]
    pa = (ctypes.c_void_p * 12)(*[x.ctypes.data + origin * 4 for x in a])
    pb = (ctypes.c_void_p * 12)(*[x.ctypes.data + origin * 4 for x in b])
    R.ref_sea_integrals(vpo(plane, origin), ssz(stride), padX, padY, H, pa)
    O.orc_sea_integral(depth, vpo(plane, origin), ssz(stride), padX, padY, H, pb)
    for k in range(12):
        assert np.array_equal(a[k], b[k]), k


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
def test_sao_and_deblock_restatement(depth):
    R, O = oracle.ref(depth), oracle.orc()
    rng = np.random.default_rng(3 + depth)
    S, rows = 200, 80
    base = rng.integers(0, 1 << depth, (rows // 4 + 1, S // 4 + 1))
    img = (np.kron(base, np.ones((4, 4), dtype=np.int64))[:rows, :S] + rng.integers(-2, 3, (rows, S))).clip(0, (1 << depth) - 1).astype(pdtype(depth)).ravel()
    for kind in range(6):
        for trial in range(5):
            r1, r2 = img.copy(), img.copy()
            buf = rng.integers(-1, 2, 200).astype(np.int8)
            b1, b2 = buf.copy(), buf.copy()
            offs = rng.integers(-7, 8, 32).astype(np.int8)
            width, height, startX = int(rng.integers(16, 65)), int(rng.integers(1, 9)), int(rng.integers(0, 2))
            ro = S * 3 + 5
            R.ref_sao_apply(kind, vpo(r1, ro), ssz(S), vpo(b1, 1), vpo(b1, 101), vp(offs), width, height, startX)
            O.orc_sao_apply(kind, depth, vpo(r2, ro), ssz(S), vpo(b2, 1), vpo(b2, 101), vp(offs), width, height, startX)
            assert np.array_equal(r1, r2) and np.array_equal(b1, b2), (kind, trial)
    diff = rng.integers(-50, 51, 64 * 64).astype(np.int16)
    for kind in (5, 0, 1, 3, 4):
        for trial in range(5):
            buf = rng.integers(-1, 2, 200).astype(np.int8)
            b1, b2 = buf.copy(), buf.copy()
            st = rng.integers(-100, 100, 32).astype(np.int32); ct = rng.integers(0, 100, 32).astype(np.int32)
            s1, c1, s2, c2 = st.copy(), ct.copy(), st.copy(), ct.copy()
            endX, endY = int(rng.integers(30, 64)), int(rng.integers(1, 64))
            ro = S * 2 + 3
            R.ref_sao_stats(kind, vp(diff), vpo(img, ro), ssz(S), vpo(b1, 2), vpo(b1, 102), endX, endY, vp(s1), vp(c1))
            O.orc_sao_stats(kind, depth, vp(diff), vpo(img, ro), ssz(S), vpo(b2, 2), vpo(b2, 102), endX, endY, vp(s2), vp(c2))
            assert np.array_equal(s1, s2) and np.array_equal(c1, c2) and np.array_equal(b1, b2), (kind, trial)
    for chroma in (0, 1):
        for trial in range(20):
            p1, p2 = img.copy(), img.copy()
            vert = trial & 1
            step, off = (S, 1) if vert else (1, S)
            a_, b_, c_ = int(rng.integers(0, 1 << (depth - 2))), int(rng.integers(-1, 40)), int(rng.integers(-1, 1))
            so = S * 20 + 50
            R.ref_deblock(chroma, vpo(p1, so), ssz(step), ssz(off), a_, b_, c_)
            O.orc_deblock(chroma, depth, vpo(p2, so), ssz(step), ssz(off), a_, b_, c_)
            assert np.array_equal(p1, p2), (chroma, trial)


@needs_ref
def test_propagate_cost_restatement():
    R, O = oracle.ref(8), oracle.orc()
    rng = np.random.default_rng(5)
    n = 4000
    pin = rng.integers(0, 65536, n).astype(np.uint16); intra = rng.integers(0, 50000, n).astype(np.int32); intra[::53] = 0
    inter = rng.integers(0, 65536, n).astype(np.uint16); invq = rng.integers(1, 70000, n).astype(np.int32)
    a, b = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    R.ref_propagate_cost(vp(a), vp(pin), vp(intra), vp(inter), vp(invq), ctypes.c_double(200.0), n)
    O.orc_propagate_cost(vp(b), vp(pin), vp(intra), vp(inter), vp(invq), ctypes.c_double(200.0), n)
    assert np.array_equal(a, b)


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
def test_tu_chain_restatement(depth):
    R, O = oracle.ref(depth), oracle.orc()
    rng = np.random.default_rng(6 + depth)
    for sizeIdx in range(4):
        N = 4 << sizeIdx
        for trial in range(12):
            fenc = rng.integers(0, 1 << depth, N * N).astype(pdtype(depth))
            kind = trial % 4
            noise = {0: 0, 1: 0, 2: 3, 3: 60}[kind] << (depth - 8)
            pred = fenc.astype(np.int64) + (int(rng.integers(-10, 11)) << (depth - 8) if kind == 1 else 0) + (rng.integers(-noise, noise + 1, N * N) if noise else 0)
            pred = pred.clip(0, (1 << depth) - 1).astype(pdtype(depth))
            useDST = int(sizeIdx == 0 and trial % 2)
            scaling = trial % 3 == 0
            qp = int(rng.integers(10, 45)) + 6 * (depth - 8)
            per, rem = qp // 6, qp % 6
            ts = 15 - depth - (sizeIdx + 2)
            qbits = 14 + per + ts
            add = 171 << (qbits - 9)
            qc = (np.full(N * N, [26214, 23302, 20560, 18396, 16384, 14564][rem]) * 16 // (rng.integers(8, 40, N * N) if scaling else 16)).astype(np.int32)
            dqc = ([40, 45, 51, 57, 64, 72][rem] * rng.integers(8, 40, N * N)).astype(np.int32) if scaling else None
            sop = per if scaling else [40, 45, 51, 57, 64, 72][rem] << per
            r1, r2 = np.zeros_like(fenc), np.zeros_like(fenc)
            c1, c2 = np.zeros(N * N, dtype=np.int16), np.zeros(N * N, dtype=np.int16)
            ns1, ss1, ss2 = np.zeros(1, dtype=np.uint32), np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64)
            R.ref_tu_pipeline(sizeIdx, useDST, vp(fenc), ssz(N), vp(pred), ssz(N), vp(r1), ssz(N), 1, 1, vp(qc), qbits, add,
                              vp(dqc) if scaling else None, sop, 6 - ts, vp(c1), vp(ns1), vp(ss1), 1)
            ns2 = O.orc_tu_chain(depth, sizeIdx, useDST, vp(fenc), ssz(N), vp(pred), ssz(N), vp(r2), ssz(N), vp(qc), qbits, add,
                                 vp(dqc) if scaling else None, sop, 6 - ts, vp(c2), vp(ss2))
            assert int(ns1[0]) == ns2 and np.array_equal(c1, c2) and np.array_equal(r1, r2) and ss1[0] == ss2[0], (depth, N, trial)
