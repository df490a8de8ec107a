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
