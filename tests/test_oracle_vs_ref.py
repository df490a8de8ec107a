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
