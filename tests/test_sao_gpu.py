"""GPU parity, batched form (many CTU jobs per launch) of the in-loop filter entries: SAO offset application, SAO statistics and
the deblocking line filters against the reference's table entries called once per job (oracle/ref_capi.cpp); the single-call
form is covered by the reference's TestBench (tests/test_testbench_gpu.py)."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle
from util import pdtype, vp, vpo, ssz

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu


def _plane(rng, depth, S, rows):
    # smooth-ish content so that all five edge classes occur
    base = rng.integers(0, 1 << depth, (rows // 4 + 2, S // 4 + 2))
    p = np.kron(base, np.ones((4, 4), dtype=np.int64))[:rows, :S] + rng.integers(-2, 3, (rows, S))
    return np.clip(p, 0, (1 << depth) - 1).astype(pdtype(depth)).ravel()


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind", [pkg.SAO_E0, pkg.SAO_E1, pkg.SAO_E1_2ROWS, pkg.SAO_E2, pkg.SAO_E3, pkg.SAO_B0])
def test_sao_apply_batched(ctx, depth, kind):
    R = oracle.ref(depth)
    rng = np.random.default_rng(100 + kind + depth)
    S, rows, n = 400, 72, 24
    rec = _plane(rng, depth, S, rows)
    e_rec = rec.copy()
    buf = rng.integers(-1, 2, n * 160).astype(np.int8)
    e_buf = buf.copy()
    noff = 32 if kind == pkg.SAO_B0 else 5
    offs = rng.integers(-7, 8, n * 32).astype(np.int8)
    jobs = np.zeros(n, dtype=pkg.SAO_JOB)
    for i in range(n):
        # disjoint 64x~16 regions: column block i % 4, row band i // 4 (12 rows apart)
        jobs[i]["recOff"] = (i // 4) * 12 * S + (i % 4) * 96 + 1
        jobs[i]["buf0"], jobs[i]["buf1"], jobs[i]["offsetOff"] = i * 160 + 1, i * 160 + 81, i * 32
        jobs[i]["width"] = int(rng.integers(17, 65)); jobs[i]["height"] = int(rng.integers(1, 9)); jobs[i]["startX"] = int(rng.integers(0, 2))
        R.ref_sao_apply(kind, vpo(e_rec, int(jobs[i]["recOff"])), ssz(S), vpo(e_buf, int(jobs[i]["buf0"])), vpo(e_buf, int(jobs[i]["buf1"])),
                        vpo(offs, i * 32), int(jobs[i]["width"]), int(jobs[i]["height"]), int(jobs[i]["startX"]))
    dR, dB, dO, dJ = ctx.to_device(rec), ctx.to_device(buf), ctx.to_device(offs), ctx.to_device(jobs)
    ctx.sao_apply_dev(kind, depth, dR, S, dJ, n, dB, dO, 64)
    assert np.array_equal(dR.download(rec.dtype), e_rec), (kind, noff)
    assert np.array_equal(dB.download(np.int8), e_buf), kind
    for b in (dR, dB, dO, dJ):
        b.free()


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind", [pkg.SAO_BO, pkg.SAO_E0, pkg.SAO_E1, pkg.SAO_E2, pkg.SAO_E3])
def test_sao_stats_batched(ctx, depth, kind):
    R = oracle.ref(depth)
    rng = np.random.default_rng(200 + kind + depth)
    S, rows, n = 300, 210, 12
    rec = _plane(rng, depth, S, rows)
    diff = rng.integers(-40, 41, n * 64 * 64).astype(np.int16)
    buf = rng.integers(-1, 2, n * 160).astype(np.int8)
    e_buf = buf.copy()
    stats = rng.integers(-1000, 1000, 3 * 32).astype(np.int32); count = rng.integers(0, 1000, 3 * 32).astype(np.int32)
    e_stats, e_count = stats.copy(), count.copy()
    jobs = np.zeros(n, dtype=pkg.SAO_JOB)
    for i in range(n):
        jobs[i]["recOff"] = (1 + (i // 4) * 66) * S + (i % 4) * 70 + 2
        jobs[i]["diffOff"] = i * 4096
        jobs[i]["buf0"], jobs[i]["buf1"] = i * 160 + 2, i * 160 + 82
        jobs[i]["offsetOff"] = (i % 3) * 32                      # several CTUs accumulate into the same statistics slot
        jobs[i]["width"] = int(rng.integers(40, 64)); jobs[i]["height"] = int(rng.integers(1, 64))
        o = int(jobs[i]["offsetOff"])
        R.ref_sao_stats(kind, vpo(diff, i * 4096), vpo(rec, int(jobs[i]["recOff"])), ssz(S), vpo(e_buf, int(jobs[i]["buf0"])), vpo(e_buf, int(jobs[i]["buf1"])),
                        int(jobs[i]["width"]), int(jobs[i]["height"]), vpo(e_stats, o), vpo(e_count, o))
    bufs = [ctx.to_device(x) for x in (diff, rec, jobs, buf, stats, count)]
    ctx.sao_stats_dev(kind, depth, bufs[0], bufs[1], S, bufs[2], n, bufs[3], bufs[4], bufs[5])
    assert np.array_equal(bufs[4].download(np.int32), e_stats) and np.array_equal(bufs[5].download(np.int32), e_count), kind
    assert np.array_equal(bufs[3].download(np.int8), e_buf), kind
    for b in bufs:
        b.free()


@pytest.mark.parametrize("depth", [8, 10])
def test_deblock_batched(ctx, depth):
    R = oracle.ref(depth)
    rng = np.random.default_rng(300 + depth)
    S, rows, n = 256, 64, 60
    for chroma in (0, 1):
        pic = _plane(rng, depth, S, rows)
        e = pic.copy()
        jobs = np.zeros(n, dtype=pkg.DEBLOCK_JOB)
        for i in range(n):
            vertical = i % 2                                       # vertical edge: lines go down (srcStep = stride), taps across (offset = 1)
            x, y = 8 + (i % 15) * 16, 8 + (i // 15) * 12
            jobs[i]["srcOff"] = y * S + x
            jobs[i]["srcStep"], jobs[i]["offset"] = (S, 1) if vertical else (1, S)
            tc = int(rng.integers(0, 1 << (depth - 2)))
            if chroma:
                jobs[i]["tcP"], jobs[i]["tcQ"], jobs[i]["maskQ"] = tc, int(rng.integers(-1, 1)), int(rng.integers(-1, 1))
            else:
                jobs[i]["tcP"], jobs[i]["tcQ"] = tc & int(rng.integers(-1, 1)), tc
            R.ref_deblock(chroma, vpo(e, int(jobs[i]["srcOff"])), ssz(int(jobs[i]["srcStep"])), ssz(int(jobs[i]["offset"])),
                          int(jobs[i]["tcP"]), int(jobs[i]["tcQ"]), int(jobs[i]["maskQ"]))
        dP, dJ = ctx.to_device(pic), ctx.to_device(jobs)
        ctx.deblock_dev(chroma, depth, dP, dJ, n)
        assert np.array_equal(dP.download(pic.dtype), e), chroma
        dP.free(); dJ.free()
