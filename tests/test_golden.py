"""Golden fixtures (tests/golden/x265_golden.npz, generated from the compiled reference by
tests/golden/make_golden.py): the CPU restatement (no GPU needed) and the CUDA path (GPU) must both
reproduce the reference's recorded outputs.  Works on boxes that have neither /root/reference nor
oracle/_ref."""
import importlib
import os

import numpy as np
import pytest

from util import CU_SIZES, LUMA_PU_SIZES, STRIDE, orc_cmp

pkg = importlib.import_module("x265-yuuki-asuna_b200")
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "x265_golden.npz"))


@pytest.mark.parametrize("depth", [8, 10])
def test_oracle_restatement_matches_golden(depth):
    px, sh, offs, offb = G["px%d" % depth], G["sh%d" % depth], G["offs"], G["offb"]
    for kind in ("sad", "satd"):
        for i, (w, h) in enumerate(LUMA_PU_SIZES):
            assert orc_cmp(kind, depth, w, h, px, STRIDE, px, 61, offs, offb) == list(map(int, G["%s%d" % (kind, depth)][i])), (kind, w, h)
    for kind in ("sa8d", "sse_pp"):
        for i, s in enumerate(CU_SIZES):
            assert orc_cmp(kind, depth, s, s, px, STRIDE, px, 61, offs, offb) == list(map(int, G["%s%d" % (kind, depth)][i])), (kind, s)
    for kind in ("sse_ss", "ssd_s"):
        for i, s in enumerate(CU_SIZES):
            assert orc_cmp(kind, depth, s, s, sh, STRIDE, sh, 61, offs, offb) == list(map(int, G["%s%d" % (kind, depth)][i])), (kind, s)


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [8, 10])
def test_cuda_block_compares_match_golden(ctx, depth):
    px, sh, offs, offb = G["px%d" % depth], G["sh%d" % depth], G["offs"], G["offb"]
    K = {"sad": pkg.CMP_SAD, "satd": pkg.CMP_SATD, "sa8d": pkg.CMP_SA8D, "sse_pp": pkg.CMP_SSE_PP, "sse_ss": pkg.CMP_SSE_SS, "ssd_s": pkg.CMP_SSD_S}
    for kind in ("sad", "satd"):
        for i, (w, h) in enumerate(LUMA_PU_SIZES):
            got = ctx.pixelcmp_host(K[kind], depth, w, h, px, STRIDE, px, 61, offs, offb)
            assert list(map(int, got)) == list(map(int, G["%s%d" % (kind, depth)][i])), (kind, w, h)
    for kind in ("sa8d", "sse_pp", "sse_ss", "ssd_s"):
        src = sh if kind in ("sse_ss", "ssd_s") else px
        for i, s in enumerate(CU_SIZES):
            got = ctx.pixelcmp_host(K[kind], depth, s, s, src, STRIDE, src, 61, offs, offb)
            assert list(map(int, got)) == list(map(int, G["%s%d" % (kind, depth)][i])), (kind, s)


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [8, 10])
def test_cuda_transforms_interp_intra_match_golden(ctx, depth):
    sh, px = G["sh%d" % depth], G["px%d" % depth]
    for idx, N in ((0, 4), (1, 8), (2, 16), (3, 32), (4, 4)):
        dS, dC, dR = ctx.to_device(sh[:N * N * 3]), ctx.empty(N * N * 3 * 2), ctx.empty(N * N * 3 * 2)
        ctx.dct_dev(idx, depth, dS, N * N, N, dC, 3)
        ctx.idct_dev(idx, depth, dC, dR, N * N, N, 3)
        assert np.array_equal(dC.download(np.int16), G["dct%d_%d" % (idx, depth)]), ("dct", idx)
        assert np.array_equal(dR.download(np.int16), G["idct%d_%d" % (idx, depth)]), ("idct", idx)
    dP = ctx.to_device(px)
    job = np.zeros(9, dtype=pkg.INTERP_JOB)
    k = 0
    for cx in (1, 2, 3):
        for cy in (1, 2, 3):
            job[k] = (10 * STRIDE + 10, k * 256, cx, cy); k += 1
    dJ, dD = ctx.to_device(job), ctx.empty(9 * 256 * px.itemsize)
    ctx.interp_dev(pkg.IP_HVPP, 8, depth, 16, 16, dP, STRIDE, dD, 16, dJ, 9)
    assert np.array_equal(dD.download(px.dtype).reshape(9, 256), G["hvpp16_%d" % depth])
    nb = G["intra8_nb_%d" % depth]
    ij = np.zeros(35, dtype=pkg.INTRA_JOB)
    for m in range(35):
        ij[m] = (0, m * 64, m, 1)
    dN, dI, dO = ctx.to_device(nb), ctx.to_device(ij), ctx.empty(35 * 64 * nb.itemsize)
    ctx.intra_pred_dev(depth, 3, dN, dO, 8, dI, 35)
    assert np.array_equal(dO.download(nb.dtype).reshape(35, 64), G["intra8_%d" % depth])


@pytest.mark.gpu
def test_cuda_motion_estimate_matches_golden(ctx):
    W, H, S, origin, merange = [int(v) for v in G["me_geom"]]
    job = G["me_jobs"].copy().view(pkg.ME_JOB).reshape(-1)
    dC, dR = ctx.to_device(G["me_cur"]), ctx.to_device(G["me_ref"])
    lam = pkg.lambda_for_qp(30, 8)
    for name, m, sub in (("hex2", 1, 2), ("star3", 3, 3), ("dia0", 0, 0), ("umh2", 2, 2)):
        dJ = ctx.to_device(job)
        ctx.me_batch_dev(8, dC.ptr + origin, S, dR.ptr + origin, S, dJ, len(job), 64, 64, m, sub, merange, lam, 1)
        out = dJ.download(pkg.ME_JOB)
        got = np.stack([out["outMvX"], out["outMvY"], out["outCost"]], axis=1)
        assert np.array_equal(got, G["me_" + name]), name
        dJ.free()
