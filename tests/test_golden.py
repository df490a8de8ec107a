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


# ---- families added late in round 1 (tests/golden/x265_golden_r1b.npz, made by tests/golden/make_golden_r1b.py) ---------------------
import ctypes  # noqa: E402

import oracle  # noqa: E402
from util import vp, vpo, ssz  # noqa: E402

G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "x265_golden_r1b.npz"))


def test_oracle_ads_and_propagate_match_golden():
    O = oracle.orc()
    sums, cost = G2["ads_sums"], G2["ads_cost"]
    for case, mvs in zip(G2["ads_cases"], G2["ads_mvs"]):
        kind, lxh, delta, thresh, off, n = [int(v) for v in case[:6]]
        enc = case[6:10].astype(np.int32)
        got = np.zeros(len(cost) + 8, dtype=np.int16)
        assert O.orc_ads(kind, lxh, vp(enc), vpo(sums, off), ssz(delta), vp(cost), vp(got), len(cost), thresh) == n
        assert np.array_equal(got[:n], mvs[:n])
    pin, intra, inter, invq, exp = G2["prop_in"], G2["prop_intra"], G2["prop_inter"], G2["prop_invq"], G2["prop_out"]   # keep the arrays alive
    dst = np.zeros(len(exp), dtype=np.int32)
    O.orc_propagate_cost(vp(dst), vp(pin), vp(intra), vp(inter), vp(invq), ctypes.c_double(192.0), len(dst))
    assert np.array_equal(dst, exp)


@pytest.mark.parametrize("depth", [8, 10])
def test_oracle_sea_sao_deblock_tu_match_golden(depth):
    O = oracle.orc()
    W, H, padX, padY = [int(v) for v in G2["sea_geom"]]
    S, rows = W + 2 * padX, H + 2 * padY
    plane = G2["sea_plane%d" % depth]
    origin = padY * S + padX
    pl = [np.zeros(S * rows, dtype=np.uint32) for _ in range(12)]
    arr = (ctypes.c_void_p * 12)(*[p.ctypes.data + origin * 4 for p in pl])
    O.orc_sea_integral(depth, vpo(plane, origin), ssz(S), padX, padY, H, arr)
    PW_, PH_ = [32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4], [32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4]
    for k in range(12):
        reg = pl[k].reshape(rows, S)[1:rows - PH_[k], :S - PW_[k]].astype(np.uint64)
        wts = (np.arange(reg.size, dtype=np.uint64).reshape(reg.shape) * np.uint64(2654435761) + np.uint64(1))
        assert [int(reg.sum(dtype=np.uint64)), int((reg * wts).sum(dtype=np.uint64))] == [int(v) for v in G2["sea_chk%d" % depth][k]], k
    img, buf, offs, diff = G2["sao_img%d" % depth], G2["sao_buf"], G2["sao_offs"], G2["sao_diff"]
    IS = 96
    for kind in range(6):
        r, b = img.copy(), buf.copy()
        O.orc_sao_apply(kind, depth, vpo(r, IS * 2 + 3), ssz(IS), vpo(b, 1), vpo(b, 81), vp(offs), 40, 6, 1)
        assert np.array_equal(r, G2["sao_apply%d_%d" % (kind, depth)]) and np.array_equal(b, G2["sao_applybuf%d_%d" % (kind, depth)]), kind
    for kind in (5, 0, 1, 3, 4):
        b = buf.copy(); st = np.arange(32, dtype=np.int32); ct = np.arange(32, dtype=np.int32) * 2
        O.orc_sao_stats(kind, depth, vp(diff), vpo(img, IS * 2 + 3), ssz(IS), vpo(b, 2), vpo(b, 82), 38, 30, vp(st), vp(ct))
        assert np.array_equal(np.stack([st, ct]), G2["sao_stats%d_%d" % (kind, depth)]) and np.array_equal(b, G2["sao_statsbuf%d_%d" % (kind, depth)]), kind
    for chroma in (0, 1):
        p = img.copy()
        for i in range(6):
            step, off = ((IS, 1) if i & 1 else (1, IS))
            O.orc_deblock(chroma, depth, vpo(p, IS * 16 + 12 + 8 * i), ssz(step), ssz(off), 3 + i, (5 if not chroma else -1), -1 * (i & 1))
        assert np.array_equal(p, G2["deblock%d_%d" % (chroma, depth)]), chroma
    for sizeIdx in range(4):
        N = 4 << sizeIdx
        fenc, pred = G2["tu_fenc%d_%d" % (sizeIdx, depth)], G2["tu_pred%d_%d" % (sizeIdx, depth)]
        qbits, add, scale, shift, ns, sse = [int(v) for v in G2["tu_meta%d_%d" % (sizeIdx, depth)]]
        qc = np.full(N * N, 16384, dtype=np.int32)
        rec, coef, s = np.zeros_like(fenc), np.zeros(N * N, dtype=np.int16), np.zeros(1, dtype=np.uint64)
        got = O.orc_tu_chain(depth, sizeIdx, 0, vp(fenc), ssz(N), vp(pred), ssz(N), vp(rec), ssz(N), vp(qc), qbits, add, None, scale, shift, vp(coef), vp(s))
        assert got == ns and int(s[0]) == sse and np.array_equal(rec, G2["tu_rec%d_%d" % (sizeIdx, depth)]) and np.array_equal(coef, G2["tu_coef%d_%d" % (sizeIdx, depth)]), sizeIdx
