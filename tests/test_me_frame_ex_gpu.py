"""GPU parity of the general frame search (x265b200_me_frame_ex_dev, csrc/me_ctu_kernels.cu) against the reference's own
MotionEstimate (oracle/_ref) at the geometry of every BASELINE config:
  config 2  1080p 8-bit ultrafast : CTU 32 / minCU 16, DIA, subme 0, 1 reference                     (param.cpp:396-410)
  config 3  2160p 8-bit medium    : CTU 64, HEX, subme 2, 3 references, 2Nx2N
  config 4  2160p 10-bit slow     : STAR, subme 3 (chroma SATD), 4 references, rect PUs, per-PU mvp + candidates
  config 5  4320p 8-bit placebo   : STAR, merange 128, subme 5, 5 references, rect + AMP
Small frames are checked PU by PU; the full-size frames are searched whole (or in a band of CTU rows) on the GPU and a
sample of CTUs -- corners, edges, interior -- is checked against the reference."""
import importlib

import numpy as np
import pytest

from me_util import ctu_jobs, ctu_layout, ref_me, ref_me_chroma, synth_sequence

pkg = importlib.import_module("x265-yuuki-asuna_b200")
pytestmark = pytest.mark.gpu


class Frames:
    """a source picture + nref references (+ chroma for csp 1) resident on the device, PicYuv-style padding"""

    def __init__(self, ctx, W, H, padX, padY, depth, nref, csp, seed, picyuv_chroma=False):
        self.ctx, self.W, self.H, self.padX, self.padY, self.depth, self.csp = ctx, W, H, padX, padY, depth, csp
        # PicYuv keeps chromaMarginX = lumaMarginX (picyuv.cpp:104-106); the default test planes halve it
        self.cpadX, self.cpadY = (padX if picyuv_chroma else padX // 2), padY // 2
        self.item = 2 if depth > 8 else 1
        self.y, self.S, self.R, self.origin = synth_sequence(W, H, padX, padY, depth, nref + 1, seed)
        self.dY = [ctx.to_device(a) for a in self.y]
        if csp:
            assert csp == 1
            self.cb, self.Sc, self.Rc, self.originC = synth_sequence(W // 2, H // 2, self.cpadX, self.cpadY, depth, nref + 1, seed + 101, max_motion=6)
            self.cr, _, _, _ = synth_sequence(W // 2, H // 2, self.cpadX, self.cpadY, depth, nref + 1, seed + 202, max_motion=6)
            self.dCb = [ctx.to_device(a) for a in self.cb]
            self.dCr = [ctx.to_device(a) for a in self.cr]

    def free(self):
        for b in self.dY + (self.dCb + self.dCr if self.csp else []):
            b.free()


def run_frame(ctx, F, C, minCu, rect, amp, method, subme, merange, nref, rng, maxCand=0, per_pu_mvp=False, far_frac=0.0,
              ctuRow0=0, ctuRows=None, picW=None, picH=None, qp=30, mvp_span=24):
    """search every PU of the CTUs [ctuRow0, ctuRow0 + ctuRows) x all columns against nref references; returns the inputs and
    the output array [ref][ctu][pu][3]"""
    ctuCols = (F.W + C - 1) // C
    rowsAll = (F.H + C - 1) // C
    ctuRows = rowsAll if ctuRows is None else ctuRows
    layout = ctu_layout(C, minCu, rect, amp)
    n, nctu = len(layout), ctuCols * ctuRows
    assert np.array_equal(pkg.me_frame_layout(C, minCu, rect, amp), layout[:, :4])
    mvpCtu = rng.integers(-mvp_span, mvp_span + 1, (nref, nctu, 2)).astype(np.int32)
    mvpPu = ncand = mvc = None
    if per_pu_mvp:
        mvpPu = (mvpCtu[:, :, None, :] + rng.integers(-12, 13, (nref, nctu, n, 2))).astype(np.int32)
        if far_frac:
            sel = rng.random((nref, nctu, n)) < far_frac
            jump = (rng.choice([-1, 1], (nref, nctu, n, 2)) * rng.integers(4 * (merange + 20), 4 * (merange + 40), (nref, nctu, n, 2))).astype(np.int32)
            mvpPu[sel] += jump[sel]
    if maxCand:
        ncand = rng.integers(0, maxCand + 1, (nref, nctu, n)).astype(np.uint8)
        mvc = (mvpCtu[:, :, None, None, :] + rng.integers(-40, 41, (nref, nctu, n, maxCand, 2))).astype(np.int32)
    lam = pkg.lambda_for_qp(qp, F.depth)
    shY = ctuRow0 * C * F.S * F.item
    params = dict(depth=F.depth, ctuSize=C, minCuSize=minCu, rect=int(rect), amp=int(amp), picWidth=picW or F.W, picHeight=picH or F.H, firstCtuRow=ctuRow0,
                  ctuCols=ctuCols, ctuRows=ctuRows, marginX=F.padX, marginY=F.padY + ctuRow0 * C, rowsTotal=F.R, searchMethod=int(method),
                  subpelRefine=subme, merange=merange, csp=F.csp, maxCand=maxCand, maxSlices=1,
                  chromaMarginX=F.cpadX if F.csp else 0, chromaMarginY=(F.cpadY + ctuRow0 * (C // 2)) if F.csp else 0)
    params["lambda"] = lam
    dOut = ctx.empty(nref * nctu * n * 12)
    dev = [ctx.to_device(a) if a is not None else None for a in (mvpCtu, mvpPu, ncand, mvc)]
    kw = {}
    if F.csp:
        shC = ctuRow0 * (C // 2) * F.Sc * F.item
        oc = F.originC * F.item
        kw = dict(curC=(F.dCb[0].ptr + oc + shC, F.dCr[0].ptr + oc + shC), curStrideC=F.Sc, refCb=[b.ptr + oc + shC for b in F.dCb[1:nref + 1]],
                  refCr=[b.ptr + oc + shC for b in F.dCr[1:nref + 1]], refStrideC=F.Sc)
    o = F.origin * F.item
    ctx.me_frame_ex_dev(params, F.dY[0].ptr + o + shY, F.S, [b.ptr + o + shY for b in F.dY[1:nref + 1]], F.S, dOut,
                        dMvpCtu=dev[0], dMvpPu=dev[1], dNumCand=dev[2], dMvc=dev[3], **kw)
    out = dOut.download(np.int32).reshape(nref, nctu, n, 3)
    dOut.free()
    for d in dev:
        if d is not None:
            d.free()
    return dict(out=out, layout=layout, mvpCtu=mvpCtu, mvpPu=mvpPu, ncand=ncand, mvc=mvc, ctuCols=ctuCols, ctuRows=ctuRows, ctuRow0=ctuRow0,
                C=C, merange=merange, method=method, subme=subme, qp=qp, picW=picW or F.W, picH=picH or F.H, nref=nref)


def check_ctus(F, r, ctus):
    """compare the listed CTUs (column, row inside the searched band) of every reference with the reference's MotionEstimate"""
    C, lay, n = r["C"], r["layout"], len(r["layout"])
    total = 0
    for ref in range(r["nref"]):
        jobs, wants, idxs = [], [], []
        for (cx, cy) in ctus:
            ctu = cy * r["ctuCols"] + cx
            mvp = r["mvpPu"][ref, ctu] if r["mvpPu"] is not None else np.repeat(r["mvpCtu"][ref, ctu][None, :], n, axis=0)
            job, searched = ctu_jobs(pkg, lay, C, cx, cy + r["ctuRow0"], r["picW"], r["picH"], mvp, r["merange"],
                                     None if r["ncand"] is None else r["ncand"][ref, ctu], None if r["mvc"] is None else r["mvc"][ref, ctu])
            assert (r["out"][ref, ctu][~searched] == [0, 0, -1]).all(), (ref, cx, cy)
            jobs.append(job[searched]); wants.append(r["out"][ref, ctu][searched]); idxs += [(cx, cy, int(i)) for i in np.nonzero(searched)[0]]
        job = np.concatenate(jobs); got = np.concatenate(wants)
        if F.csp:
            ex, ey, ec = ref_me_chroma(F.depth, F.csp, F.y[0], F.y[1 + ref], F.S, F.origin, (F.cb[0], F.cr[0]), (F.cb[1 + ref], F.cr[1 + ref]), F.Sc, F.originC,
                                       job, r["method"], r["subme"], r["merange"], r["qp"], threads=8)
        else:
            ex, ey, ec = ref_me(F.depth, F.y[0], F.y[1 + ref], F.S, F.origin, job, r["method"], r["subme"], r["merange"], r["qp"], threads=8)
        bad = np.nonzero((got[:, 0] != ex) | (got[:, 1] != ey) | (got[:, 2] != ec))[0]
        assert not len(bad), (ref, idxs[bad[0]], job[bad[0]], got[bad[0]].tolist(), int(ex[bad[0]]), int(ey[bad[0]]), int(ec[bad[0]]), len(bad), len(job))
        total += len(job)
    return total


SMALL = [
    # depth C minCu rect amp method subme merange csp maxCand perPu far
    (8, 64, 8, False, False, pkg.ME_HEX, 2, 57, 0, 0, False, 0.0),
    (8, 32, 16, False, False, pkg.ME_DIA, 0, 57, 0, 0, False, 0.0),
    (8, 64, 8, True, False, pkg.ME_STAR, 3, 24, 1, 3, True, 0.0),
    (8, 64, 8, True, True, pkg.ME_UMH, 5, 16, 1, 3, True, 0.0),
    (8, 64, 16, True, True, pkg.ME_HEX, 2, 16, 0, 2, True, 0.15),
    (8, 16, 8, True, True, pkg.ME_HEX, 7, 16, 1, 2, True, 0.0),
    (10, 64, 8, True, False, pkg.ME_STAR, 3, 24, 1, 3, True, 0.1),
    (10, 32, 8, False, False, pkg.ME_UMH, 2, 40, 0, 2, True, 0.0),
    (8, 64, 8, False, False, pkg.ME_FULL, 1, 6, 0, 0, True, 0.0),
    (8, 64, 8, True, True, pkg.ME_STAR, 5, 128, 1, 2, True, 0.0),       # config-5 shape: two TMA boxes per window, 206 KB of shared memory
]


@pytest.mark.parametrize("depth,C,minCu,rect,amp,method,subme,merange,csp,maxCand,perPu,far", SMALL)
def test_small_frame_every_pu(ctx, depth, C, minCu, rect, amp, method, subme, merange, csp, maxCand, perPu, far):
    cols, rows, nref = (3 if C > 16 else 4), 2, 2          # plane strides stay multiples of 16 bytes (TMA), as every x265 plane is
    pad = 64 + merange + 64 + (80 if far else 0)
    pad = (pad + 31) & ~31
    F = Frames(ctx, cols * C, rows * C, pad, pad, depth, nref, csp, seed=7000 + depth + C + subme)
    r = run_frame(ctx, F, C, minCu, rect, amp, method, subme, merange, nref, np.random.default_rng(11 + subme), maxCand=maxCand, per_pu_mvp=perPu, far_frac=far)
    check_ctus(F, r, [(x, y) for y in range(rows) for x in range(cols)])
    F.free()


def test_picyuv_chroma_margins(ctx):
    """the chroma planes of a real PicYuv keep chromaMarginX = lumaMarginX (picyuv.cpp:104-106): explicit chroma margins in the params"""
    C, cols, rows, nref, merange = 64, 3, 2, 2, 24
    F = Frames(ctx, cols * C, rows * C, 160, 160, 8, nref, 1, seed=7150, picyuv_chroma=True)
    r = run_frame(ctx, F, C, 8, True, False, pkg.ME_HEX, 3, merange, nref, np.random.default_rng(7), maxCand=2, per_pu_mvp=True)
    check_ctus(F, r, [(x, y) for y in range(rows) for x in range(cols)])
    F.free()


def test_picture_smaller_than_ctu_grid(ctx):
    """CUs that leave the picture are not searched (cost -1); clipMv against the real picture size."""
    C, cols, rows = 64, 3, 2
    F = Frames(ctx, cols * C, rows * C, 160, 160, 8, 1, 0, seed=7100)
    r = run_frame(ctx, F, C, 8, True, False, pkg.ME_HEX, 2, 32, 1, np.random.default_rng(5), per_pu_mvp=True, picW=cols * C - 40, picH=rows * C - 24)
    check_ctus(F, r, [(x, y) for y in range(rows) for x in range(cols)])
    assert (r["out"][0, cols - 1, 0] == [0, 0, -1]).all()          # the 64x64 PU of the last column leaves the picture
    F.free()


def test_equals_the_2Nx2N_frame_entry(ctx):
    """the 2Nx2N-only entry (x265b200_me_frame_dev) and the general one agree PU for PU on the config-3 shape"""
    C, cols, rows, nref, merange = 64, 4, 3, 3, 57
    F = Frames(ctx, cols * C, rows * C, 160, 160, 8, nref, 0, seed=7200)
    r = run_frame(ctx, F, C, 8, False, False, pkg.ME_HEX, 2, merange, nref, np.random.default_rng(9))
    dMvp = ctx.to_device(r["mvpCtu"])
    per_level = [cols * rows * (1 << l) ** 2 for l in range(4)]
    dOut = ctx.empty(nref * sum(per_level) * 12)
    o = F.origin
    ctx.me_frame_dev(8, F.dY[0].ptr + o, F.S, [b.ptr + o for b in F.dY[1:]], F.S, F.padX, F.padY, F.R, cols, rows, 15, dMvp, pkg.ME_HEX, 2, merange,
                     pkg.lambda_for_qp(30, 8), dOut)
    old = dOut.download(np.int32).reshape(nref, sum(per_level), 3)
    lay = r["layout"]
    off = np.cumsum([0] + per_level)
    for ref in range(nref):
        for ctu in range(cols * rows):
            cx, cy = ctu % cols, ctu // cols
            for i, (x, y, w, h, _, _, S) in enumerate(lay.tolist()):
                level = {64: 0, 32: 1, 16: 2, 8: 3}[S]
                per = 1 << level
                gx, gy = cx * per + x // S, cy * per + y // S
                assert (old[ref, off[level] + gy * cols * per + gx] == r["out"][ref, ctu, i]).all(), (ref, ctu, i)
    dMvp.free(); dOut.free(); F.free()


def test_host_buffer_entries_equal_the_device_entry(ctx):
    """x265b200_me_frame_ex_host and its _begin / _end halves (source picture from host memory over the copy stream, results to
    host memory) return what x265b200_me_frame_ex_dev returns; work queued between _begin and _end does not disturb them; so does
    x265b200_me_frame_host against x265b200_me_frame_dev."""
    import ctypes
    C, cols, rows, nref, merange = 64, 3, 2, 2, 24
    F = Frames(ctx, cols * C, rows * C, 160, 160, 8, nref, 1, seed=7300)
    r = run_frame(ctx, F, C, 8, True, False, pkg.ME_HEX, 3, merange, nref, np.random.default_rng(3))
    want = r["out"]
    n, nctu = len(r["layout"]), cols * rows
    params = dict(depth=8, ctuSize=C, minCuSize=8, rect=1, amp=0, picWidth=F.W, picHeight=F.H, firstCtuRow=0, ctuCols=cols, ctuRows=rows, marginX=F.padX,
                  marginY=F.padY, rowsTotal=F.R, searchMethod=int(pkg.ME_HEX), subpelRefine=3, merange=merange, csp=1, maxCand=0, maxSlices=1,
                  chromaMarginX=F.cpadX, chromaMarginY=F.cpadY)
    params["lambda"] = pkg.lambda_for_qp(30, 8)
    dMvp = ctx.to_device(r["mvpCtu"])
    o, oc = F.origin, F.originC
    # fresh device planes for the source picture: the host entries must fill them
    dY, dCb, dCr = ctx.to_device(np.zeros_like(F.y[0])), ctx.to_device(np.zeros_like(F.cb[0])), ctx.to_device(np.zeros_like(F.cr[0]))
    kw = dict(curC=(dCb.ptr + oc, dCr.ptr + oc), curStrideC=F.Sc, refCb=[b.ptr + oc for b in F.dCb[1:]], refCr=[b.ptr + oc for b in F.dCr[1:]], refStrideC=F.Sc,
              hostC=(F.cb[0].ctypes.data, F.cr[0].ctypes.data), devCBase=(dCb.ptr, dCr.ptr), bytesC=F.cb[0].nbytes)
    dOut = ctx.empty(nref * nctu * n * 12)
    for mode in ("sync", "begin_end"):
        host = np.full((nref, nctu, n, 3), -7, dtype=np.int32)
        args = (params, dY.ptr + o, F.S, [b.ptr + o for b in F.dY[1:]], F.S, F.y[0].ctypes.data, dY.ptr, F.y[0].nbytes, dOut, host.ctypes.data, host.nbytes)
        if mode == "sync":
            ctx.me_frame_ex_host(*args, dMvpCtu=dMvp, **kw)
        else:
            ctx.me_frame_ex_host_begin(*args, dMvpCtu=dMvp, **kw)
            scratch = ctx.to_device(np.arange(1 << 16, dtype=np.int32))        # unrelated work on the compute stream in between
            ctx.me_frame_host_end()
            scratch.free()
        assert np.array_equal(host, want), mode
        assert np.array_equal(dY.download(F.y[0].dtype), F.y[0])
        dY.upload(np.zeros_like(F.y[0]))
    # two calls pending: the second is queued before the first is read
    dY.upload(np.zeros_like(F.y[0]))
    hostA, hostB = np.full((nref, nctu, n, 3), -7, dtype=np.int32), np.full((nref, nctu, n, 3), -7, dtype=np.int32)
    dOutB = ctx.empty(nref * nctu * n * 12)
    base = (params, dY.ptr + o, F.S, [b.ptr + o for b in F.dY[1:]], F.S, F.y[0].ctypes.data, dY.ptr, F.y[0].nbytes)
    ctx.me_frame_ex_host_begin(*base, dOut, hostA.ctypes.data, hostA.nbytes, dMvpCtu=dMvp, **kw)
    ctx.me_frame_ex_host_begin(*base, dOutB, hostB.ctypes.data, hostB.nbytes, dMvpCtu=dMvp, **kw)
    with pytest.raises(pkg.X265B200Error):
        ctx.me_frame_ex_host_begin(*base, dOutB, hostB.ctypes.data, hostB.nbytes, dMvpCtu=dMvp, **kw)        # a third one is refused
    ctx.me_frame_host_end()
    assert np.array_equal(hostA, want)
    ctx.me_frame_host_end()
    assert np.array_equal(hostB, want)
    dOutB.free()
    # the 2Nx2N entry
    per_level = [nctu * (1 << l) ** 2 for l in range(4)]
    dOld = ctx.empty(nref * sum(per_level) * 12)
    ctx.me_frame_dev(8, F.dY[0].ptr + o, F.S, [b.ptr + o for b in F.dY[1:]], F.S, F.padX, F.padY, F.R, cols, rows, 15, dMvp, pkg.ME_HEX, 2, merange,
                     params["lambda"], dOld)
    old = dOld.download(np.int32)
    host = np.zeros_like(old)
    ctx.me_frame_host(8, F.y[0].ctypes.data, F.y[0].nbytes, dY.ptr, F.S, [b.ptr + o for b in F.dY[1:]], F.S, F.padX, F.padY, F.R, cols, rows, 15, dMvp,
                      pkg.ME_HEX, 2, merange, params["lambda"], dOld, host.ctypes.data, host.nbytes)
    assert np.array_equal(host, old)
    for b in (dY, dCb, dCr, dOut, dOld, dMvp):
        b.free()
    F.free()


def _sample_ctus(cols, rows, k, rng):
    pts = {(0, 0), (cols - 1, 0), (0, rows - 1), (cols - 1, rows - 1), (cols // 2, rows // 2)}
    while len(pts) < k:
        pts.add((int(rng.integers(0, cols)), int(rng.integers(0, rows))))
    return sorted(pts)


def test_config2_1080p_ctu32_dia(ctx):
    """BASELINE config 2 at its own geometry: 1920x1080, CTU 32 / minCU 16, DIA, subme 0, merange 57, 1 reference."""
    F = Frames(ctx, 1920, 1080, 64, 64 + 8, 8, 1, 0, seed=7300)          # picture rows 1080: the last CTU row is cut
    r = run_frame(ctx, F, 32, 16, False, False, pkg.ME_DIA, 0, 57, 1, np.random.default_rng(1), mvp_span=0)
    assert r["ctuCols"] == 60 and r["ctuRows"] == 34
    n = check_ctus(F, r, _sample_ctus(60, 34, 40, np.random.default_rng(2)) + [(x, 16) for x in range(60)])
    assert n > 400
    F.free()


def test_config4_2160p_10bit_star_chroma_rect(ctx):
    """BASELINE config 4 at its own geometry: 3840x2160 10-bit 4:2:0, STAR, subme 3 (chroma SATD), 4 references, rect PUs,
    per-PU predictors and candidates; the whole frame is searched, sampled CTUs (incl. the cut last row) are checked."""
    F = Frames(ctx, 3840, 2160, 96, 80, 10, 4, 1, seed=7400)
    r = run_frame(ctx, F, 64, 8, True, False, pkg.ME_STAR, 3, 57, 4, np.random.default_rng(3), maxCand=3, per_pu_mvp=True)
    assert r["ctuCols"] == 60 and r["ctuRows"] == 34 and len(r["layout"]) == 425
    n = check_ctus(F, r, _sample_ctus(60, 34, 8, np.random.default_rng(4)))
    assert n > 4 * 5 * 300
    F.free()


def test_config5_4320p_star_merange128_amp(ctx):
    """BASELINE config 5 at its own geometry: 7680x4320 8-bit 4:2:0, STAR, merange 128, subme 5, 5 references, rect + AMP.
    A band of 2 CTU rows in the middle of the frame is searched on the full-size planes; sampled CTUs are checked."""
    F = Frames(ctx, 7680, 4320, 96, 80, 8, 5, 1, seed=7500)
    r = run_frame(ctx, F, 64, 8, True, True, pkg.ME_STAR, 5, 128, 5, np.random.default_rng(5), maxCand=2, per_pu_mvp=True, ctuRow0=33, ctuRows=2)
    assert len(r["layout"]) == 593
    n = check_ctus(F, r, [(0, 0), (119, 1), (57, 0), (88, 1)])
    assert n > 5 * 4 * 500
    F.free()
