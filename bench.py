#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 x265 hot-path backend (contract in the task brief).

`--config N` selects one of BASELINE.json's configurations (default 3, the one the metric is quoted on):
  2  1080p 8-bit ultrafast : CTU 32 / minCU 16, DIA, subme 0, merange 57, 1 reference   -- "ME primitives": streaming SAD + frame search
  3  2160p 8-bit medium    : the full primitive mix of SURVEY.md 8d config 3 (below)    -- DEFAULT
  4  2160p 10-bit slow     : STAR, subme 3 (chroma SATD), 4 references, rect PUs        -- streaming SAD + frame search
  5  4320p 8-bit placebo   : STAR, merange 128, subme 5, 5 references, rect + AMP       -- streaming SAD + frame search

A "step" is one frame through the hot path.  Config 3 (HEX, subme 2, merange 57, 3 references, 2Nx2N PUs 64..8, mvp = 0):
  1. sad_pred  : SAD at the predictor of every PU level x reference (the streaming TMA kernel of sad_stream_kernels.cu)
  2. me_search : MotionEstimate::motionEstimate for every PU x reference (me_frame kernel, TMA-staged windows)
  3. mc        : one 8-tap luma interpolation per PU and level (all 15 fractions)
  4. residual  : fused residual pipeline (fenc - pred -> DCT -> quant -> dequant -> IDCT -> recon -> SSE) on every 32/16/8/4 TU
  5. intra     : neighbour filter + all 35 modes on every 8/16/32 block
  6. lookahead : Lowres::init + lowresIntraEstimate of the new frame and, per batch of 8 frames, the estimateFrameCost list
                 searches medium triggers (bframes 4: 5 frame-triples = 10 list searches per frame, --lookahead-slices 8),
                 on a second stream, overlapped with 1-5 and INSIDE the timed region.
`value` = frames/s with frames resident in HBM (ring of 32 frames = 318 MB > L2, no flush needed);
`e2e`   = the same with every new frame arriving from pinned HOST memory through the host-buffer C-ABI entry
          (x265b200_me_frame_host / x265b200_me_frame_ex_host: H2D of the frame, search, D2H of the {mv,cost} records).
`--impl reference` times the reference's own CPU code (oracle/_ref, compiled from the unmodified x265 sources; C table +
the SSE-intrinsic DCT table, no nasm here) for the same step on all host cores, on a bounded sample of CTU rows.
`--shard frames` (default; weak scaling: every rank its own frames + all_gather of the new reference plane + gather of the
results, on a comm stream overlapped with the next step) or `--shard ctu-rows` (strong scaling of ONE frame: a band of CTU
rows per rank).
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# ---- config 3 geometry (module-level names: tests/ and scripts/ use them) ------------------------------------------------------
W, H, CTU, PAD = 3840, 2160, 64, 128          # luma; PAD >= merange + 8 + halo, multiple of 64
STRIDE = W + 2 * PAD
ROWS = H + 2 * PAD + (64 - (H % 64)) % 64      # 34 CTU rows (2176) + margins
CTU_COLS, CTU_ROWS = W // CTU, (H + CTU - 1) // CTU
NREF, MERANGE, SUBME, QP = 3, 57, 2, 30
LEVELS = [64, 32, 16, 8]
METRIC = "2160p preset-medium fps at 1/2/4/8 B200; ME SAD achieved HBM GB/s vs peak"
LA_SMS = 0                                    # SMs set aside for the lookahead stream (0: none), see --la-sms
LA_BATCH, LA_BFRAMES, LA_SLICES = int(os.environ.get("BENCH_LA_BATCH", "16")), 4, 8      # frames per estimateFrameCost launch; bframes; --lookahead-slices (param.cpp:173)

ME_DIA, ME_HEX, ME_UMH, ME_STAR = 0, 1, 2, 3
CONFIGS = {
    2: dict(name="1080p 8-bit ultrafast: CTU 32 / minCU 16, DIA, subme 0, merange 57, 1 reference (param.cpp:396-410)", W=1920, H=1080, depth=8, C=32, minCu=16,
            rect=0, amp=0, method=ME_DIA, subme=0, merange=57, nref=1, csp=0, NF=64),
    3: dict(name="2160p 8-bit medium primitive mix", W=W, H=H, depth=8, C=64, minCu=8, rect=0, amp=0, method=ME_HEX, subme=SUBME, merange=MERANGE, nref=NREF, csp=0, NF=32),
    4: dict(name="2160p 10-bit slow: STAR, subme 3 (chroma SATD), 4 references, rect PUs, 4:2:0", W=3840, H=2160, depth=10, C=64, minCu=8, rect=1, amp=0,
            method=ME_STAR, subme=3, merange=57, nref=4, csp=1, NF=20),
    5: dict(name="4320p 8-bit placebo: STAR, merange 128, subme 5, 5 references, rect + AMP, 4:2:0", W=7680, H=4320, depth=8, C=64, minCu=8, rect=1, amp=1,
            method=ME_STAR, subme=5, merange=128, nref=5, csp=1, NF=12),
}
WORKLOAD3 = ("2160p-8bit-medium primitive mix (SURVEY.md 8d config 3): SAD at the predictor + HEX subme2 merange57 search of every 2Nx2N PU 64..8 x 3 refs; "
             "one 8-tap MC interpolation per PU and level (all 15 fractions); residual -> DCT/quant/dequant/IDCT on every 32/16/8/4 TU -> recon; "
             "intra neighbour smoothing + all 35 modes on every 8/16/32 block (fused); lookahead (lowres init + intra estimate + 10 estimateFrameCost list "
             "searches per frame, --lookahead-slices 8; the searches of %d queued frames go in one launch, within medium's rc-lookahead of 20) on a second stream, "
             "inside the timed region" % LA_BATCH)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def box_blur(a, k=5):
    for ax in (0, 1):
        c = np.cumsum(a, axis=ax, dtype=np.float64)
        pad = [(0, 0), (0, 0)]; pad[ax] = (k, 0)
        c = np.pad(c, pad)
        n = a.shape[ax]
        a = (np.take(c, np.arange(k, n + k), axis=ax) - np.take(c, np.arange(0, n), axis=ax)) / k
    return a


def synth_planes(nframes, rows, stride, depth=8, seed=1234, step=6, noise=2.0):
    """band-limited noise + per-frame global translation + noise, padded planes of rows x stride pixels"""
    rng = np.random.default_rng(seed)
    big = box_blur(rng.integers(0, 256, (rows + 96, stride + 96)).astype(np.float64))
    big = ((big - big.min()) / (big.max() - big.min()) * 255.0).astype(np.float32)
    mrng = np.random.default_rng(5678)
    scale, pmax = 1 << (depth - 8), (1 << depth) - 1
    dt = np.uint16 if depth > 8 else np.uint8
    frames = []
    x, y = 48, 48
    for f in range(nframes):
        dx, dy = mrng.integers(-step, step + 1, 2)
        x = int(np.clip(x + dx, 0, 95)); y = int(np.clip(y + dy, 0, 95))
        fr = big[y:y + rows, x:x + stride] + rng.normal(0, noise, (rows, stride)).astype(np.float32)
        frames.append(np.clip(np.rint(fr * scale), 0, pmax).astype(dt))
    return frames


def synth_frames(nframes, seed=1234):
    """the config-3 frames (ROWS x STRIDE, 8-bit)"""
    return synth_planes(nframes, ROWS, STRIDE, 8, seed)


def build_jobs(pkg, ctu_rows=None):
    """one ME job per (ref, level, PU) of config 3: mvp = 0, window = +-merange."""
    rows = range(CTU_ROWS) if ctu_rows is None else ctu_rows
    js = []
    for r in range(NREF):
        for s in LEVELS:
            per = CTU // s
            for cy in rows:
                for cx in range(CTU_COLS):
                    for py in range(per):
                        y = cy * CTU + py * s
                        for px in range(per):
                            js.append((cx * CTU + px * s, y, s, r))
    a = np.array(js, dtype=np.int32)
    job = np.zeros(len(a), dtype=pkg.ME_JOB)
    job["puX"], job["puY"], job["w"], job["h"], job["refIdx"] = a[:, 0], a[:, 1], a[:, 2], a[:, 2], a[:, 3]
    job["mvminX"] = job["mvminY"] = -MERANGE
    job["mvmaxX"] = job["mvmaxY"] = MERANGE
    return job


TU_SIZES = [(3, 32), (2, 16), (1, 8), (0, 4)]      # (sizeIdx, N): every TU size of the residual plane
INTRA_SIZES = [(1, 8, 3), (2, 16, 4), (3, 32, 5)]   # (sizeIdx, N, log2N): all 35 modes on every block
IP_HPP, IP_VPP, IP_HVPP = 0, 2, 6                   # X265B200_IP_* / ref_interp_batch kinds


def interp_jobs(pkg, s, rows=None):
    """one luma interpolation per PU of level s: block at its own position in reference 0, fraction cycling through
    the 15 non-zero quarter-pel positions.  Returns {kind: job array}; srcOff is relative to
    the plane origin (STRIDE pitch), dstOff addresses a W-pitch prediction plane."""
    per_row = W // s
    out = {IP_HPP: [], IP_VPP: [], IP_HVPP: []}
    ys = range(0, CTU_ROWS * CTU, s) if rows is None else [cy * CTU + py * s for cy in rows for py in range(CTU // s)]
    for y in ys:
        for bx in range(per_row):
            k = ((y // s) * per_row + bx) % 15 + 1
            fx, fy = k & 3, k >> 2
            x = bx * s
            src, dst = y * STRIDE + x, y * W + x
            if fy == 0:
                out[IP_HPP].append((src, dst, fx, 0))
            elif fx == 0:
                out[IP_VPP].append((src, dst, fy, 0))
            else:
                out[IP_HVPP].append((src, dst, fx, fy))
    res = {}
    for kind, lst in out.items():
        a = np.zeros(len(lst), dtype=pkg.INTERP_JOB)
        if lst:
            t = np.array(lst, dtype=np.int64)
            a["srcOff"], a["dstOff"], a["idxX"], a["idxY"] = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
        res[kind] = a
    return res


def neighbour_arrays(frame, N, rows=None):
    """[topLeft, top 2N, left 2N] (intrapred.cpp:36-50 layout) of every N x N block, taken from the padded frame."""
    f = frame.reshape(ROWS, STRIDE)
    ys = np.arange(0, CTU_ROWS * CTU, N) if rows is None else np.array([cy * CTU + k * N for cy in rows for k in range(CTU // N)])
    xs = np.arange(0, W, N)
    Y, X = np.meshgrid(ys + PAD, xs + PAD, indexing="ij")
    Y, X = Y.ravel(), X.ravel()
    out = np.empty((len(Y), 4 * N + 1), dtype=np.uint8)
    out[:, 0] = f[Y - 1, X - 1]
    k = np.arange(2 * N)
    out[:, 1:2 * N + 1] = f[(Y - 1)[:, None], X[:, None] + k[None, :]]
    out[:, 2 * N + 1:] = f[Y[:, None] + k[None, :], (X - 1)[:, None]]
    return out


def quant_params(N):
    log2 = int(np.log2(N))
    qbits = 14 + 5 + (15 - 8 - log2)                # QUANT_SHIFT + per(qp 30) + transformShift (quant.cpp:397-480)
    return qbits, 171 << (qbits - 9)


class ClockSampler(threading.Thread):
    def __init__(self, idx):
        super().__init__(daemon=True)
        self.idx, self.samples, self.stop_flag = idx, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append([x.strip() for x in o])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons)}


def geometry(cfg):
    """PicYuv-style padded geometry of a config (picyuv.cpp:87-89: marginX = maxCU + 32, marginY = maxCU + 16); config 3 keeps the
    128-pixel pad of the round-1 bench (tests share its constants)."""
    C = cfg["C"]
    cols, rows = (cfg["W"] + C - 1) // C, (cfg["H"] + C - 1) // C
    if cfg is CONFIGS[3]:
        return dict(cols=cols, rows=rows, padX=PAD, padY=PAD, stride=STRIDE, planeRows=ROWS)
    padX, padY = C + 32, C + 16
    if cfg["merange"] > 64:                      # the search window of merange 128 reaches further than the encoder's own padding
        padX, padY = 64 + cfg["merange"] + 16, 64 + cfg["merange"] + 16
    padX = (padX + 31) & ~31; padY = (padY + 1) & ~1
    return dict(cols=cols, rows=rows, padX=padX, padY=padY, stride=cols * C + 2 * padX, planeRows=rows * C + 2 * padY)


# =====================================================================================================================
# the reference arm: the reference's own CPU code on the host cores
# =====================================================================================================================
def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle
    pkg = importlib.import_module("x265-yuuki-asuna_b200")
    cfg = CONFIGS[args.config]
    R = oracle.ref(cfg["depth"])
    cores = os.cpu_count() or 1
    if R is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libx265ref%d.so not present" % cfg["depth"]}))
        return
    intr = R.ref_enable_intrinsics()              # the SSE-intrinsic DCT table (vec/): the optimised table that builds without nasm
    if args.config == 3:
        step, frac, sample, la = reference_mix(R, pkg, cores)
    else:
        step, frac, sample, la = reference_me(R, pkg, cfg, cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    per_frame = dt / frac + la["s_per_frame_all_cores"]
    fps = 1.0 / per_frame
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_frame * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8" if cfg["depth"] == 8 else "u16",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD3 if args.config == 3 else cfg["name"], "baseline_config": args.config, "sample": sample},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                             "sample": sample + "; C table + %d SSE-intrinsic slots (common/vec), no nasm asm in this image; ME / MC / TU / intra jobs handed to %d "
                                       "threads from an atomic queue; lookahead leg: %s" % (intr, cores, la["how"])},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def reference_lookahead(R, frames, stride, origin, width, height, cores, depth=8):
    """one frame's worth of lookahead on the reference's own Lookahead objects: Lowres::init + lowresIntraEstimate + the 5
    frame-triples of `bframes 4` (singleCost, the non-pool path: one thread); reported per frame as if the encoder's pool
    scaled it perfectly over the cores (the most favourable reading for the CPU)."""
    R.ref_la_create.restype = ctypes.c_void_p
    R.ref_la_frame_cost.restype = ctypes.c_int64
    n = 2 * (LA_BFRAMES + 1) + 1
    h = ctypes.c_void_p(R.ref_la_create(width, height, LA_BFRAMES, 0))
    item = frames[0].itemsize
    t0 = time.perf_counter()
    for f in frames[:n]:
        R.ref_la_add_frame(h, ctypes.c_void_p(f.ctypes.data + origin * item), ctypes.c_ssize_t(stride))
    t_init = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    R.ref_la_intra(h, LA_BFRAMES + 1)
    t_intra = time.perf_counter() - t0
    b = LA_BFRAMES + 1
    t0 = time.perf_counter()
    for dd in range(1, LA_BFRAMES + 2):
        R.ref_la_frame_cost(h, b - dd, b + dd, b, 0)
    t_est = time.perf_counter() - t0
    R.ref_la_destroy(h)
    one = t_init + t_intra + t_est
    return {"s_per_frame_1thread": one, "s_per_frame_all_cores": one / cores,
            "how": "Lowres::init %.1f ms + lowresIntraEstimate %.1f ms + 5 frame-triples (10 list searches) %.1f ms on ONE thread, divided by %d cores" % (
                t_init * 1e3, t_intra * 1e3, t_est * 1e3, cores)}


def reference_mix(R, pkg, cores):
    from me_util import REF_ME_JOB
    frames = synth_frames(max(NREF + 1, 2 * (LA_BFRAMES + 1) + 1))
    sample_rows = [0, CTU_ROWS // 3, (2 * CTU_ROWS) // 3, CTU_ROWS - 1]          # top, two interior rows, bottom
    job = build_jobs(pkg, sample_rows)
    frac = len(sample_rows) / CTU_ROWS
    origin = PAD * STRIDE + PAD
    cur = frames[NREF].ravel()
    # everything a step needs is prepared here, outside the timed region
    rjs = []
    rng = np.random.default_rng(3)
    for r in range(NREF):
        jr = job[job["refIdx"] == r]
        rj = np.zeros(len(jr), dtype=REF_ME_JOB)
        for f in ("puX", "puY", "w", "h", "mvminX", "mvminY", "mvmaxX", "mvmaxY", "mvpX", "mvpY", "numCand", "mvc"):
            rj[f] = jr[f]
        rjs.append(rj[rng.permutation(len(rj))])      # sizes interleaved as well (the queue is dynamic anyway)
    ip_jobs = {sz: interp_jobs(pkg, sz, sample_rows) for sz in LEVELS}
    pred = {sz: np.zeros(CTU_ROWS * CTU * W, dtype=np.uint8) for sz in LEVELS}
    qtab = np.full(1024, 26214, dtype=np.int32)
    recon = np.zeros(CTU * W, dtype=np.uint8); tu_coef = np.zeros(CTU * W, dtype=np.int16)
    tu_ns = np.zeros((W // 4) * (CTU // 4), dtype=np.uint32); tu_sse = np.zeros((W // 4) * (CTU // 4), dtype=np.uint64)
    nbr = {N: neighbour_arrays(cur, N, sample_rows) for _, N, _ in INTRA_SIZES}
    filt = {N: np.empty_like(nbr[N]) for N in nbr}
    intra_out = {N: np.empty(len(nbr[N]) * 35 * N * N, dtype=np.uint8) for N in nbr}
    parts = {sz: R.ref_partition_from_sizes(sz, sz) for sz in LEVELS}
    refs = [frames[NREF - 1 - r].ravel() for r in range(NREF)]
    vp = lambda a, off=0: ctypes.c_void_p(a.ctypes.data + off)

    def step():
        for r in range(NREF):
            R.ref_me_batch(vp(cur, origin), ctypes.c_ssize_t(STRIDE), vp(refs[r], origin), ctypes.c_ssize_t(STRIDE),
                           vp(rjs[r]), ctypes.c_int64(len(rjs[r])), 1, SUBME, MERANGE, QP, 1, cores)
        for sz in LEVELS:
            for kind, ja in ip_jobs[sz].items():
                if len(ja):
                    R.ref_interp_batch(kind, parts[sz], vp(refs[0], origin), ctypes.c_ssize_t(STRIDE), vp(pred[sz]), ctypes.c_ssize_t(W), vp(ja), ctypes.c_int64(len(ja)), cores)
        for cy in sample_rows:
            y0 = cy * CTU
            for idx, N in TU_SIZES:
                qbits, add = quant_params(N)
                R.ref_tu_pipeline(idx, 0, vp(cur, origin + y0 * STRIDE), ctypes.c_ssize_t(STRIDE), vp(pred[16], y0 * W), ctypes.c_ssize_t(W), vp(recon), ctypes.c_ssize_t(W),
                                  W // N, CTU // N, vp(qtab), qbits, add, None, 40 << 5, 9, vp(tu_coef), vp(tu_ns), vp(tu_sse), cores)
        for idx, N, _ in INTRA_SIZES:
            R.ref_intra_batch(idx, vp(nbr[N]), vp(filt[N]), vp(intra_out[N]), ctypes.c_int64(len(nbr[N])), cores)

    la = reference_lookahead(R, frames, STRIDE, origin, W, CTU_ROWS * CTU, cores)
    sample = "CTU rows %s of %d (x%.1f to a frame): %d PU searches (3 refs), %d MC interpolations, DCT/quant/dequant/IDCT of their 32/16/8/4 TUs, 35 intra modes on their 8/16/32 blocks" % (
        sample_rows, CTU_ROWS, 1 / frac, len(job), sum(len(a) for d in ip_jobs.values() for a in d.values()))
    return step, frac, sample, la


def reference_me(R, pkg, cfg, cores):
    """configs 2 / 4 / 5: the frame search of the config's partition set on sampled CTU rows, through the reference's MotionEstimate"""
    from me_util import REF_ME_JOB, ctu_jobs, ctu_layout
    g = geometry(cfg)
    C, depth, nref = cfg["C"], cfg["depth"], cfg["nref"]
    ys = synth_planes(nref + 1, g["planeRows"], g["stride"], depth, seed=1234)
    origin = g["padY"] * g["stride"] + g["padX"]
    item = ys[0].itemsize
    lay = ctu_layout(C, cfg["minCu"], cfg["rect"], cfg["amp"])
    rows = sorted(set([0, g["rows"] // 2, g["rows"] - 1]))
    if args_budget_small(cfg):
        rows = [g["rows"] // 2]
    mvp0 = np.zeros((len(lay), 2), dtype=np.int32)
    jobs = []
    for cy in rows:
        for cx in range(g["cols"]):
            j, searched = ctu_jobs(pkg, lay, C, cx, cy, cfg["W"], cfg["H"], mvp0, cfg["merange"])
            jobs.append(j[searched])
    job = np.concatenate(jobs)
    rj = np.zeros(len(job), dtype=REF_ME_JOB)
    for f in ("puX", "puY", "w", "h", "mvminX", "mvminY", "mvmaxX", "mvmaxY", "mvpX", "mvpY", "numCand", "mvc"):
        rj[f] = job[f]
    rj = rj[np.random.default_rng(3).permutation(len(rj))]
    frac = len(rows) / g["rows"]
    cur = ys[nref].ravel()
    refs = [ys[nref - 1 - r].ravel() for r in range(nref)]
    vp = lambda a, off=0: ctypes.c_void_p(a.ctypes.data + off * item)
    if cfg["csp"]:
        sc, rc = g["stride"] // 2, g["planeRows"] // 2
        oc = (g["padY"] // 2) * sc + g["padX"] // 2
        cbs = [p.ravel() for p in synth_planes(nref + 1, rc, sc, depth, seed=77, step=3)]
        crs = [p.ravel() for p in synth_planes(nref + 1, rc, sc, depth, seed=99, step=3)]

    def step():
        for r in range(nref):
            if cfg["csp"]:
                R.ref_me_batch_chroma(vp(cur, origin), vp(cbs[nref], oc), vp(crs[nref], oc), ctypes.c_ssize_t(g["stride"]), ctypes.c_ssize_t(sc),
                                      vp(refs[r], origin), vp(cbs[nref - 1 - r], oc), vp(crs[nref - 1 - r], oc), ctypes.c_ssize_t(g["stride"]), ctypes.c_ssize_t(sc),
                                      int(cfg["csp"]), ctypes.c_void_p(rj.ctypes.data), ctypes.c_int64(len(rj)), int(cfg["method"]), int(cfg["subme"]), int(cfg["merange"]), QP, 1, cores)
            else:
                R.ref_me_batch(vp(cur, origin), ctypes.c_ssize_t(g["stride"]), vp(refs[r], origin), ctypes.c_ssize_t(g["stride"]),
                               ctypes.c_void_p(rj.ctypes.data), ctypes.c_int64(len(rj)), int(cfg["method"]), int(cfg["subme"]), int(cfg["merange"]), QP, 1, cores)

    sample = "CTU rows %s of %d: %d PU searches per reference x %d references (the reference's MotionEstimate::motionEstimate, %d PUs per CTU)" % (
        rows, g["rows"], len(rj), nref, len(lay))
    return step, frac, sample, {"s_per_frame_all_cores": 0.0, "how": "not part of this config's step"}


def args_budget_small(cfg):
    """configs whose per-row CPU cost is large get one sampled CTU row (the run must end within minutes)"""
    return cfg["merange"] > 64 or cfg["subme"] >= 3


# =====================================================================================================================
# the B200 arm
# =====================================================================================================================
class Workload:
    """frames of one config resident in a pool; step(t) = one frame through the config's hot path"""

    def __init__(self, args, cfg, pkg, torch, ctx, dev, stream, rank, world, band=None):
        self.args, self.cfg, self.pkg, self.torch, self.ctx, self.dev, self.stream = args, cfg, pkg, torch, ctx, dev, stream
        self.rank, self.world = rank, world
        g = self.g = geometry(cfg)
        self.depth, self.C, self.nref = cfg["depth"], cfg["C"], cfg["nref"]
        self.item = 2 if self.depth > 8 else 1
        self.tdt = torch.int16 if self.item == 2 else torch.uint8
        self.NF = cfg["NF"]
        self.lam = pkg.lambda_for_qp(QP, self.depth)
        seed = 1234 + (rank if args.shard == "frames" else 0)        # ctu-rows: every rank holds the SAME frames
        ys = synth_planes(self.NF, g["planeRows"], g["stride"], self.depth, seed)
        self.host_y = [torch.from_numpy(f.view(np.int16) if self.item == 2 else f).pin_memory() for f in ys]
        self.pool = torch.empty((self.NF, g["planeRows"], g["stride"]), dtype=self.tdt, device=dev)
        for i, h in enumerate(self.host_y):
            self.pool[i].copy_(h)
        self.origin = (g["padY"] * g["stride"] + g["padX"]) * self.item            # bytes from a plane's base to pixel (0,0)
        self.plane_bytes = g["planeRows"] * g["stride"] * self.item
        self.csp = cfg["csp"]
        if self.csp:
            sc, rc = g["stride"] // 2, g["planeRows"] // 2
            self.sc, self.rc = sc, rc
            self.originC = ((g["padY"] // 2) * sc + g["padX"] // 2) * self.item
            self.poolC = []
            self.hostC = []
            for k, sd in enumerate((77, 99)):
                cs = synth_planes(self.NF, rc, sc, self.depth, sd + seed, step=3)
                hp = [torch.from_numpy(f.view(np.int16) if self.item == 2 else f).pin_memory() for f in cs]
                p = torch.empty((self.NF, rc, sc), dtype=self.tdt, device=dev)
                for i, h in enumerate(hp):
                    p[i].copy_(h)
                self.poolC.append(p); self.hostC.append(hp)
            self.planeC_bytes = rc * sc * self.item
        # CTU rows this rank searches (ctu-rows sharding: a band of ONE frame per rank)
        self.row0, self.rows = (0, g["rows"]) if band is None else band
        # streaming SAD outputs: [group][ref][grid]
        nctu = g["cols"] * g["rows"]
        self.sad_out = [torch.empty(max(1, self.NF // (self.nref + 1)) * self.nref * nctu * (64 // s) ** 2, dtype=torch.int32, device=dev) for s in (8, 16, 32, 64)]
        self.launches_extra = 0

    def frame(self, t):
        return t % self.NF

    def e2e_retire(self, drain=False):
        """a step's {mv, cost} records are read (x265b200_me_frame_host_end) once the NEXT step has been queued, so the host stays one step
        ahead of the device; every step's H2D and D2H still happen inside the timed region"""
        self.e2e_pending = getattr(self, "e2e_pending", 0) + (0 if drain else 1)
        while self.e2e_pending > (0 if drain else 1):
            self.ctx.me_frame_host_end()
            self.e2e_pending -= 1

    def refs_of(self, t):
        return [(t - 1 - r) % self.NF for r in range(self.nref)]

    def yptr(self, f, base=False):
        return self.pool[f].data_ptr() + (0 if base else self.origin)

    def sad_stream(self, groups, out=None):
        g = self.g
        out = out or self.sad_out
        self.ctx.sad_stream_dev(self.depth, self.pool.data_ptr() + self.origin, g["planeRows"] * g["stride"], g["stride"], g["padX"], g["padY"], g["planeRows"], self.NF,
                                (self.cfg["W"] + 63) // 64 if self.C != 64 else g["cols"], (g["rows"] * self.C) // 64, groups, self.nref, *[o.data_ptr() for o in out])

    def one_group(self, t):
        grp = np.zeros(1, dtype=self.pkg.SAD_GROUP)
        grp[0]["cur"] = self.frame(t)
        grp[0]["ref"][:self.nref] = self.refs_of(t)
        return grp

    def sad_bytes(self, ngroups):
        g = self.g
        w64, h64 = ((self.cfg["W"] + 63) // 64 if self.C != 64 else g["cols"]) * 64, ((g["rows"] * self.C) // 64) * 64
        nb = sum((w64 // s) * (h64 // s) for s in (8, 16, 32, 64))
        return ngroups * ((1 + self.nref) * w64 * h64 * self.item + self.nref * nb * 4)

    def join(self):
        pass


class MEWorkload(Workload):
    """configs 2 / 4 / 5: streaming SAD at the predictor + the frame search with the config's partition set"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        cfg, g, pkg, torch = self.cfg, self.g, self.pkg, self.torch
        self.npu = len(pkg.me_frame_layout(self.C, cfg["minCu"], cfg["rect"], cfg["amp"]))
        self.nout = self.nref * g["cols"] * self.rows * self.npu
        self.me_out = [torch.empty((self.nout, 3), dtype=torch.int32, device=self.dev) for _ in range(2)]
        self.res_h = [torch.empty((self.nout, 3), dtype=torch.int32).pin_memory() for _ in range(2)]
        self.params = dict(depth=self.depth, ctuSize=self.C, minCuSize=cfg["minCu"], rect=cfg["rect"], amp=cfg["amp"], picWidth=cfg["W"], picHeight=cfg["H"],
                           ctuCols=g["cols"], ctuRows=self.rows, marginX=g["padX"], marginY=g["padY"] + self.row0 * self.C, rowsTotal=g["planeRows"],
                           searchMethod=cfg["method"], subpelRefine=cfg["subme"], merange=cfg["merange"], csp=cfg["csp"], maxCand=0, maxSlices=1,
                           firstCtuRow=self.row0, sliceTotalRows=g["rows"])
        self.params["lambda"] = self.lam
        self.shY = self.row0 * self.C * g["stride"] * self.item
        self.shC = self.row0 * (self.C // 2) * (g["stride"] // 2) * self.item if self.csp else 0
        self.units = {"pu_searches": None, "pus_per_ctu": self.npu, "ctus": g["cols"] * self.rows, "references": self.nref}
        self.workload = cfg["name"] + "; step = streaming SAD at the predictor (all 2Nx2N levels x refs) + the frame search of every PU of the partition set, mvp = 0"

    def chroma_kw(self, t):
        if not self.csp:
            return {}
        f, refs = self.frame(t), self.refs_of(t)
        o = self.originC + self.shC
        return dict(curC=(self.poolC[0][f].data_ptr() + o, self.poolC[1][f].data_ptr() + o), curStrideC=self.sc,
                    refCb=[self.poolC[0][r].data_ptr() + o for r in refs], refCr=[self.poolC[1][r].data_ptr() + o for r in refs], refStrideC=self.sc)

    def step(self, t, k=0):
        if self.row0 == 0:
            self.sad_stream(self.one_group(t))
        f, refs = self.frame(t), self.refs_of(t)
        self.ctx.me_frame_ex_dev(self.params, self.yptr(f) + self.shY, self.g["stride"], [self.yptr(r) + self.shY for r in refs], self.g["stride"], self.me_out[k].data_ptr(),
                                 **self.chroma_kw(t))
        return self.me_out[k]

    def e2e_step(self, t, k=0):
        """the new frame arrives from pinned host memory through the host-buffer C-ABI entry; results land in pinned host memory"""
        f, refs = self.frame(t), self.refs_of(t)
        kw = self.chroma_kw(t)
        if self.csp:
            kw.update(hostC=(self.hostC[0][f].data_ptr(), self.hostC[1][f].data_ptr()), devCBase=(self.poolC[0][f].data_ptr(), self.poolC[1][f].data_ptr()),
                      bytesC=self.planeC_bytes)
        # _begin queues H2D (copy stream) -> search -> D2H (copy stream); the rest of the step is queued behind the search and
        # _end blocks until the {mv, cost} records are in host memory
        self.ctx.me_frame_ex_host_begin(self.params, self.yptr(f) + self.shY, self.g["stride"], [self.yptr(r) + self.shY for r in refs], self.g["stride"],
                                        self.host_y[f].data_ptr(), self.yptr(f, base=True), self.plane_bytes, self.me_out[k].data_ptr(), self.res_h[k].data_ptr(),
                                        self.nout * 12, **kw)
        if self.row0 == 0:
            self.sad_stream(self.one_group(t))
        self.e2e_retire()

    def h2d_bytes(self):
        return self.plane_bytes + (2 * self.planeC_bytes if self.csp else 0)

    def d2h_bytes(self):
        return self.nout * 12

    def count_units(self):
        out = self.me_out[0].cpu().numpy()
        self.units["pu_searches"] = int((out[:, 2] >= 0).sum())


class MixWorkload(Workload):
    """config 3: the full primitive mix + the lookahead on a second stream"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        pkg, torch, dev, ctx = self.pkg, self.torch, self.dev, self.ctx
        P = lambda t: t.data_ptr()
        self.P = P
        job_h = build_jobs(pkg)
        self.njobs = len(job_h)
        self.me_out = [torch.empty((self.njobs, 3), dtype=torch.int32, device=dev) for _ in range(2)]
        self.res_h = [torch.empty((self.njobs, 3), dtype=torch.int32).pin_memory() for _ in range(2)]
        n32 = (W // 32) * (CTU_ROWS * 2)
        self.n32 = n32
        self.qcoef = torch.empty(n32 * 1024, dtype=torch.int16, device=dev)
        self.recon = torch.empty((CTU_ROWS * CTU, W), dtype=torch.uint8, device=dev)
        self.qtab = torch.full((1024,), 26214, dtype=torch.int32, device=dev)          # quantScales[qp%6=0] flat list
        self.numsig = torch.empty((W // 4) * (CTU_ROWS * CTU // 4), dtype=torch.int32, device=dev)
        self.tu_sse = torch.empty((W // 4) * (CTU_ROWS * CTU // 4), dtype=torch.int64, device=dev)
        self.ip_jobs_h = {sz: interp_jobs(pkg, sz) for sz in LEVELS}
        self.ip_jobs_d = {sz: {k: torch.from_numpy(a.view(np.uint8).copy()).to(dev) for k, a in self.ip_jobs_h[sz].items() if len(a)} for sz in LEVELS}
        self.n_interp = sum(len(a) for d in self.ip_jobs_h.values() for a in d.values())
        self.pred = {sz: torch.zeros((CTU_ROWS * CTU, W), dtype=torch.uint8, device=dev) for sz in LEVELS}
        frame0 = self.host_y[0].numpy().ravel()
        self.nbr_d, self.n_intra = {}, 0
        for _, N, _ in INTRA_SIZES:
            a = neighbour_arrays(frame0, N)
            self.nbr_d[N] = torch.from_numpy(a).to(dev)
            self.n_intra += len(a)
        self.allangs_out = torch.empty(35 * W * CTU_ROWS * CTU, dtype=torch.uint8, device=dev)       # 35 modes x every pixel, reused per size
        self.units = {"pu_searches": self.njobs, "mc_interpolations": self.n_interp, "tu_per_size": {str(N): (W // N) * (CTU_ROWS * CTU // N) for _, N in TU_SIZES},
                      "intra_blocks_x35_modes": self.n_intra, "lookahead_list_searches_per_frame": 2 * (LA_BFRAMES + 1)}
        self.workload = WORKLOAD3
        self.init_lookahead()

    # ---- the lookahead half of the path, on its own stream / context --------------------------------------------------
    def init_lookahead(self):
        torch, pkg, dev = self.torch, self.pkg, self.dev
        # same priority as the main stream: measured 4.25 ms/step; giving the lookahead priority makes its CTAs displace frame-search
        # CTAs on every SM and costs more than it hides (5.2 ms/step), see DESIGN.md 7
        self.la_stream = self.args.la_stream if self.args.la_stream is not None else torch.cuda.Stream(device=dev)
        self.la_ctx = pkg.Ctx(dev.index, stream=self.la_stream.cuda_stream)
        Wc, Hc = W, CTU_ROWS * CTU
        mx, my = PAD, 80
        self.wcu, self.hcu = (Wc // 2 + 7) // 8, (Hc // 2 + 7) // 8
        self.lw, self.ll = self.wcu * 8, self.hcu * 8
        self.ls = (Wc // 2 + 2 * mx + 31) & ~31
        self.lmx, self.lmy = mx, my
        planesize, padoff, ncu = self.ls * (self.ll + 2 * my), self.ls * my + mx, self.wcu * self.hcu
        self.ncu = ncu
        NL = self.NF
        self.la_planes = torch.zeros((NL, 4, planesize), dtype=torch.uint8, device=dev)
        self.la_ptrs = np.array([[self.la_planes[i, k].data_ptr() + padoff for k in range(4)] for i in range(NL)], dtype=np.int64)
        self.la_plane_ptrs_d = torch.from_numpy(self.la_ptrs.copy()).to(dev)
        self.la_intra_cost = torch.empty((NL, ncu), dtype=torch.int32, device=dev)
        self.la_intra_ptrs_d = torch.tensor([self.la_intra_cost[i].data_ptr() for i in range(NL)], dtype=torch.int64, device=dev)
        self.la_im = torch.empty(ncu, dtype=torch.uint8, device=dev)
        self.la_lc0 = torch.empty(ncu, dtype=torch.int16, device=dev)
        self.la_rs0 = torch.empty(self.hcu, dtype=torch.int32, device=dev)
        self.la_sm0 = torch.empty(2, dtype=torch.int32, device=dev)
        self.la_lam = pkg.lambda_for_qp(12, 8)
        nslots = NL * 2 * (LA_BFRAMES + 2)
        self.la_mv = torch.zeros(nslots * ncu * 2, dtype=torch.int32, device=dev)
        self.la_mvc = torch.zeros(nslots * ncu, dtype=torch.int32, device=dev)
        ntr = LA_BATCH * (LA_BFRAMES + 1)
        self.la_lc = torch.empty(ntr * ncu, dtype=torch.int16, device=dev)
        self.la_rs = torch.empty(ntr * self.hcu, dtype=torch.int32, device=dev)
        self.la_sm = torch.empty(ntr * 4, dtype=torch.int32, device=dev)
        self.la_pending = []
        # every frame of the ring gets its lowres planes + intra costs once (the triples reach LA_BFRAMES + 1 frames back and forth)
        for i in range(NL):
            self.la_frame(i)
        self.la_ctx.sync()

    def la_frame(self, f):
        """Lowres::init + lowresIntraEstimate of frame f"""
        c, P = self.la_ctx, self.P
        c.lowres_init_dev(8, self.yptr(f), STRIDE, self.la_ptrs[f], self.ls, self.lw, self.ll, self.lmx, self.lmy)
        c.la_intra_dev(8, self.la_ptrs[f, 0], self.ls, self.wcu, self.hcu, None, 5 * int(self.la_lam), P(self.la_intra_cost[f]), P(self.la_im), P(self.la_lc0), P(self.la_rs0), P(self.la_sm0))

    def la_estimate(self, frames):
        """estimateFrameCost for the queued frames: per frame b the triples (b - d, b + d, b), d = 1 .. bframes + 1, both lists searched"""
        NL = self.NF
        # (the ring wraps: frames within bframes + 1 of its ends search the same amount of work around a legal centre)
        centre = lambda b: min(max(b, LA_BFRAMES + 1), NL - LA_BFRAMES - 2)
        wave = [(centre(b) - d, centre(b) + d, centre(b)) for b in frames for d in range(1, LA_BFRAMES + 2)]
        tr = np.zeros(len(wave), dtype=self.pkg.LA_TRIPLE)
        for t, (p0, p1, b) in enumerate(wave):
            d = t % (LA_BFRAMES + 1) + 1
            # frame indices must be ordered p0 < b < p1 for the kernel only as identities of planes; keep the ring's
            tr[t]["b"], tr[t]["p0"], tr[t]["p1"] = b, p0, p1
            for lst in (0, 1):
                tr[t]["mvSlot"][lst] = (b * 2 + lst) * (LA_BFRAMES + 2) + d
                tr[t]["doSearch"][lst] = 1
        P = self.P
        self.la_ctx.la_estimate_dev(8, P(self.la_plane_ptrs_d), self.ls, self.wcu, self.hcu, tr, P(self.la_mv), P(self.la_mvc), P(self.la_intra_ptrs_d), None,
                                    P(self.la_lc), P(self.la_rs), P(self.la_sm), self.la_lam, lookaheadSlices=LA_SLICES)

    def lookahead_step(self, t, ev_frame_ready):
        """called once per step: the new frame's lowres work now, the batch's list searches every LA_BATCH frames"""
        self.la_stream.wait_event(ev_frame_ready)
        f = self.frame(t)
        self.la_frame(f)
        self.la_pending.append(f)
        if len(self.la_pending) == LA_BATCH:
            self.la_estimate(self.la_pending)
            self.la_pending = []

    def join(self):
        if self.la_pending:                       # a partial batch at the end of the timed region still has to be searched
            self.la_estimate(self.la_pending)
            self.la_pending = []
        self.stream.wait_stream(self.la_stream)

    # ---- one frame ------------------------------------------------------------------------------------------------------
    def stages_after_search(self, t):
        ctx, P = self.ctx, self.P
        f, refs = self.frame(t), self.refs_of(t)
        cptr, r0 = self.yptr(f), self.yptr(refs[0])
        HH = CTU_ROWS * CTU
        ctx.interp_multi_dev(8, 8, [(kind, sz, sz, r0, STRIDE, P(self.pred[sz]), W, P(jd), len(self.ip_jobs_h[sz][kind]))
                                    for sz in LEVELS for kind, jd in self.ip_jobs_d[sz].items()])          # every level and filter kind in one launch
        for idx, N in TU_SIZES:
            qbits, add = quant_params(N)
            ctx.tu_pipeline_dev(idx, 8, 0, cptr, STRIDE, P(self.pred[16]), W, P(self.recon), W, W // N, HH // N, P(self.qtab), qbits, add, None, 40 << 5, 9,
                                P(self.qcoef), P(self.numsig), P(self.tu_sse))
        for _, N, log2N in INTRA_SIZES:
            ctx.intra_modes_dev(8, log2N, P(self.nbr_d[N]), P(self.allangs_out), int(N <= 16), self.nbr_d[N].shape[0])

    def step(self, t, k=0, events=None):
        torch = self.torch
        parts = os.environ.get("BENCH_PARTS", "both")          # diagnosis only: "main" / "la" time one half of the step alone
        if parts != "main":
            ev = torch.cuda.Event(); ev.record(self.stream)
            self.lookahead_step(t, ev)
        if parts == "la":
            return self.me_out[k]
        f, refs = self.frame(t), self.refs_of(t)
        if events is not None:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(self.stream)
        self.sad_stream(self.one_group(t))
        if events is not None:
            e[1].record(self.stream)
        self.ctx.me_frame_dev(8, self.yptr(f), STRIDE, [self.yptr(r) for r in refs], STRIDE, PAD, PAD, ROWS, CTU_COLS, CTU_ROWS, 15, None,
                              ME_HEX, SUBME, MERANGE, self.lam, self.P(self.me_out[k]))
        if events is not None:
            e[2].record(self.stream); events.append(e)
        self.stages_after_search(t)
        return self.me_out[k]

    def e2e_step(self, t, k=0):
        torch = self.torch
        f, refs = self.frame(t), self.refs_of(t)
        self.ctx.me_frame_host_begin(8, self.host_y[f].data_ptr(), self.plane_bytes, self.yptr(f, base=True), STRIDE, [self.yptr(r) for r in refs], STRIDE, PAD, PAD, ROWS,
                                     CTU_COLS, CTU_ROWS, 15, None, ME_HEX, SUBME, MERANGE, self.lam, self.P(self.me_out[k]), self.res_h[k].data_ptr(), self.njobs * 12)
        ev = torch.cuda.Event(); ev.record(self.stream)
        self.lookahead_step(t, ev)
        self.sad_stream(self.one_group(t))
        self.stages_after_search(t)
        self.e2e_retire()

    def h2d_bytes(self):
        return self.plane_bytes

    def d2h_bytes(self):
        return self.njobs * 12

    def count_units(self):
        pass


class LookaheadWorkload(MixWorkload):
    """--shard triples (config 3): ONE stream of frames goes through the lookahead half of the path on N GPUs.  The rank
    that owns a frame (t % N) runs Lowres::init + lowresIntraEstimate and broadcasts the four hpel planes and the intra costs; the
    frame-triples of every batch are dealt round-robin (they are independent, slicetype.cpp:1942-1968); rank 0 gathers lowresCosts,
    rowSatds and the frame sums.  Strong scaling: the work is fixed as N grows."""

    def __init__(self, args, cfg, pkg, torch, ctx, dev, stream, rank, world, band=None):
        Workload.__init__(self, args, cfg, pkg, torch, ctx, dev, stream, rank, world, band)
        import torch.distributed as dist
        self.dist = dist
        self.P = lambda t: t.data_ptr()
        self.njobs = 0
        self.me_out = [torch.empty((1, 3), dtype=torch.int32, device=dev) for _ in range(2)]
        self.units = {"lookahead_list_searches_per_frame": 2 * (LA_BFRAMES + 1), "frames_per_launch": LA_BATCH}
        self.workload = ("2160p-8bit-medium LOOKAHEAD half only (SURVEY.md 8e row 2): Lowres::init + lowresIntraEstimate per frame on the owning rank, broadcast of the "
                         "4 hpel planes + intra costs, the %d frame-triples of every %d-frame batch dealt round-robin over the ranks (--lookahead-slices 8), "
                         "gather of lowresCosts / rowSatds / sums on rank 0" % (LA_BATCH * (LA_BFRAMES + 1), LA_BATCH))
        self.n32 = (W // 32) * (CTU_ROWS * 2)
        self.init_lookahead()
        ntr = LA_BATCH * (LA_BFRAMES + 1)
        self.per_rank = pkg.triples_per_rank_max(ntr, world)
        n = self.per_rank
        self.res_send = torch.zeros(n * (self.ncu * 2 + self.hcu * 4 + 16), dtype=torch.uint8, device=dev)
        self.res_all = [torch.empty_like(self.res_send) for _ in range(world)] if rank == 0 and world > 1 else None
        self.res_host = torch.empty((world, self.res_send.numel()), dtype=torch.uint8).pin_memory() if rank == 0 else None
        self.e2e = False

    def la_estimate(self, frames):
        torch, dist = self.torch, self.dist
        NL = self.NF
        centre = lambda b: min(max(b, LA_BFRAMES + 1), NL - LA_BFRAMES - 2)
        wave = [(centre(b) - d, centre(b) + d, centre(b)) for b in frames for d in range(1, LA_BFRAMES + 2)]
        mine = self.pkg.triples_of_rank(len(wave), self.rank, self.world)
        tr = np.zeros(len(mine), dtype=self.pkg.LA_TRIPLE)
        for j, t in enumerate(mine):
            p0, p1, b = wave[t]
            d = t % (LA_BFRAMES + 1) + 1
            tr[j]["b"], tr[j]["p0"], tr[j]["p1"] = b, p0, p1
            for lst in (0, 1):
                tr[j]["mvSlot"][lst] = (b * 2 + lst) * (LA_BFRAMES + 2) + d
                tr[j]["doSearch"][lst] = 1
        P = self.P
        if len(mine):
            self.la_ctx.la_estimate_dev(8, P(self.la_plane_ptrs_d), self.ls, self.wcu, self.hcu, tr, P(self.la_mv), P(self.la_mvc), P(self.la_intra_ptrs_d), None,
                                        P(self.la_lc), P(self.la_rs), P(self.la_sm), self.la_lam, lookaheadSlices=LA_SLICES)
        with torch.cuda.stream(self.la_stream):
            n, ncu, hcu = len(mine), self.ncu, self.hcu
            o1, o2 = self.per_rank * ncu * 2, self.per_rank * (ncu * 2 + hcu * 4)
            self.res_send[:n * ncu * 2].copy_(self.la_lc[:n * ncu].view(torch.uint8))
            self.res_send[o1:o1 + n * hcu * 4].copy_(self.la_rs[:n * hcu].view(torch.uint8))
            self.res_send[o2:o2 + n * 16].copy_(self.la_sm[:n * 4].view(torch.uint8))
            if self.world > 1:
                dist.gather(self.res_send, self.res_all, dst=0)
            if self.e2e and self.rank == 0:
                src = self.res_all if self.world > 1 else [self.res_send]
                for r in range(self.world):
                    self.res_host[r].copy_(src[r], non_blocking=True)

    def step(self, t, k=0, events=None):
        self.e2e = False
        return self.one_frame(t, k)

    def one_frame(self, t, k):
        torch, dist = self.torch, self.dist
        f, owner = self.frame(t), t % self.world
        with torch.cuda.stream(self.la_stream):
            if owner == self.rank:
                if self.e2e:
                    self.pool[f].copy_(self.host_y[f], non_blocking=True)        # the new frame arrives from pinned host memory
                self.la_frame(f)
            if self.world > 1:
                dist.broadcast(self.la_planes[f], src=owner)
                dist.broadcast(self.la_intra_cost[f], src=owner)
        self.la_pending.append(f)
        if len(self.la_pending) == LA_BATCH:
            self.la_estimate(self.la_pending)
            self.la_pending = []
        return self.me_out[k]

    def e2e_step(self, t, k=0):
        self.e2e = True
        self.one_frame(t, k)

    def e2e_retire(self, drain=False):
        pass

    def h2d_bytes(self):
        return self.plane_bytes

    def d2h_bytes(self):
        return (LA_BFRAMES + 1) * (self.ncu * 2 + self.hcu * 4 + 16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--shard", default="frames", choices=["frames", "ctu-rows", "triples"])
    ap.add_argument("--la-sms", type=int, default=LA_SMS, help="config 3: SMs set aside for the lookahead stream (x265b200_sm_partition); 0 = no partition")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("x265-yuuki-asuna_b200")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.la_stream = None
    if args.config == 3 and args.la_sms > 0:
        # the lookahead keeps a few latency-bound warps per SM alive for milliseconds while the frame search fills every register
        # file it runs on: give each its own SMs (green contexts) so that they run side by side instead of in turn
        s_la, s_main, args.sm_split = pkg.sm_partition(local, args.la_sms)
        stream = torch.cuda.ExternalStream(s_main, device=local)
        args.la_stream = torch.cuda.ExternalStream(s_la, device=local)
    else:
        stream = torch.cuda.Stream(device=local)           # a real (non-NULL) stream shared by torch and the C ABI
    torch.cuda.set_stream(stream)
    ctx = pkg.Ctx(local, stream=stream.cuda_stream)       # fails loudly without the CUDA library
    assert stream.cuda_stream != 0 and ctx.stream == stream.cuda_stream
    dev = torch.device("cuda", local)
    cfg = CONFIGS[args.config]
    warm = max(args.warmup, 3)
    K = args.steps

    band = None
    if args.shard == "ctu-rows" and world > 1:
        band = pkg.band_rows(geometry(cfg)["rows"], rank, world)
    if args.config == 3 and args.shard == "ctu-rows":
        raise SystemExit("--shard ctu-rows applies to the frame-search configs (2, 4, 5)")
    if args.shard == "triples" and args.config != 3:
        raise SystemExit("--shard triples applies to the lookahead of config 3")
    wl = (LookaheadWorkload if args.shard == "triples" else (MixWorkload if args.config == 3 else MEWorkload))(args, cfg, pkg, torch, ctx, dev, stream, rank, world, band)

    # ---- N > 1: the path's exchange on a comm stream, overlapped with the next step ----------------------------
    comm = torch.cuda.Stream(device=local) if world > 1 else None
    g = wl.g
    if world > 1 and args.shard != "triples":
        if args.shard == "frames":
            # every rank needs the reference pixels the others produced: all_gather of the new luma plane; rank 0 collects the {mv,cost} records
            gathered = [torch.empty((world, wl.pool[0].shape[0], wl.pool[0].shape[1] * wl.item), dtype=torch.uint8, device=dev) for _ in range(2)]   # bytes: NCCL has no int16
        else:
            # ctu-rows: the bands of reconstructed rows are all_gathered (equal-sized chunks of the plane), results gathered
            rows_px = g["rows"] * wl.C
            chunk = pkg.plane_chunk(rows_px, world)
            gathered = [torch.empty((world, chunk, g["stride"] * wl.item), dtype=torch.uint8, device=dev) for _ in range(2)]
        nres = wl.me_out[0].shape[0]
        res_max = torch.tensor([nres], device=dev); dist.all_reduce(res_max, op=dist.ReduceOp.MAX); res_max = int(res_max)
        send = [torch.zeros((res_max, 3), dtype=torch.int32, device=dev) for _ in range(2)]
        res_all = [[torch.empty((res_max, 3), dtype=torch.int32, device=dev) for _ in range(world)] for _ in range(2)] if rank == 0 else [None, None]
    comm_events = [None, None]

    def exchange(t, k, out):
        if world == 1 or args.shard == "triples":          # the lookahead workload runs its own broadcast / gather
            return
        done = torch.cuda.Event(); done.record(stream)
        with torch.cuda.stream(comm):
            comm.wait_event(done)
            if args.shard == "frames":
                dist.all_gather_into_tensor(gathered[k], wl.pool[wl.frame(t)].view(torch.uint8))
            else:
                y0 = wl.g["padY"] + rank * gathered[k].shape[1]
                dist.all_gather_into_tensor(gathered[k], wl.pool[wl.frame(t)][y0:y0 + gathered[k].shape[1]].view(torch.uint8))
            send[k][:out.shape[0]].copy_(out)
            dist.gather(send[k], res_all[k], dst=0)
            comm_events[k] = torch.cuda.Event(); comm_events[k].record(comm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    t_first = wl.nref + LA_BFRAMES + 2

    def run_steps(first, count, events=None):
        for i in range(count):
            k = i & 1
            if comm_events[k] is not None:
                stream.wait_event(comm_events[k])          # result buffer k has been gathered
            out = wl.step(first + i, k, events) if isinstance(wl, MixWorkload) else wl.step(first + i, k)
            exchange(first + i, k, out)
        wl.join()
        if comm is not None:
            stream.wait_stream(comm)

    # ---- warm-up ---------------------------------------------------------------------------------------
    run_steps(t_first, warm)
    barrier()
    wl.count_units()
    launches0 = ctx.launches + (wl.la_ctx.launches if hasattr(wl, "la_ctx") else 0)

    sampler = ClockSampler(local); sampler.start()
    # ---- device-resident timing: K steps between two events on the main stream (the lookahead and comm streams are joined
    # before the closing event); frames come from a ring larger than L2, so no flush is needed
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run_steps(t_first + warm, K)
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = ctx.launches + (wl.la_ctx.launches if hasattr(wl, "la_ctx") else 0) - launches0

    # ---- per-kernel times of the two ME kernels (a separate short pass: events inside the step) ---------------------------
    stage_events = []
    if isinstance(wl, MixWorkload) and not isinstance(wl, LookaheadWorkload):
        run_steps(t_first, 4, stage_events)
        torch.cuda.synchronize()
    sad_ms = [e[0].elapsed_time(e[1]) for e in stage_events]
    me_ms = [e[1].elapsed_time(e[2]) for e in stage_events]

    # ---- roofline of the streaming ME-SAD kernel: ONE launch over `groups` disjoint (frame, references) groups -- every plane
    # byte comes from HBM exactly once (the ring is larger than L2 and each launch walks all of it), timed as a CUDA graph so
    # that the events bracket GPU time only
    groups = max(1, wl.NF // (wl.nref + 1))
    grp = np.zeros(groups, dtype=pkg.SAD_GROUP)
    for gi in range(groups):
        grp[gi]["cur"] = gi * (wl.nref + 1)
        grp[gi]["ref"][:wl.nref] = [gi * (wl.nref + 1) + 1 + r for r in range(wl.nref)]
    # the roofline legs time a kernel ALONE on the whole device: with an SM partition they get a plain stream / context of their own
    rf_stream, rf_ctx, step_ctx = stream, ctx, ctx
    if args.la_stream is not None:
        rf_stream = torch.cuda.Stream(device=local)
        rf_ctx = pkg.Ctx(local, stream=rf_stream.cuda_stream)
        wl.ctx = rf_ctx
        torch.cuda.set_stream(rf_stream)
    wl.sad_stream(grp); torch.cuda.synchronize()
    sad_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(sad_graph, stream=rf_stream):
        wl.sad_stream(grp)
    sad_loop_ms = []
    for rep in range(max(K, 5) + 1):
        r0 = torch.cuda.Event(enable_timing=True); r1 = torch.cuda.Event(enable_timing=True)
        r0.record(rf_stream); sad_graph.replay(); r1.record(rf_stream)
        torch.cuda.synchronize()
        if rep:                                  # first repetition is the warm-up
            sad_loop_ms.append(r0.elapsed_time(r1))
    del sad_graph

    dct_line = None
    if isinstance(wl, MixWorkload):
        # the transform stage measured the same way (DCT32 over the residual plane, 8 back-to-back launches on 8 planes)
        resid_ring = [torch.randint(-255, 256, (CTU_ROWS * CTU, W), dtype=torch.int16, device=dev) for _ in range(8)]
        coef_ring = [torch.empty(wl.n32 * 1024, dtype=torch.int16, device=dev) for _ in range(8)]

        def dct_loop():
            for kk in range(8):
                rf_ctx.dct_plane_dev(3, 8, resid_ring[kk].data_ptr(), W, W // 32, (CTU_ROWS * CTU) // 32, coef_ring[kk].data_ptr())
        dct_loop(); torch.cuda.synchronize()
        dct_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(dct_graph, stream=rf_stream):
            dct_loop()
        dct_ms = []
        for rep in range(4):
            r0 = torch.cuda.Event(enable_timing=True); r1 = torch.cuda.Event(enable_timing=True)
            r0.record(rf_stream); dct_graph.replay(); r1.record(rf_stream); torch.cuda.synchronize()
            if rep:
                dct_ms.append(r0.elapsed_time(r1) / 8)
        del dct_graph, resid_ring, coef_ring
        dct_line = dct_ms
    wl.ctx = step_ctx
    torch.cuda.set_stream(stream)

    # ---- end-to-end timing (host buffers in, results out, through the host-buffer C-ABI entry) ---------------------------
    def e2e_run(first, count):
        for i in range(count):
            k = i & 1
            if comm_events[k] is not None:
                stream.wait_event(comm_events[k])
            wl.e2e_step(first + i, k)
            exchange(first + i, k, wl.me_out[k])
        wl.e2e_retire(drain=True)
        wl.join()
        if comm is not None:
            stream.wait_stream(comm)
        torch.cuda.synchronize()

    e2e_run(t_first, 2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(t_first + 2, K)
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True; sampler.join(timeout=2)

    tt = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(tt[0]), float(tt[1])

    if rank == 0:
        pk, pk_kind = peaks()
        frames_done = K * (world if args.shard == "frames" else 1)
        fps = frames_done / (total_ms / 1e3)
        e2e_fps = frames_done / e2e_s
        sad_bytes = wl.sad_bytes(groups)
        sad_t = float(np.median(sad_loop_ms)) / 1e3
        achieved = sad_bytes / sad_t / 1e9
        traffic = SAD_STREAM_DRAM_BYTES.get((args.config, groups))
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": warm,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak" if args.shard == "frames" else "strong", "vs_baseline": None,
                "dtype": "u8" if cfg["depth"] == 8 else "u16", "data": "synthetic",
                "config": {"workload": wl.workload, "baseline_config": args.config, "units_per_step": wl.units,
                           "l2": "frames come from a ring of %d frames (%.0f MB) larger than L2; no flush" % (wl.NF, wl.NF * wl.plane_bytes / 1e6),
                           "sm_partition": ({"lookahead_sms": args.sm_split[0], "main_sms": args.sm_split[1], "how": "x265b200_sm_partition (CUDA green contexts): the lookahead stream and the main stream own disjoint SMs"} if args.la_stream is not None else None),
                           "parallelism": ("frame-parallel x%d (all_gather of the new plane + gather of results on a comm stream)" % world) if args.shard == "frames"
                                          else ("CTU-row bands of one frame x%d (all_gather of the bands' rows + gather of results)" % world) if args.shard == "ctu-rows"
                                          else ("lookahead frame-triples round-robin x%d (broadcast of the new frame's lowres planes + intra costs, gather of the cost records)" % world)},
                "clocks": sampler.summary(), "gpu_launches": int(launches),
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": wl.h2d_bytes(), "d2h_bytes_per_step": wl.d2h_bytes(),
                        "how": "the owning rank copies the new frame from pinned host memory, rank 0 copies the gathered cost records back to pinned host memory" if args.shard == "triples" else "x265b200_me_frame%s_host_begin / _host_end: per step the frame is copied from pinned host memory (copy stream), searched, and the {mv,cost} records are copied back to pinned host memory; the step's other stages are queued in between" % ("" if args.config == 3 else "_ex")},
                "roofline": {"kernel": "sad_stream_kernel (streaming ME SAD at the predictor: all 4 PU levels x %d references, %d frame groups in one launch, TMA ring)" % (wl.nref, groups),
                             "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
                             "traffic_src": "profiles/r02_sad_stream.txt (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of the same launch)" if traffic else None,
                             "peak_kind": pk_kind, "algorithmic_bytes": sad_bytes,
                             "bytes_formula": "groups x ((1 + refs) x W x H x sizeof(pixel) + 4 x refs x (n8 + n16 + n32 + n64)): the source frame is counted ONCE per group",
                             "us_per_launch": sad_t * 1e6,
                             "how": "one launch over %d disjoint groups (%d planes, %.0f MB > L2), CUDA-graph replay between events on the launching stream, median of %d"
                                    % (groups, groups * (wl.nref + 1), groups * (wl.nref + 1) * wl.plane_bytes / 1e6, len(sad_loop_ms)) +
                                    ("; in-step single-group launch: %.1f us" % (float(np.mean(sad_ms)) * 1e3) if sad_ms else "")}}
        if dct_line:
            dct_bytes = wl.n32 * 1024 * 2 * 2
            dct_t = float(np.mean(dct_line)) / 1e3
            line["roofline_dct32"] = {"kernel": "xform_mma_kernel<32,fwd> (IMMA)", "bound": "hbm", "achieved": dct_bytes / dct_t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                      "frac": dct_bytes / dct_t / 1e9 / pk["hbm_gbs"], "us_per_launch": dct_t * 1e6, "blocks": wl.n32}
        if me_ms:
            me_bytes = (1 + NREF) * W * (CTU_ROWS * CTU) + wl.njobs * 12
            me_t = float(np.mean(me_ms)) / 1e3
            line["roofline_me_search"] = {"kernel": "me_frame_kernel (TMA-staged windows; HEX + subme 2, %d searches)" % wl.njobs, "bound": "hbm", "achieved": me_bytes / me_t / 1e9,
                                          "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": me_bytes / me_t / 1e9 / pk["hbm_gbs"], "ms_per_step": me_t * 1e3,
                                          "note": "instruction-issue bound pattern search over smem-staged windows (DESIGN.md 5); HBM figure shown for scale only"}
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(args.config) if not os.environ.get("BENCH_NO_CPU") else {"skipped": "BENCH_NO_CPU (diagnosis run)"}
        print(json.dumps(line))
    if hasattr(wl, "la_ctx"):
        wl.la_ctx.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# DRAM bytes of one sad_stream launch under `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum), keyed by (config, groups)
SAD_STREAM_DRAM_BYTES = {(3, 8): 292216832}          # profiles/r02_sad_stream.txt: 278.55 MB read + 13.67 MB written; algorithmic 284.03 MB


def cpu_baseline(config):
    """bounded sample of the same step on the host cores through the compiled reference (kind 'reference')."""
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", str(config), "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900)
    try:
        ref = json.loads(out.stdout.strip().splitlines()[-1])
        return ref.get("cpu_baseline", {"unavailable": ref.get("unavailable")})
    except Exception as e:      # noqa: BLE001
        return {"unavailable": "reference arm failed: %s %s" % (e, out.stderr[-200:])}


if __name__ == "__main__":
    main()
