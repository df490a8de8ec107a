#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 x265 hot-path backend (contract in the task brief).

Workload ("step" = one 2160p 8-bit frame through the hot path, BASELINE.json configs[2] knobs:
HEX search, subme 2, merange 57, 3 references, 2Nx2N PUs 64/32/16/8 of every CTU, mvp = 0):
  1. sad_pred  : SAD at the predictor for every PU level x reference (grid-mode pixelcmp kernel; the
                 streaming "ME SAD" kernel whose HBM GB/s is the second half of BASELINE's metric)
  2. me_search : MotionEstimate::motionEstimate for every PU x reference (me_batch kernel)
  3. mc        : one 8-tap luma interpolation per PU and level (all 15 fractions)
  4. residual  : fused residual pipeline (fenc - pred -> DCT -> quant -> dequant -> IDCT -> recon -> SSE) on every 32/16/8/4 TU
  5. intra     : neighbour filter + all 35 modes on every 8/16/32 block
`value` = frames/s with inputs resident in HBM; `e2e` = the same through the C ABI with the new frame
coming from pinned HOST memory and the per-PU {MV,cost} results copied back, every step.

--impl reference times the reference's own C implementation (oracle/_ref, compiled from the
unmodified x265 sources) of the same step on the host cores, on a bounded sample of the frame.
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

W, H, CTU, PAD = 3840, 2160, 64, 128          # luma; PAD >= merange + 8 + halo, multiple of 64
STRIDE = W + 2 * PAD
ROWS = H + 2 * PAD + (64 - (H % 64)) % 64      # 34 CTU rows (2176) + margins
CTU_COLS, CTU_ROWS = W // CTU, (H + CTU - 1) // CTU
NREF, MERANGE, SUBME, QP = 3, 57, 2, 30
LEVELS = [64, 32, 16, 8]
# DRAM bytes one sad_pyramid launch (3 references, 2160p 8-bit) moved under `ncu --set full` (profiles/r01_pyramid_dct_v3.txt)
SAD_PYRAMID_DRAM_BYTES = 33437696
WORKLOAD = ("2160p-8bit-medium primitive mix (SURVEY.md 8d config 3): SAD at the predictor + HEX subme2 merange57 search of every 2Nx2N PU 64..8 x 3 refs; "
            "one 8-tap MC interpolation per PU and level (all 15 fractions); residual -> DCT/quant/dequant/IDCT on every 32/16/8/4 TU -> recon; "
            "intra neighbour smoothing + all 35 modes on every 8/16/32 block (fused)")
METRIC = "2160p preset-medium fps at 1/2/4/8 B200; ME SAD achieved HBM GB/s vs peak"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def synth_frames(nframes, seed=1234):
    """band-limited noise + per-frame global translation + noise."""
    rng = np.random.default_rng(seed)
    big = rng.integers(0, 256, (ROWS + 96, STRIDE + 96), dtype=np.uint8).astype(np.float32)
    k = np.ones(5, dtype=np.float32) / 5
    for ax in (0, 1):
        big = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), ax, big)
    big = (big - big.min()) / (big.max() - big.min()) * 255.0
    mrng = np.random.default_rng(5678)
    frames = []
    x, y = 48, 48
    for f in range(nframes):
        dx, dy = mrng.integers(-6, 7, 2)
        x = int(np.clip(x + dx, 0, 95)); y = int(np.clip(y + dy, 0, 95))
        fr = big[y:y + ROWS, x:x + STRIDE] + rng.normal(0, 2.0, (ROWS, STRIDE)).astype(np.float32)
        frames.append(np.clip(np.rint(fr), 0, 255).astype(np.uint8))
    return frames


def build_jobs(pkg, ctu_rows=None):
    """one job per (ref, level, PU): mvp = 0, window = +-merange."""
    rows = range(CTU_ROWS) if ctu_rows is None else ctu_rows
    js = []
    for r in range(NREF):
        for s in LEVELS:
            per = CTU // s
            for cy in rows:
                for cx in range(CTU_COLS):
                    for py in range(per):
                        y = cy * CTU + py * s
                        if y + s > H + (64 - H % 64) % 64:
                            continue
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


def run_reference(args, rank, world):
    """The reference's own C path (oracle/_ref) on host cores, bounded sample: one CTU row."""
    if rank != 0:
        return
    import oracle
    from me_util import REF_ME_JOB
    pkg = importlib.import_module("x265-yuuki-asuna_b200")
    R = oracle.ref(8)
    cores = os.cpu_count() or 1
    if R is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libx265ref8.so not present"}))
        return
    frames = synth_frames(NREF + 1)
    sample_rows = [CTU_ROWS // 2]
    job = build_jobs(pkg, sample_rows)
    frac = len(sample_rows) / CTU_ROWS
    origin = PAD * STRIDE + PAD
    cur = frames[NREF].ravel()

    def step():
        for r in range(NREF):
            jr = job[job["refIdx"] == r]
            rj = np.zeros(len(jr), dtype=REF_ME_JOB)
            for f in ("puX", "puY", "w", "h", "mvminX", "mvminY", "mvmaxX", "mvmaxY", "mvpX", "mvpY", "numCand", "mvc"):
                rj[f] = jr[f]
            ref = frames[NREF - 1 - r].ravel()
            R.ref_me_batch(ctypes.c_void_p(cur.ctypes.data + origin), ctypes.c_ssize_t(STRIDE),
                           ctypes.c_void_p(ref.ctypes.data + origin), ctypes.c_ssize_t(STRIDE),
                           ctypes.c_void_p(rj.ctypes.data), ctypes.c_int64(len(rj)), 1, SUBME, MERANGE, QP, 1, cores)
        # MC: one luma interpolation per PU of the sample row, all 15 fractions
        ref0 = frames[NREF - 1].ravel()
        for sz in LEVELS:
            part = R.ref_partition_from_sizes(sz, sz)
            for kind, ja in ip_jobs[sz].items():
                if len(ja):
                    R.ref_interp_batch(kind, part, ctypes.c_void_p(ref0.ctypes.data + origin), ctypes.c_ssize_t(STRIDE),
                                       ctypes.c_void_p(pred[sz].ctypes.data), ctypes.c_ssize_t(W), ctypes.c_void_p(ja.ctypes.data), ctypes.c_int64(len(ja)), cores)
        # residual of the sample row against the 16x16-level prediction, every TU size: the table entries chained as
        # Quant::transformNxN / invtransformNxN chain them (sub_ps, dct, quant, dequant, idct | DC fill | zero, add_ps, sse_pp)
        y0 = sample_rows[0] * CTU
        for idx, N in TU_SIZES:
            nb = (W // N) * (CTU // N)
            qbits, add = quant_params(N)
            R.ref_tu_pipeline(idx, 0, ctypes.c_void_p(cur.ctypes.data + origin + y0 * STRIDE), ctypes.c_ssize_t(STRIDE),
                              ctypes.c_void_p(pred[16].ctypes.data + y0 * W), ctypes.c_ssize_t(W), ctypes.c_void_p(recon.ctypes.data), ctypes.c_ssize_t(W),
                              W // N, CTU // N, ctypes.c_void_p(qtab.ctypes.data), qbits, add, None, 40 << 5, 9,
                              ctypes.c_void_p(tu_coef.ctypes.data), ctypes.c_void_p(tu_ns.ctypes.data), ctypes.c_void_p(tu_sse.ctypes.data), cores)
        # intra: filter + all 35 modes on every 8/16/32 block of the sample row
        for idx, N, _ in INTRA_SIZES:
            nb = len(nbr[N])
            R.ref_intra_batch(idx, ctypes.c_void_p(nbr[N].ctypes.data), ctypes.c_void_p(filt[N].ctypes.data), ctypes.c_void_p(intra_out[N].ctypes.data), ctypes.c_int64(nb), cores)

    ip_jobs = {sz: interp_jobs(pkg, sz, sample_rows) for sz in LEVELS}
    pred = {sz: np.zeros(CTU_ROWS * CTU * W, dtype=np.uint8) for sz in LEVELS}
    qtab = np.full(1024, 26214, dtype=np.int32)
    recon = np.zeros(CTU * W, dtype=np.uint8); tu_coef = np.zeros(CTU * W, dtype=np.int16)
    tu_ns = np.zeros((W // 4) * (CTU // 4), dtype=np.uint32); tu_sse = np.zeros((W // 4) * (CTU // 4), dtype=np.uint64)
    nbr = {N: neighbour_arrays(cur, N, sample_rows) for _, N, _ in INTRA_SIZES}
    filt = {N: np.empty_like(nbr[N]) for N in nbr}
    intra_out = {N: np.empty(len(nbr[N]) * 35 * N * N, dtype=np.uint8) for N in nbr}

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    fps = frac / dt
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "sample": "1 of %d CTU rows, scaled" % CTU_ROWS},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                             "sample": "CTU row %d of %d (x%d to a frame): %d PU searches (3 refs), %d MC interpolations, DCT/quant/dequant/IDCT of its 32/16/8/4 TUs, 35 intra modes on its 8/16/32 blocks; C table, no nasm asm" % (sample_rows[0], CTU_ROWS, CTU_ROWS, len(job), sum(len(a) for d in ip_jobs.values() for a in d.values()))},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
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
    stream = torch.cuda.Stream(device=local)               # a real (non-NULL) stream shared by torch and the C ABI
    torch.cuda.set_stream(stream)
    ctx = pkg.Ctx(local, stream=stream.cuda_stream)       # fails loudly without the CUDA library
    assert stream.cuda_stream != 0 and ctx.stream == stream.cuda_stream
    dev = torch.device("cuda", local)
    lam = pkg.lambda_for_qp(QP, 8)

    # ---- resident data: a ring of frames (> L2), static PU descriptors ------------------------------
    NF = 32                       # 32 x 9.9 MB padded planes = 318 MB > L2 (126 MB)
    frames_h = synth_frames(NF, seed=1234 + rank)
    pinned = [torch.from_numpy(f).pin_memory() for f in frames_h]
    ring = [torch.empty((ROWS, STRIDE), dtype=torch.uint8, device=dev) for _ in range(NF)]
    for d, h in zip(ring, pinned):
        d.copy_(h)
    job_h = build_jobs(pkg)
    job_h = job_h[np.argsort(-job_h["w"], kind="stable")]          # group by PU size (64, 32, 16, 8)
    njobs = len(job_h)
    job_bytes = job_h.dtype.itemsize
    level_ranges = {}
    for s in LEVELS:
        idx = np.nonzero(job_h["w"] == s)[0]
        level_ranges[s] = (int(idx[0]), int(len(idx)))
    jobs_d = torch.from_numpy(job_h.view(np.uint8).reshape(njobs, -1).copy()).to(dev)
    origin = PAD * STRIDE + PAD
    ref_ptr_table = torch.tensor([[ring[(t - 1 - r) % NF].data_ptr() + origin for r in range(NREF)] for t in range(NF)],
                                 dtype=torch.int64).to(dev)
    level_n = {s: (W // s) * ((CTU_ROWS * CTU) // s) for s in LEVELS}
    sad_out = {s: torch.empty(NREF * level_n[s], dtype=torch.int32, device=dev) for s in LEVELS}
    n32 = (W // 32) * (CTU_ROWS * 2)
    qcoef = torch.empty(n32 * 1024, dtype=torch.int16, device=dev)
    recon = torch.empty((CTU_ROWS * CTU, W), dtype=torch.uint8, device=dev)
    qtab = torch.full((1024,), 26214, dtype=torch.int32, device=dev)          # quantScales[qp%6=0] flat list
    numsig = torch.empty((W // 4) * (CTU_ROWS * CTU // 4), dtype=torch.int32, device=dev)
    tu_sse = torch.empty((W // 4) * (CTU_ROWS * CTU // 4), dtype=torch.int64, device=dev)
    # MC jobs (static), prediction planes, intra neighbour arrays (taken once from frame 0: static inputs) and outputs
    ip_jobs_h = {sz: interp_jobs(pkg, sz) for sz in LEVELS}
    ip_jobs_d = {sz: {k: torch.from_numpy(a.view(np.uint8).copy()).to(dev) for k, a in ip_jobs_h[sz].items() if len(a)} for sz in LEVELS}
    n_interp = sum(len(a) for d in ip_jobs_h.values() for a in d.values())
    pred = {sz: torch.zeros((CTU_ROWS * CTU, W), dtype=torch.uint8, device=dev) for sz in LEVELS}
    nbr_d, n_intra = {}, 0
    for _, N, _ in INTRA_SIZES:
        a = neighbour_arrays(frames_h[0].ravel(), N)
        nbr_d[N] = torch.from_numpy(a).to(dev)
        n_intra += len(a)
    allangs_out = torch.empty(35 * W * CTU_ROWS * CTU, dtype=torch.uint8, device=dev)       # 35 modes x every pixel, reused per size
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    me_out = torch.empty((njobs, 3), dtype=torch.int32, device=dev)
    res_h = torch.empty((njobs, 3), dtype=torch.int32).pin_memory()
    P = lambda t: t.data_ptr()

    sad_events = []
    me_events = []

    def hot_path(t, time_sad=False, out=None):
        out = me_out if out is None else out
        cur = ring[t % NF]
        refs = [ring[(t - 1 - r) % NF] for r in range(NREF)]
        ref_ptrs = ref_ptr_table[t % NF]
        cptr = P(cur) + origin
        # 1. SAD at the predictor for every PU level x ref: ONE streaming pass per reference (SAD pyramid)
        if time_sad:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        ctx.sad_pyramid_dev(8, cptr, STRIDE, P(ref_ptrs), NREF, STRIDE, CTU_COLS, CTU_ROWS, None,
                            P(sad_out[8]), P(sad_out[16]), P(sad_out[32]), P(sad_out[64]))
        if time_sad:
            e1.record(); sad_events.append((e0, e1))
        # 2. full motion search for every PU x ref: TMA-staged windows, one CTA per (CTU, ref)
        if time_sad:
            m0 = torch.cuda.Event(enable_timing=True); m1 = torch.cuda.Event(enable_timing=True); m0.record()
        ctx.me_frame_dev(8, cptr, STRIDE, [P(r) + origin for r in refs], STRIDE, PAD, PAD, ROWS, CTU_COLS, CTU_ROWS, 15, None,
                         pkg.ME_HEX, SUBME, MERANGE, lam, P(out))
        if time_sad:
            m1.record(); me_events.append((m0, m1))
        # 3. MC: one 8-tap interpolation per PU and level from reference 0 (HPP / VPP / HVPP by fraction)
        HH = CTU_ROWS * CTU
        r0 = P(refs[0]) + origin
        for sz in LEVELS:
            for kind, jd in ip_jobs_d[sz].items():
                ctx.interp_dev(kind, 8, 8, sz, sz, r0, STRIDE, P(pred[sz]), W, P(jd), len(ip_jobs_h[sz][kind]), 0)
        # 4. residual against the 16x16-level prediction -> DCT -> quant -> dequant -> IDCT -> recon -> SSE on every TU size:
        #    the fused residual pipeline (Quant::transformNxN + invtransformNxN chain, one launch per TU size)
        for idx, N in TU_SIZES:
            qbits, add = quant_params(N)
            ctx.tu_pipeline_dev(idx, 8, 0, cptr, STRIDE, P(pred[16]), W, P(recon), W, W // N, HH // N, P(qtab), qbits, add, None, 40 << 5, 9,
                                P(qcoef), P(numsig), P(tu_sse))
        # 5. intra: neighbour smoothing + DC + planar + the 33 angular modes of every 8/16/32 block, one fused launch per size
        #    (the prediction half of Search::estIntraPredQT, search.cpp:1358-1400)
        for _, N, log2N in INTRA_SIZES:
            ctx.intra_modes_dev(8, log2N, P(nbr_d[N]), P(allangs_out), int(N <= 16), nbr_d[N].shape[0])

    # ---- N > 1: the path's one real exchange: every rank needs the reference pixels the others
    # produced, and rank 0 collects the per-PU {mv,cost}.  One all_gather of the new luma plane + one gather.
    if world > 1:
        gathered = torch.empty((world * ROWS, STRIDE), dtype=torch.uint8, device=dev)   # concatenated form
        res_all = [torch.empty((njobs, 3), dtype=torch.int32, device=dev) for _ in range(world)] if rank == 0 else None

    def exchange(t, out=None):
        if world == 1:
            return
        dist.all_gather_into_tensor(gathered, ring[t % NF])
        dist.gather(me_out if out is None else out, res_all, dst=0)

    # end-to-end: every step uploads its frame from pinned host memory and reads its {mv,cost} results back.  The copies run on
    # a second stream so that the upload of frame t+1 and the read-back of step t overlap the kernels of the neighbouring step
    # (double-buffered result arrays); every copy of every step is inside the timed region.
    copy_stream = torch.cuda.Stream(device=local)
    out2 = [me_out, torch.empty_like(me_out)]
    res_h2 = [res_h, torch.empty((njobs, 3), dtype=torch.int32).pin_memory()]

    def e2e_run(first, count):
        def upload(t):
            with torch.cuda.stream(copy_stream):
                ring[t % NF].copy_(pinned[t % NF], non_blocking=True)           # H2D: the new frame
                ev = torch.cuda.Event(); ev.record(copy_stream)
            return ev
        ev_next = upload(first)
        d2h_ev = [None, None]
        for i in range(count):
            t, k = first + i, i & 1
            stream.wait_event(ev_next)
            if i + 1 < count:
                ev_next = upload(t + 1)
            if d2h_ev[k] is not None:
                stream.wait_event(d2h_ev[k])                                     # result buffer k has been read back
            hot_path(t, out=out2[k])
            exchange(t, out=out2[k])
            done = torch.cuda.Event(); done.record(stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                res_h2[k].copy_(out2[k], non_blocking=True)                     # D2H: {mvx, mvy, cost} per PU
                d2h_ev[k] = torch.cuda.Event(); d2h_ev[k].record(copy_stream)
        torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        hot_path(NREF + i)
        exchange(NREF + i)
    barrier()
    launches0 = ctx.launches

    sampler = ClockSampler(local); sampler.start()
    # ---- device-resident timing: per-step CUDA events, L2 flushed between steps -----------------------------
    evs = []
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 255)                                                   # > L2 (126 MB): evicts frames and tables
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        hot_path(NREF + i, time_sad=True)
        exchange(NREF + i)
        e1.record()
        evs.append((e0, e1))
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    sad_ms = [a.elapsed_time(b) for a, b in sad_events]
    me_ms = [a.elapsed_time(b) for a, b in me_events]
    launches = ctx.launches - launches0

    # ---- roofline of the streaming ME-SAD kernel: 8 back-to-back launches over DISJOINT frame sets (each plane
    # byte comes from HBM exactly once: L2 flushed first, 8 x 4 planes = 318 MB > L2), CUDA events around the loop
    groups = NF // (NREF + 1)
    grp_ptrs = torch.tensor([[ring[g * (NREF + 1) + 1 + r].data_ptr() + origin for r in range(NREF)] for g in range(groups)], dtype=torch.int64).to(dev)

    def sad_loop():
        for g in range(groups):
            ctx.sad_pyramid_dev(8, P(ring[g * (NREF + 1)]) + origin, STRIDE, P(grp_ptrs[g]), NREF, STRIDE, CTU_COLS, CTU_ROWS, None,
                                P(sad_out[8]), P(sad_out[16]), P(sad_out[32]), P(sad_out[64]))
    # the 8 launches are captured in a CUDA graph so the GPU is not waiting on Python/ctypes launch overhead
    sad_loop(); torch.cuda.synchronize()
    sad_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(sad_graph, stream=stream):
        sad_loop()
    sad_loop_ms = []
    for rep in range(max(args.steps, 3) + 1):
        flush.fill_(rep & 255)
        r0 = torch.cuda.Event(enable_timing=True); r1 = torch.cuda.Event(enable_timing=True)
        r0.record(); sad_graph.replay(); r1.record()
        torch.cuda.synchronize()
        if rep:                                  # first repetition is the warm-up
            sad_loop_ms.append(r0.elapsed_time(r1) / groups)

    # the transform stage measured the same way (DCT32 over the residual plane, 8 back-to-back launches on 8 planes)
    resid_ring = [torch.randint(-255, 256, (CTU_ROWS * CTU, W), dtype=torch.int16, device=dev) for _ in range(8)]
    coef_ring = [torch.empty(n32 * 1024, dtype=torch.int16, device=dev) for _ in range(8)]
    def dct_loop():
        for k in range(8):
            ctx.dct_plane_dev(3, 8, P(resid_ring[k]), W, W // 32, (CTU_ROWS * CTU) // 32, P(coef_ring[k]))
    dct_loop(); torch.cuda.synchronize()
    dct_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(dct_graph, stream=stream):
        dct_loop()
    dct_ms = []
    for rep in range(4):
        flush.fill_(rep)
        r0 = torch.cuda.Event(enable_timing=True); r1 = torch.cuda.Event(enable_timing=True)
        r0.record(); dct_graph.replay(); r1.record(); torch.cuda.synchronize()
        if rep:
            dct_ms.append(r0.elapsed_time(r1) / 8)
    del dct_graph, resid_ring, coef_ring

    # ---- end-to-end timing (host buffers in, results out) ---------------------------------------------------
    e2e_run(NREF, 2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(NREF, args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True; sampler.join(timeout=2)

    tt = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(tt[0]), float(tt[1])

    if rank == 0:
        pk, pk_kind = peaks()
        fps = world * args.steps / (total_ms / 1e3)
        e2e_fps = world * args.steps / e2e_s
        # roofline of the streaming ME-SAD kernel: algorithmic bytes = 2*W*H + nPU*4 per (level, ref) launch
        sad_launches = 1
        sad_bytes = (2 * W * (CTU_ROWS * CTU) + sum(level_n[s] * 4 for s in LEVELS)) * NREF
        sad_t = float(np.mean(sad_loop_ms)) / 1e3
        achieved = sad_bytes / sad_t / 1e9
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD, "units_per_step": {"pu_searches": njobs, "mc_interpolations": n_interp, "tu_per_size": {str(N): (W // N) * (CTU_ROWS * CTU // N) for _, N in TU_SIZES}, "intra_blocks_x35_modes": n_intra},
                           "l2": "512 MiB flush between timed steps", "parallelism": "frame-parallel x%d" % world},
                "clocks": sampler.summary(), "gpu_launches": int(launches),
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": ROWS * STRIDE, "d2h_bytes_per_step": njobs * 12},
                "roofline": {"kernel": "sad_pyramid_kernel (streaming ME SAD at the predictor, all 4 PU levels in one pass per reference)", "bound": "hbm", "achieved": achieved,
                             "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": SAD_PYRAMID_DRAM_BYTES,
                             "traffic_src": "profiles/r01_pyramid_dct_v3.txt (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch; the current frame is fetched from DRAM once and re-read by the 2 other references through L2, outputs stay in L2)",
                             "peak_kind": pk_kind, "launches_per_step": sad_launches, "us_per_launch": sad_t * 1e6,
                             "how": "%d back-to-back launches (one CUDA graph) on disjoint frame sets after an L2 flush, CUDA events; in-step (single launch between events): %.1f us" % (groups, float(np.mean(sad_ms)) * 1e3)}}
        dct_bytes = n32 * 1024 * 2 * 2
        dct_t = float(np.mean(dct_ms)) / 1e3
        line["roofline_dct32"] = {"kernel": "xform_mma_kernel<32,fwd> (IMMA)", "bound": "hbm", "achieved": dct_bytes / dct_t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                  "frac": dct_bytes / dct_t / 1e9 / pk["hbm_gbs"], "us_per_launch": dct_t * 1e6, "blocks": n32}
        me_bytes = NREF * (2 * W * (CTU_ROWS * CTU)) + njobs * 8
        me_t = float(np.mean(me_ms)) / 1e3
        line["roofline_me_search"] = {"kernel": "me_frame_kernel (TMA-staged windows; HEX + subme 2, %d searches)" % njobs, "bound": "hbm", "achieved": me_bytes / me_t / 1e9,
                                      "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": me_bytes / me_t / 1e9 / pk["hbm_gbs"], "ms_per_step": me_t * 1e3,
                                      "note": "instruction-issue/fetch-bound pattern search over smem-staged windows (DESIGN.md 5, profiles/r01_me_frame_v5.txt); HBM figure shown for scale only"}
        if world == 1:
            line["cpu_baseline"] = cpu_baseline()
            try:
                line["lookahead"] = lookahead_leg(pkg, ctx, ring, origin)
            except Exception as e:      # noqa: BLE001  (extra measurement; never lose the main line to it)
                line["lookahead"] = {"unavailable": str(e)[:200]}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def lookahead_leg(pkg, ctx, ring, origin, nframes=8, bframes=4):
    """The lookahead half of the path (NOT part of `value`): Lowres::init + lowresIntraEstimate per frame and the
    estimateFrameCost list searches `nframes` frames trigger at preset medium (bframes 4: 2 x 5 per frame), batched
    into ONE launch the way the encoder's lookahead queue allows.  Timed with host clocks around stream syncs."""
    Wc, Hc = W, CTU_ROWS * CTU
    mx, my = PAD, 80
    wcu, hcu = (Wc // 2 + 7) // 8, (Hc // 2 + 7) // 8
    lw, ll = wcu * 8, hcu * 8
    ls = (Wc // 2 + 2 * mx + 31) & ~31
    planesize, padoff, ncu = ls * (ll + 2 * my), ls * my + mx, wcu * hcu
    NF = nframes + 2 * (bframes + 1)
    planes = [[ctx.to_device(np.zeros(planesize, dtype=np.uint8)) for _ in range(4)] for _ in range(NF)]
    ptrs = np.array([[b.ptr + padoff for b in fr] for fr in planes], dtype=np.int64)
    lam = pkg.lambda_for_qp(12, 8)
    dIC = [ctx.empty(ncu * 4) for _ in range(NF)]
    dIM, dLC0, dRS0, dSm0 = ctx.empty(ncu), ctx.empty(ncu * 2), ctx.empty(hcu * 4), ctx.empty(8)

    def init_and_intra():
        for i in range(NF):
            ctx.lowres_init_dev(8, ring[i].data_ptr() + origin, STRIDE, ptrs[i], ls, lw, ll, mx, my)
            ctx.la_intra_dev(8, ptrs[i, 0], ls, wcu, hcu, None, 5 * int(lam), dIC[i], dIM, dLC0, dRS0, dSm0)
    init_and_intra(); ctx.sync()
    t0 = time.perf_counter(); init_and_intra(); ctx.sync()
    pre_ms = (time.perf_counter() - t0) * 1e3 / NF
    dPlanePtrs = ctx.to_device(ptrs)
    dIntraPtrs = ctx.to_device(np.array([b.ptr for b in dIC], dtype=np.int64))
    nslots = NF * 2 * (bframes + 2)
    dMv, dMvC = ctx.to_device(np.zeros(nslots * ncu * 2, dtype=np.int32)), ctx.to_device(np.zeros(nslots * ncu, dtype=np.int32))
    wave = [(b - dd, b + dd, b) for b in range(bframes + 1, bframes + 1 + nframes) for dd in range(1, bframes + 2)]
    tr = np.zeros(len(wave), dtype=pkg.LA_TRIPLE)
    for t, (p0, p1, b) in enumerate(wave):
        tr[t]["b"], tr[t]["p0"], tr[t]["p1"] = b, p0, p1
        for lst, dist in ((0, b - p0), (1, p1 - b)):
            tr[t]["mvSlot"][lst] = (b * 2 + lst) * (bframes + 2) + dist
            tr[t]["doSearch"][lst] = 1
    dLC, dRS, dSm = ctx.empty(len(wave) * ncu * 2), ctx.empty(len(wave) * hcu * 4), ctx.empty(len(wave) * 16)
    run = lambda: ctx.la_estimate_dev(8, dPlanePtrs, ls, wcu, hcu, tr, dMv, dMvC, dIntraPtrs, None, dLC, dRS, dSm, lam)
    run(); ctx.sync()
    t0 = time.perf_counter(); run(); ctx.sync()
    est_ms = (time.perf_counter() - t0) * 1e3
    for b in [x for fr in planes for x in fr] + dIC + [dIM, dLC0, dRS0, dSm0, dPlanePtrs, dIntraPtrs, dMv, dMvC, dLC, dRS, dSm]:
        b.free()
    per_frame = pre_ms + est_ms / nframes
    return {"included_in_value": False, "lowres": "%dx%d (%dx%d CUs)" % (lw, ll, wcu, hcu), "frames_batched": nframes,
            "list_searches": 2 * len(wave), "lowres_init_plus_intra_ms_per_frame": pre_ms, "estimate_launch_ms": est_ms,
            "ms_per_frame": per_frame, "frames_per_s": 1e3 / per_frame,
            "note": "estimateFrameCost is a dependent wavefront (~560 CU steps of ~30 us per (frame, list) field): one launch costs ~17-25 ms "
                    "whatever the batch, so throughput comes from batching the fields of several queued frames (DESIGN.md 5b, profiles/r01_lookahead.txt)"}


def cpu_baseline():
    """bounded sample of the same step on the host cores through the compiled reference (kind 'reference')."""
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    try:
        ref = json.loads(out.stdout.strip().splitlines()[-1])
        return ref.get("cpu_baseline", {"unavailable": ref.get("unavailable")})
    except Exception as e:      # noqa: BLE001
        return {"unavailable": "reference arm failed: %s %s" % (e, out.stderr[-200:])}


if __name__ == "__main__":
    main()
