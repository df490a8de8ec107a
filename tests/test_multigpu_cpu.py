"""CPU (gloo, world_size 2): the N > 1 host logic of bench.py -- frames are sharded round-robin over
ranks, every rank all_gathers the new reference planes and rank 0 gathers the per-PU results.  Uses the
oracle as the stand-in compute so the plumbing is testable without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from util import orc_cmp
    H, W = 32, 64
    nframes = 6
    rng = np.random.default_rng(7)
    frames = [rng.integers(0, 256, (H, W), dtype=np.int64).astype(np.uint8) for _ in range(nframes)]
    mine = list(range(rank, nframes, world))                      # round-robin frame ownership
    results = []
    prev = torch.zeros((world, H, W), dtype=torch.uint8)
    for step, f in enumerate(mine):
        cur = torch.from_numpy(frames[f].copy())
        flat = torch.empty((world * H, W), dtype=torch.uint8)
        dist.all_gather_into_tensor(flat, cur)
        gathered = flat.view(world, H, W)                 # reference pixels of every rank's new frame
        # every rank now sees the same set of planes
        for r in range(world):
            assert np.array_equal(gathered[r].numpy(), frames[step * world + r])
        ref = prev[(rank + 1) % world].numpy().ravel()             # a plane produced by ANOTHER rank in the previous step
        offs = np.arange(0, 8, dtype=np.int64) * 8
        sad = orc_cmp("sad", 8, 8, 8, frames[f].ravel(), W, ref, W, offs, offs)
        out = torch.tensor(sad, dtype=torch.int32)
        parts = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
        dist.gather(out, parts, dst=0)
        if rank == 0:
            results.append([p.tolist() for p in parts])
        prev = gathered
    if rank == 0:
        q.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_allgather_gather_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(res) == 3 and all(len(step) == world for step in res)
    # rank r's second-step SADs are against the plane the other rank produced in step 0
    assert all(len(v) == 8 for step in res for v in step)


def _band_worker(rank, world, port, q):
    """--shard ctu-rows: every rank holds the same frame, searches its band of CTU rows with the frame-relative clipping
    (firstCtuRow), contributes its rows of the plane to the all-gather and its padded records to the gather"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("x265-yuuki-asuna_b200")
    import me_util
    C, W, H, pad, merange, nref = 16, 96, 80, 48, 16, 2                   # 5 CTU rows: bands of 2 and 3 rows
    cols, rows = W // C, H // C
    frames, stride, _, origin = me_util.synth_sequence(W, H, pad, pad, 8, nref + 1, seed=5, max_motion=6)
    layout = me_util.ctu_layout(C, 8, False, False)
    npu = len(layout)
    mvp = np.zeros((npu, 2), dtype=np.int32)

    def search(row0, nrows):
        out = np.zeros((nref, nrows * cols * npu, 3), dtype=np.int32)
        for r in range(nref):
            for cy in range(row0, row0 + nrows):
                for cx in range(cols):
                    job, searched = me_util.ctu_jobs(pkg, layout, C, cx, cy, W, H, mvp, merange)   # picH of the WHOLE frame
                    mx, my, cost = me_util.ref_me(8, frames[0], frames[1 + r], stride, origin, job, 1, 2, merange, 30)
                    base = ((cy - row0) * cols + cx) * npu
                    out[r, base:base + npu] = np.stack([mx, my, cost], axis=1)
        return out.reshape(-1, 3)

    lo, n = pkg.band_rows(rows, rank, world)
    mine = search(lo, n)
    assert mine.shape[0] == pkg.records_per_band(rows, cols, npu, nref, rank, world)
    res_max = torch.tensor([mine.shape[0]]); dist.all_reduce(res_max, op=dist.ReduceOp.MAX); res_max = int(res_max)
    send = torch.zeros((res_max, 3), dtype=torch.int32); send[:mine.shape[0]] = torch.from_numpy(mine)
    parts = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, parts, dst=0)
    # the bands' rows of the (reconstructed) plane: equal chunks, the last one padded
    chunk = pkg.plane_chunk(H, world)
    plane = torch.from_numpy(frames[0].reshape(-1, stride)[pad:pad + H].copy())
    padded = torch.zeros((world * chunk, stride), dtype=torch.uint8); padded[:H] = plane
    flat = torch.empty((world * chunk, stride), dtype=torch.uint8)
    dist.all_gather_into_tensor(flat, padded[rank * chunk:(rank + 1) * chunk].contiguous())
    assert torch.equal(flat[:H], plane)
    if rank == 0:
        whole = search(0, rows).reshape(nref, -1, 3)
        got = pkg.assemble_bands([p.numpy() for p in parts], rows, cols, npu, nref, world)
        q.put(bool(np.array_equal(got, whole)))
    dist.barrier()
    dist.destroy_process_group()


def test_ctu_row_bands_gather_gloo():
    from util import oracle
    if oracle.ref(8) is None:
        pytest.skip("oracle/_ref not built")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "bands searched with firstCtuRow and put back in rank order differ from the whole-frame search"


def test_band_partition_arithmetic():
    import importlib
    pkg = importlib.import_module("x265-yuuki-asuna_b200")
    for rows in (1, 5, 17, 34, 68):
        for world in (1, 2, 3, 4, 8):
            bands = [pkg.band_rows(rows, r, world) for r in range(world)]
            assert bands[0][0] == 0 and sum(n for _, n in bands) == rows
            assert all(bands[i][0] + bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            assert max(n for _, n in bands) - min(n for _, n in bands) <= 1
            assert pkg.plane_chunk(rows * 64, world) * world >= rows * 64


def _triple_worker(rank, world, port, q):
    """--shard triples: the frame-triples of a batch dealt round-robin; each rank's records (here: a function of the triple) are padded
    to the largest share and gathered; rank 0 puts them back in batch order"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import importlib
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("x265-yuuki-asuna_b200")
    ntr, reclen = 25, 7                                                    # 5 frames x (bframes 4 + 1) triples
    wave = [(b - d, b + d, b) for b in range(10, 15) for d in range(1, 6)]
    rec = lambda t: np.array([wave[t][0], wave[t][1], wave[t][2], t, t * t, 1, 2], dtype=np.int32)
    mine = pkg.triples_of_rank(ntr, rank, world)
    n = pkg.triples_per_rank_max(ntr, world)
    send = torch.zeros((n, reclen), dtype=torch.int32)
    for j, t in enumerate(mine):
        send[j] = torch.from_numpy(rec(t))
    parts = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, parts, dst=0)
    # the owning rank's new lowres planes reach every rank
    plane = torch.full((4, 64), rank, dtype=torch.uint8)
    for owner in range(world):
        buf = plane.clone() if owner == rank else torch.empty_like(plane)
        dist.broadcast(buf, src=owner)
        assert int(buf[0, 0]) == owner
    if rank == 0:
        got = pkg.assemble_triples([p.numpy() for p in parts], ntr, world)
        q.put(bool(np.array_equal(got, np.stack([rec(t) for t in range(ntr)]))))
    dist.barrier()
    dist.destroy_process_group()


def test_lookahead_triples_round_robin_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_triple_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_triple_partition_arithmetic():
    import importlib
    pkg = importlib.import_module("x265-yuuki-asuna_b200")
    for n in (0, 1, 5, 80, 81):
        for world in (1, 2, 3, 8):
            shares = [pkg.triples_of_rank(n, r, world) for r in range(world)]
            assert sorted(t for s in shares for t in s) == list(range(n))
            assert max(len(s) for s in shares) == pkg.triples_per_rank_max(n, world) or n == 0
