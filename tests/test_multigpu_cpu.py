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
