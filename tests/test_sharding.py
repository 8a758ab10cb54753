"""N > 1 host logic on CPU: LPT sharding by exact sample counts and the output gather, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import grail_rs_b200 as g
from grail_rs_b200 import sharding
from grail_rs_b200 import workloads as W


def test_lpt_assign_balances_and_covers():
    rng = np.random.default_rng(0)
    counts = rng.integers(1000, 100000, 200)
    for world in (1, 2, 4, 8):
        a = sharding.lpt_assign(counts, world)
        assert sorted(np.concatenate(a).tolist()) == list(range(200))
        loads = [counts[x].sum() for x in a]
        assert max(loads) - min(loads) <= counts.max()


def test_shard_batch_roundtrip():
    elems, offs, vp = W.config4(9)
    counts = g.count_samples(elems, offs, vp)
    parts = sharding.lpt_assign(counts, 2)
    seen = []
    for mine in parts:
        e, o, v = sharding.shard_batch(elems, offs, vp, mine)
        assert np.array_equal(g.count_samples(e, o, v), counts[mine])
        assert e.tobytes() == np.concatenate([elems[offs[u]:offs[u + 1]] for u in mine]).tobytes()
        seen += mine.tolist()
    assert sorted(seen) == list(range(9))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        elems, offs, vp = W.config4(7)
        counts = g.count_samples(elems, offs, vp).astype(np.int64) // 1000     # small stand-in lengths
        assign = sharding.lpt_assign(counts, world)
        mine = assign[rank]
        # stand-in for the device output of this rank's shard: sample i of utterance u is u + i/1e6
        local = np.concatenate([u + np.arange(counts[u]) * 1e-6 for u in mine]).astype(np.float32) if len(mine) else np.zeros(0, np.float32)
        full = sharding.gather_outputs(local, counts[mine], assign, counts).numpy()
        want = np.concatenate([u + np.arange(counts[u]) * 1e-6 for u in range(7)]).astype(np.float32)
        q.put((rank, bool(np.array_equal(full, want))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_outputs_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
