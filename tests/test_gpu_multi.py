"""N > 1 on real GPUs (skipped with fewer than two): one batch split over two ranks by LPT on the exact sample counts,
each rank synthesizes its shard with no data-path collective, the outputs are gathered over NCCL and put into batch
order by grail_cuda_copy_segments.  The gathered batch must equal the single-GPU rendering of the whole batch:
sample counts exactly, F_t-dependent structure implicitly, samples to rounding level (the interpolating fast path is
chosen per warp, so a different batch composition may round differently: DESIGN.md section 2), every rank's own piece
bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus() -> int:
    import grail_rs_b200 as g
    return g._ffi.lib().grail_cuda_device_count()


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import grail_rs_b200 as g
    from grail_rs_b200 import sharding
    from grail_rs_b200 import workloads as W
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        elems, offs, vp = W.config4(96, first_utt=4000)
        e2, o2, v2 = W.config2(8, 3)                                  # a few default-voice utterances in the same batch
        elems = np.concatenate([elems, e2])
        offs = np.concatenate([offs, offs[-1] + o2[1:]]).astype(np.uint32)
        vp = np.concatenate([vp, v2])
        counts = g.count_samples(elems, offs, vp)
        assign = sharding.lpt_assign(counts, world)
        mine = assign[rank]
        se, so, sv = sharding.shard_batch(elems, offs, vp, mine)
        with g.Context(rank) as ctx:
            plan = ctx.plan(se, so, sv)
            out = torch.empty(plan.total_samples, dtype=torch.float32, device="cuda")
            plan.launch(out.data_ptr())
            ctx.synchronize()
            assert np.array_equal(np.diff(plan.out_offsets.astype(np.int64)), counts[mine].astype(np.int64))
            full = sharding.gather_outputs(out, counts[mine], assign, counts, ctx=ctx)
            torch.cuda.synchronize()
            all_off = np.concatenate([[0], np.cumsum(counts.astype(np.int64))])
            oo = plan.out_offsets
            own_ok = all(bool(torch.equal(full[all_off[u]:all_off[u + 1]], out[int(oo[k]):int(oo[k + 1])]))
                         for k, u in enumerate(mine.tolist()))
            plan.close()
            res = {"rank": rank, "own_ok": own_ok, "n": int(full.numel()), "total": int(counts.sum())}
            if rank == 0:
                whole = ctx.plan(elems, offs, vp)
                ref = torch.empty(whole.total_samples, dtype=torch.float32, device="cuda")
                whole.launch(ref.data_ptr())
                ctx.synchronize()
                res["counts_equal"] = bool(np.array_equal(np.diff(whole.out_offsets.astype(np.int64)), counts.astype(np.int64)))
                res["max_diff_vs_single_gpu"] = float((full - ref).abs().max().item())
                res["bit_equal_fraction"] = float((full == ref).float().mean().item())
                whole.close()
        q.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_rank_shard_and_nccl_gather():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r["rank"])
    for p in procs:
        p.join(120)
    print(res)
    for r in res:
        assert r["own_ok"] and r["n"] == r["total"], r
    assert res[0]["counts_equal"]
    assert res[0]["max_diff_vs_single_gpu"] <= 1e-5, res[0]   # rounding level: the two batch compositions interpolate over different 8 / 16-sample blocks
