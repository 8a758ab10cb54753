"""Device-to-host copy ceiling of one box: each rank copies a config-2-sized f32 output (0.9 GB) from its GPU into its
own pinned host buffer -- first one rank at a time (the others idle), then all ranks at once -- with and without the
NUMA binding bench.py uses.  Prints one JSON line on rank 0: per-GPU GB/s alone, per-GPU GB/s concurrently, aggregate.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/d2h_concurrency.py

This is the measurement behind the end-to-end figures at N = 8 (bench.py e2e): the path is bound by this aggregate."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import bind_to_gpu_cpus

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 225767424                      # samples of one config-2 batch
res = {}
for numa in (0, 1):
    cpus = bind_to_gpu_cpus(local) if numa else None
    dev = torch.empty(N, dtype=torch.float32, device="cuda").normal_()
    host = torch.empty(N, dtype=torch.float32, pin_memory=True)
    host.zero_()                   # first touch under the current CPU mask
    s = torch.cuda.Stream()

    def copy(reps=3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for _ in range(reps):
                host.copy_(dev, non_blocking=True)
        s.synchronize()
        return N * 4 * reps / (time.perf_counter() - t0) / 1e9

    copy(1)
    alone = [0.0] * world
    for r in range(world):         # one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            alone[r] = copy()
    if world > 1:
        dist.barrier()
    together = copy()              # all ranks at once
    t = torch.tensor([alone[rank], together], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
    else:
        allv = [t]
    if rank == 0:
        a = [float(v[0]) for v in allv]
        c = [float(v[1]) for v in allv]
        res["numa_bound" if numa else "unbound"] = {"alone_gb_s": [round(x, 1) for x in a], "concurrent_gb_s": [round(x, 1) for x in c],
                                                     "aggregate_concurrent_gb_s": round(sum(c), 1), "cpus_in_mask": cpus}
    del dev, host
if rank == 0:
    res["n_gpus"] = world
    res["bytes_per_copy"] = N * 4
    res["f32_samples_per_s_ceiling"] = res["numa_bound"]["aggregate_concurrent_gb_s"] * 1e9 / 4
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
