"""one-shot config-2 batch, host records in / pinned f32 out: ms per call against the number of overlapped utterance groups"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.config2(1024, 10)
n = 1024 * 220476
oo = (np.arange(1025, dtype=np.uint64) * 220476)
pinned = ctx.pinned_empty(n)
dev = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
host = torch.from_numpy(pinned)
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); host.copy_(dev, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"bare D2H of the same 0.9 GB: {dt*1e3:.2f} ms ({n*4/dt/1e9:.1f} GB/s)")
for G in (-1, 1, 2, 3, 4, 5, 6):
    ctx.set_option("e2e_groups", G)
    ctx.synthesize_batch(elems, offs, vp, out=pinned, out_offsets=oo)
    t0 = time.perf_counter()
    for i in range(5): ctx.synthesize_batch(elems, offs, vp, out=pinned, out_offsets=oo)
    dt = (time.perf_counter() - t0) / 5
    print(f"e2e_groups {G:2d}: {dt*1e3:7.2f} ms/call  {n/dt:.3e} samples/s")
