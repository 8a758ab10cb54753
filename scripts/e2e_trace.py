import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.config2(1024, 10)
n = 1024 * 220476
oo = (np.arange(1025, dtype=np.uint64) * 220476)
pinned = ctx.pinned_empty(n)
for G in (3,):
    ctx.set_option("e2e_groups", G)
    for i in range(3):
        t0 = time.perf_counter(); ctx.synthesize_batch(elems, offs, vp, out=pinned, out_offsets=oo); dt = time.perf_counter() - t0
        print(f"G={G} call {i}: {dt*1e3:.2f} ms", file=sys.stderr)
