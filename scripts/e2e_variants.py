import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.config2(1024, 10)
n = 1024 * 220476
oo = (np.arange(1025, dtype=np.uint64) * 220476)
pinned = ctx.pinned_empty(n)
pageable = np.empty(n, np.float32)
for name, buf, zc in (("pinned zero-copy", pinned, 1), ("pinned memcpy", pinned, 0), ("pageable", pageable, 0)):
    ctx.set_option("zero_copy_out", zc)
    ctx.synthesize_batch(elems, offs, vp, out=buf, out_offsets=oo)
    t0 = time.perf_counter()
    for i in range(3): ctx.synthesize_batch(elems, offs, vp, out=buf, out_offsets=oo)
    dt = (time.perf_counter() - t0) / 3
    print(f"{name:18s} {dt*1e3:7.2f} ms/step  {n/dt:.3e} samples/s  checksum {float(np.abs(buf[:220476]).sum()):.4f}")
