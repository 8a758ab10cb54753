"""where an end-to-end step goes: plan build (host schedule + H2D), kernels, D2H -- through the plan API, and the
one-shot grail_cuda_synthesize_batch for comparison"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.config2()
n = 1024 * 220476
host = ctx.pinned_empty(n, np.float32)
oo = None
for rep in range(3):
    t0 = time.perf_counter()
    plan = ctx.plan(elems, offs, vp)
    ctx.synchronize()
    t1 = time.perf_counter()
    plan.launch(); ctx.synchronize()
    t2 = time.perf_counter()
    plan.read_output(out=host)
    t3 = time.perf_counter()
    oo = plan.out_offsets.copy()
    plan.close()
    t4 = time.perf_counter()
    print("plan %.2f ms  kernels %.2f ms  d2h %.2f ms  close %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
for rep in range(3):
    t0 = time.perf_counter()
    ctx.synthesize_batch(elems, offs, vp, out=host, out_offsets=oo)
    t1 = time.perf_counter()
    print("synthesize_batch %.2f ms  -> %.3e samples/s" % ((t1 - t0) * 1e3, n / (t1 - t0)))
