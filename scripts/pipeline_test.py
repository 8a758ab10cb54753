import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream_handle)
elems, offs, vp = W.config2(1024, 10)
import itertools
for pipe, tl, lean in [(0, 0, -1), (1, 0, -1), (0, 23680, -1), (1, 23680, -1), (0, 18944, -1), (1, 18944, -1), (1, 14208, -1), (1, 21312, -1)]:
    if True:
        ctx.set_option("pipeline", pipe); ctx.set_option("target_lanes", tl); ctx.set_option("phase_lean", lean)
        plan = ctx.plan(elems, offs, vp)
        out = torch.empty(plan.total_samples, dtype=torch.float32, device="cuda")
        for _ in range(3): plan.launch(out.data_ptr())
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 20
        e0.record(stream)
        for _ in range(K): plan.launch(out.data_ptr())
        plan.join(); e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1) / K
        print("pipeline", pipe, "target_lanes", tl, "lean", lean, f"{ms:.3f} ms/step  {plan.total_samples/ms/1e-3:.3e} samples/s", {k: round(v, 3) for k, v in plan.timings().items()}, "chk", float(out[:220476].abs().sum()))
        plan.close()
