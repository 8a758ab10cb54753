"""config 3 (one 10-minute utterance): fixed-point phase scan (k_ps_*) against the chunk-parallel path with CTA-wide scans"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.config3()
outs = []
for mode, pc in ((1, 0), (2, 0), (2, 2048), (2, 4096)):
    ctx.set_option("phase_mode", mode); ctx.set_option("phase_chunk", pc)
    plan = ctx.plan(elems, offs, vp)
    out = torch.empty(plan.total_samples, dtype=torch.float32, device="cuda")
    best = None
    for i in range(4):
        plan.launch(out.data_ptr()); ctx.synchronize()
        t = plan.timings()
        if best is None or t["total_ms"] < best["total_ms"]: best = t
    print(f"phase_mode={mode} phase_chunk={pc}: " + " ".join(f"{k}={v:.3f}" for k, v in best.items()), plan.phase_stats() if hasattr(plan, "phase_stats") else "")
    outs.append(out.clone()); plan.close()
for i in range(1, len(outs)):
    print(f"output {i} bit-equal to the fixed-point scan's: {bool(torch.equal(outs[0], outs[i]))}  max diff {float((outs[0]-outs[i]).abs().max()):.3e}")
