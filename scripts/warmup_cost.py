import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.config2(1024, 10)
for fpt in (1, 2):
    ctx.set_option("formants_per_lane", fpt)
    for D in (0.001, 3.45, 6.9, 13.8):
        ctx.set_option("warmup_nepers", D)
        plan = ctx.plan(elems, offs, vp); d = plan.device_output()
        for i in range(3): plan.launch(d); ctx.synchronize()
        print("fpt", fpt, "D", D, "formant_ms", round(plan.timings()["formant_ms"], 3))
        plan.close()
