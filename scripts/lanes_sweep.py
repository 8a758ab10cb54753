import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.config2(1024, 10)
for fpt in (2, 1):
    ctx.set_option("formants_per_lane", fpt)
    for per_sm in ((4, 5, 6, 7, 8) if fpt == 2 else (4, 5, 6)):
        ctas = 148 * per_sm
        ctx.set_option("target_lanes", ctas * 32)
        plan = ctx.plan(elems, offs, vp); d = plan.device_output()
        best = 1e9
        for i in range(4):
            plan.launch(d); ctx.synchronize(); best = min(best, plan.timings()["formant_ms"])
        print("fpt", fpt, "CTAs/SM", per_sm, "formant_ms", round(best, 3))
        plan.close()
