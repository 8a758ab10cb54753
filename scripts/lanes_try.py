"""config 2 with a few explicit lane targets (CTAs/SM of k_formant): lanes_try.py FPT ctas..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
ctx.set_option("formants_per_lane", int(sys.argv[1]))
elems, offs, vp = W.config2()
for ctas in [int(a) for a in sys.argv[2:]] or [0]:
    ctx.set_option("target_lanes", 148 * ctas * 32)
    plan = ctx.plan(elems, offs, vp)
    d = plan.device_output()
    ts = []
    for i in range(5):
        plan.launch(d); ctx.synchronize(); ts.append(plan.timings())
    t = ts[-1]
    print("FPT", sys.argv[1], "ctas/SM", ctas or "auto", "formant %.3f phase %.3f freq %.3f total %.3f" % (t["formant_ms"], t["phase_ms"], t["frequency_ms"], t["total_ms"]))
    plan.close()
