import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
for n in (1, 16, 148, 296, 592, 1024):
    elems, offs, vp = W.config2(n, 10)
    plan = ctx.plan(elems, offs, vp); d = plan.device_output()
    for i in range(2): plan.launch(d); ctx.synchronize()
    t = plan.timings(); print(n, {k: round(v, 3) for k, v in t.items()}); plan.close()
