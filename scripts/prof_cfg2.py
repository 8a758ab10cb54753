"""one config-2 plan, a few launches (for ncu launch lists / captures)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = os.environ.get("GRAIL_CFG", "2")
elems, offs, vp = {"2": lambda: W.config2(1024, 10), "3": W.config3, "4": lambda: W.config4(4096)}[cfg]()
plan = ctx.plan(elems, offs, vp)
d = plan.device_output()
for i in range(n):
    plan.launch(d)
    ctx.synchronize()
print(plan.timings(), plan.phase_stats())
