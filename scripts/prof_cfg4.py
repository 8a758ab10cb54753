"""profiling driver: config 4 slice"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.config4(4096)
plan = ctx.plan(elems, offs, vp)
d = plan.device_output()
for i in range(3):
    plan.launch(d); ctx.synchronize()
print(plan.timings())
