"""profiling driver: config 2 (optionally smaller) resident plan, a few launches (used under ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = g.Context(0)
if len(sys.argv) > 3: ctx.set_option("formants_per_lane", int(sys.argv[3]))
if len(sys.argv) > 4: ctx.set_option("target_lanes", 148 * 32 * int(sys.argv[4]))
elems, offs, vp = W.config2(n_utts, 10)
plan = ctx.plan(elems, offs, vp)
d = plan.device_output()
for i in range(reps):
    plan.launch(d)
    ctx.synchronize()
    print(plan.timings())
