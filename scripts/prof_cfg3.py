"""Config 3 (one 10-minute utterance): steady-state timings and phase-scan statistics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.config3(1200)
plan = ctx.plan(elems, offs, vp)
d = plan.device_output()
for i in range(4):
    plan.launch(d); ctx.synchronize(); print(plan.timings(), plan.phase_scan_stats())
