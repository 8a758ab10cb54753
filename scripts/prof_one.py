"""profiling driver: ONE 5 s utterance (the serial carrier-phase chain in isolation)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
elems, offs, vp = W.from_phonemes([[0, 4, 3, 3, 4, 3, 3, 4, 3, 3]], g.voices.generic(), [0])
plan = ctx.plan(elems, offs, vp)
d = plan.device_output()
for i in range(3):
    plan.launch(d); ctx.synchronize(); print(plan.timings())
