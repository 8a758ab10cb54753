"""config 4 slice (4096 utterances, per-utterance random voices): steady-state kernel times"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
for k, v in (a.split("=") for a in sys.argv[1:]):
    ctx.set_option(k, float(v))
elems, offs, vp = W.config4(4096)
plan = ctx.plan(elems, offs, vp)
d = plan.device_output()
for i in range(4):
    plan.launch(d); ctx.synchronize()
t = plan.timings()
print("config4/4096: samples %d  freq %.3f phase %.3f formant %.3f total %.3f ms  -> %.3e samples/s" % (
    plan.total_samples, t["frequency_ms"], t["phase_ms"], t["formant_ms"], t["total_ms"], plan.total_samples / t["total_ms"] * 1e3))
