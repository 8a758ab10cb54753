"""k_phase cost against the carrier frequency (fewer wraps -> fewer redos): config 2 with the pitch scaled"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
for k, v in (a.split("=") for a in sys.argv[1:]):
    ctx.set_option(k, float(v))
for n_utts in (64, 1024):
    for scale in (1.0, 0.25, 0.02):
        elems, offs, vp = W.config2(n_utts)
        elems["elem"]["frequency"] *= np.float32(scale)
        vp["jitter_delta_frequency"] *= np.float32(scale)
        plan = ctx.plan(elems, offs, vp)
        d = plan.device_output()
        for i in range(4):
            plan.launch(d); ctx.synchronize()
        t = plan.timings()
        print("utts %4d pitch x%.2f  phase %.3f ms  (%.2f ns/sample)" % (n_utts, scale, t["phase_ms"], t["phase_ms"] * 1e6 / 220476))
        plan.close()
