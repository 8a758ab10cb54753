import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from oracle import oracle as O
ctx = g.Context(0)
elems, offs, vp = W.config2(1024, 10)
want = O.synthesize(elems[offs[5]:offs[6]], vp[5])[0]
for fpt in (1, 2):
    ctx.set_option("formants_per_lane", fpt)
    for tl in (0, 16384, 24576, 32768, 49152):
        ctx.set_option("target_lanes", tl)
        plan = ctx.plan(elems, offs, vp); d = plan.device_output()
        for i in range(3): plan.launch(d); ctx.synchronize()
        t = plan.timings()
        out = plan.read_output(); oo = plan.out_offsets
        st = W.parity_stats(out[oo[5]:oo[6]], want)
        print("fpt", fpt, "target_lanes", tl, {k: round(v, 3) for k, v in t.items()}, "max_abs %.2e snr %.1f" % (st["max_abs"], st["snr_db"]))
        plan.close()
