import sys, os
os.environ["GRAIL_PSCAN_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from oracle import oracle as O
ctx = g.Context(0)
ctx.set_option("pscan_min_samples", 1)
ctx.set_option("pscan_cost_model", 0)
v = g.voices.generic()
for name, (elems, offs, vp) in [("voiced2", W.from_phonemes([[3, 4]], v, [3])), ("sil", W.from_phonemes([[0, 0]], v, [1])), ("mixed", W.from_phonemes([[0, 4, 3, 0, 0, 3]], v, [1]))]:
    plan = ctx.plan(elems, offs, vp)
    plan.launch()
    ctx.synchronize()
    print(name, plan.total_samples, plan.phase_scan_stats(), plan.timings())
    f, ph, saw = plan.read_intermediates()
    want, tr, _ = O.synthesize(elems, vp[0], trace=True)
    print("  phase exact:", np.array_equal(ph.view(np.uint32), tr["carrier_phase"].view(np.uint32)))
    plan.close()
