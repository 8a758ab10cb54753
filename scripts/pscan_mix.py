"""Cost-model check: batches mixing long and short utterances, scan on / off."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
v = g.voices.generic()
def run(label, elems, offs, vp):
    for mn in (1 << 18, 0xFFFFFFFF):
        ctx.set_option("pscan_min_samples", mn)
        plan = ctx.plan(elems, offs, vp)
        d = plan.device_output()
        for i in range(3):
            plan.launch(d); ctx.synchronize()
        t = plan.timings(); ps = plan.phase_scan_stats()
        print(label, "scan" if mn < 1e9 else "chain", "phase_ms %.3f total %.3f" % (t["phase_ms"], t["total_ms"]), ps)
        plan.close()
e3, o3, v3 = W.config3(60)      # one 30 s utterance
run("1x30s", e3, o3, v3)
e3, o3, v3 = W.config3(20)      # one 10 s utterance
run("1x10s", e3, o3, v3)
# 4 x 10 s + 200 short
parts = [W.config3(20) for _ in range(4)]
es, os_, vs = W.config4(200)
elems = np.concatenate([p[0] for p in parts] + [es])
offs = [0]
for p in parts: offs.append(offs[-1] + len(p[0]))
base = offs[-1]
for i in range(1, len(os_)): offs.append(base + int(os_[i]))
vp = np.concatenate([p[2] for p in parts] + [vs])
run("4x10s+200", elems, np.array(offs, dtype=np.uint32), vp)
# 40 x 10 s
parts = [W.config3(20) for _ in range(40)]
elems = np.concatenate([p[0] for p in parts]); offs = np.arange(41, dtype=np.uint32) * len(parts[0][0]); vp = np.concatenate([p[2] for p in parts])
run("40x10s", elems, offs, vp)
