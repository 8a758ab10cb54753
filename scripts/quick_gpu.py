"""ad-hoc GPU check: parity numbers on a few cases and first timings (not a bench)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from oracle import oracle as O

ctx = g.Context(0)
print("probe", ctx.probe_fp32_peak())
for name, ph, seed in [("sil_a", [0, 3], 0), ("ten", [0, 4, 3, 3, 4, 3, 3, 4, 3, 3], 0), ("sil3a", [0, 0, 0, 3], 0)]:
    elems, offs, vp = W.from_phonemes([ph], g.voices.generic(), [seed])
    plan = ctx.plan(elems, offs, vp)
    plan.launch()
    out = plan.read_output()
    f, p, s = plan.read_intermediates()
    want, tr, _ = O.synthesize(elems, vp[0], trace=True)
    print(name, len(out), len(want), "F exact", np.array_equal(f.view(np.uint32), tr["frequency"].view(np.uint32)),
          "phase exact", np.array_equal(p.view(np.uint32), tr["carrier_phase"].view(np.uint32)),
          W.parity_stats(out, want) if len(out) == len(want) else "LEN MISMATCH", plan.timings())
    if not np.array_equal(f.view(np.uint32), tr["frequency"].view(np.uint32)):
        bad = np.flatnonzero(f.view(np.uint32) != tr["frequency"].view(np.uint32))
        print("  first F mismatches", bad[:10], f[bad[:5]], tr["frequency"][bad[:5]])
    plan.close()
for n_utts in (64, 1024):
    elems, offs, vp = W.config2(n_utts, 10)
    t0 = time.time(); plan = ctx.plan(elems, offs, vp); t1 = time.time()
    d = plan.device_output()
    for i in range(3):
        plan.launch(d); ctx.synchronize()
        t = plan.timings()
        print(n_utts, "plan_s", round(t1 - t0, 4), t, "samples/s", plan.total_samples / (t["total_ms"] * 1e-3))
    plan.close()
