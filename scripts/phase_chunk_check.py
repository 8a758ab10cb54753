"""GPU check of the chunk-parallel exact carrier phase against the oracle's trace and against the serial chains."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from oracle import oracle as O

ctx = g.Context(0)


def exact(a, b):
    return bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))


def check(name, elems, offs, vp, n_check=6):
    res = {}
    for mode in (1, 0):
        ctx.set_option("phase_mode", mode)
        plan = ctx.plan(elems, offs, vp)
        plan.launch()
        out = plan.read_output()
        f, p, s = plan.read_intermediates()
        res[mode] = (out, f, p, s, plan.out_offsets.copy(), plan.phase_stats(), plan.timings())
        plan.close()
    ctx.set_option("phase_mode", 1)
    out, f, p, s, oo, st, tm = res[1]
    ok_modes = exact(p, res[0][2]) and exact(s, res[0][3]) and exact(out, res[0][0])
    ok_or = True
    for u in list(range(min(n_check, len(offs) - 1))):
        want, tr, _ = O.synthesize(elems[offs[u]:offs[u + 1]], vp[u], trace=True)
        ok_or &= exact(p[oo[u]:oo[u + 1]], tr["carrier_phase"])
    print(f"{name}: phase/saw/audio identical to serial chains: {ok_modes}; phase == oracle on {n_check} utts: {ok_or}; "
          f"stats {st}; new {tm}; old phase_ms {res[0][6]['phase_ms']:.3f}", flush=True)
    return ok_modes and ok_or


v = g.voices.generic()
good = True
good &= check("kat", *W.from_phonemes([[0, 3], [0, 4, 3, 3, 4, 3, 3, 4, 3, 3], [0, 0, 0, 3], [3], [], [0, 0, 0, 3, 0, 0, 4, 0]], v,
                                      [0, 0, 0, 1, 2, 5]))
good &= check("config2x64", *W.config2(64, 10))
for rate in (16000.0, 48000.0):
    good &= check(f"config4x256@{rate}", *W.config4(256, rate))
good &= check("config3/20", *W.from_phonemes([W.config3_phonemes(60)], v, [0]), n_check=1)
for pc in (256, 1024, 4096):
    ctx.set_option("phase_chunk", pc)
    good &= check(f"config2x32 pc={pc}", *W.config2(32, 10), n_check=2)
ctx.set_option("phase_chunk", 2048)
print("ALL GOOD" if good else "FAILURES", flush=True)

# timings at full size
for label, (elems, offs, vp) in (("config2", W.config2(1024, 10)), ("config3", W.config3()), ("config4x4096", W.config4(4096))):
    for mode, pc in ((1, 1024), (1, 2048), (1, 4096), (0, 2048)):
        ctx.set_option("phase_mode", mode)
        ctx.set_option("phase_chunk", pc)
        plan = ctx.plan(elems, offs, vp)
        d = plan.device_output()
        for i in range(3):
            plan.launch(d); ctx.synchronize()
        t = plan.timings()
        print(label, "mode", mode, "pc", pc, {k: round(v, 4) for k, v in t.items()}, plan.phase_stats(),
              "samples/s %.3e" % (plan.total_samples / (t["total_ms"] * 1e-3)), flush=True)
        plan.close()
ctx.set_option("phase_mode", 1)
ctx.set_option("phase_chunk", 2048)
