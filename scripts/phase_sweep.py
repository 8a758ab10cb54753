"""phase_ms of the chunk-parallel carrier phase for a few chunk lengths (one library build per process)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
cfg = os.environ.get("GRAIL_CFG", "2")
elems, offs, vp = {"2": lambda: W.config2(1024, 10), "3": W.config3, "4": lambda: W.config4(4096)}[cfg]()
for pc in [int(x) for x in sys.argv[1:]] or [1024, 2048, 3072, 4096]:
    ctx.set_option("phase_chunk", pc)
    plan = ctx.plan(elems, offs, vp)
    d = plan.device_output()
    ts = []
    for i in range(4):
        plan.launch(d); ctx.synchronize()
        ts.append(plan.timings())
    t = ts[-1]
    st = plan.phase_stats()
    print(os.environ.get("GRAIL_CUDA_LIB", "default").split("/")[-1], "cfg", cfg, "pc", pc, "phase_ms %.4f total_ms %.4f" % (t["phase_ms"], t["total_ms"]),
          "walks/chunk %.3f rounds %d unproven %d" % (st["walks"] / max(1, st["chunks"]), st["repair_rounds"], st["unproven_utterances"]), flush=True)
    plan.close()
