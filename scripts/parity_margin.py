"""worst-case parity margins (max-abs, SNR) of the CUDA path against the oracle over several configs"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from oracle import oracle as O
O.lib()
ctx = g.Context(0)
def run(label, elems, offs, vp, pick):
    plan = ctx.plan(elems, offs, vp); plan.launch(); out = plan.read_output(); oo = plan.out_offsets
    worst_abs, worst_snr = 0.0, 1e9
    for u in pick:
        want, _, _ = O.synthesize(elems[offs[u]:offs[u + 1]], vp[u])
        st = W.parity_stats(out[oo[u]:oo[u + 1]], want)
        worst_abs = max(worst_abs, st["max_abs"]); worst_snr = min(worst_snr, st["snr_db"])
    print(f"{label:28s} utts {len(pick):3d}  max_abs {worst_abs:.3e}  min SNR {worst_snr:.1f} dB")
    plan.close()
run("config2 (1024 x 10 ph)", *W.config2(), pick=[0, 1, 511, 1023])
run("config2 64 utts", *W.config2(64), pick=list(range(0, 64, 7)))
run("config4 2048 random voices", *W.config4(2048), pick=list(range(0, 2048, 97)))
for rate in (16000.0, 22050.0, 48000.0):
    run(f"config5 {int(rate)} Hz", *W.config2(64, 10, rate), pick=[0, 31, 63])
run("config3 40 phonemes", *W.config3(40), pick=[0])
