"""parity vs warm-up depth (nepers) and chunk length: how much zero-state error survives"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from oracle import oracle as O
ctx = g.Context(0)
elems, offs, vp = W.config2(4, 10)
e4, o4, v4 = W.config4(6, first_utt=40)
want = [O.synthesize(elems[offs[u]:offs[u+1]], vp[u])[0] for u in range(4)]
want4 = [O.synthesize(e4[o4[u]:o4[u+1]], v4[u])[0] for u in range(6)]
for chunk in (1 << 22, 6912, 2048):
    ctx.set_option("min_chunk", chunk); ctx.set_option("target_lanes", 1 if chunk > 100000 else 1 << 20)
    for D in (9.2, 10.4, 11.5, 13.8):
        ctx.set_option("warmup_nepers", D)
        out, oo = ctx.synthesize_batch(elems, offs, vp)
        st = [W.parity_stats(out[oo[u]:oo[u+1]], want[u]) for u in range(4)]
        out4, oo4 = ctx.synthesize_batch(e4, o4, v4)
        st4 = [W.parity_stats(out4[oo4[u]:oo4[u+1]], want4[u]) for u in range(6)]
        print(f"chunk {chunk:8d} D {D:5.1f}  default voice: max_abs {max(s['max_abs'] for s in st):.2e} snr {min(s['snr_db'] for s in st):6.1f} dB"
              f" | random voices: max_abs {max(s['max_abs'] for s in st4):.2e} snr {min(s['snr_db'] for s in st4):6.1f} dB")
        if chunk > 100000: break
