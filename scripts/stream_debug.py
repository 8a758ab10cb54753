import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from oracle import oracle as O
ctx = g.Context(0)
phon = [0, 4, 3, 0, 0, 3, 4, 4]
elems, offs, vp = W.from_phonemes([phon], g.voices.generic(), [11])
elems = elems.copy(); elems["length"] = np.array([0.5, 0.3, 0.5, 0.11, 0.5, 0.25, 0.5, 0.4], np.float32)
want, tr, _ = O.synthesize(elems, vp[0], trace=True)
for windows in ([777, 10000, 1, 255, 256, 257, 50000] * 3, [1000] * 30, [8] * 5 + [9] * 5 + [16, 17, 23, 24, 25]):
    st = ctx.stream(vp[0]); st.push(elems); st.finish()
    pos = 0
    for w in windows:
        x = st.pull(w)
        if len(x) == 0: break
        err = np.abs(x - want[pos:pos + len(x)]).max()
        print(f"window {w:6d} got {len(x):6d} at {pos:7d} phoneme {tr['phoneme_index'][pos]} max_err {err:.2e}")
        pos += len(x)
    st.close(); print("----")
