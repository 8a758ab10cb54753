import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from oracle import oracle as O
ctx = g.Context(0)
phon = [3, 4, 3]
elems, offs, vp = W.from_phonemes([phon], g.voices.generic(), [11])
want, tr, _ = O.synthesize(elems, vp[0], trace=True)
np.set_printoptions(precision=5, linewidth=200)
for first in (8, 16, 24, 25, 255, 256, 257, 1000, 1024):
    st = ctx.stream(vp[0]); st.push(elems); st.finish()
    x = st.pull(first)
    y = st.pull(64)
    print(first, "first-window err", np.abs(x - want[:len(x)]).max(), "second-window err", np.abs(y - want[len(x):len(x)+len(y)]).max())
    if first in (8, 257):
        print("  got ", x[:8]); print("  want", want[:8])
        print("  got2 ", y[:8]); print("  want2", want[len(x):len(x)+8])
    st.close()
