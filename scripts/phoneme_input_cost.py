"""plan creation from Sequencer records against phoneme ids (Selector + Intonator on the device): 65 536 short
utterances of the default voice"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
v = g.voices.generic()
rng = np.random.default_rng(3)
n = 65536
lists = [[0] + [int(x) for x in rng.integers(3, 5, 2 + u % 3 - 1)] for u in range(n)]
elems, offs, vp = W.from_phonemes(lists, v, list(range(n)))
ids = np.concatenate([np.asarray(p, np.uint8) for p in lists])
cf = np.full(n, v.center_frequency, np.float32)
st = v.storage()
for rep in range(3):
    t0 = time.perf_counter(); p1 = ctx.plan(elems, offs, vp); ctx.synchronize(); t1 = time.perf_counter()
    p2 = ctx.plan_phonemes(ids, offs, st, vp, center_frequency=cf); ctx.synchronize(); t2 = time.perf_counter()
    print("records: %.2f ms (%.1f MB H2D)   phoneme ids: %.2f ms (%.2f MB H2D)   samples %d" % (
        (t1 - t0) * 1e3, elems.nbytes / 1e6, (t2 - t1) * 1e3, (ids.nbytes + cf.nbytes + st.nbytes) / 1e6, p1.total_samples))
    p1.close(); p2.close()
