"""small end-to-end run for compute-sanitizer (memcheck / racecheck): batch with chunking, long-form scan, stream"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
ctx.set_option("min_chunk", 1024); ctx.set_option("target_lanes", 1 << 20)
elems, offs, vp = W.from_phonemes([[0, 4, 3], [3], [0, 0, 4, 3]], g.voices.generic(), [1, 2, 3])
elems = elems.copy(); elems["length"] = 0.05
out, oo = ctx.synthesize_batch(elems, offs, vp)
print("batch", len(out), float(np.abs(out).sum()))
ctx.set_option("pscan_min_samples", 1)
ctx.set_option("pscan_cost_model", 0)
out2, _ = ctx.synthesize_batch(elems, offs, vp)
print("pscan", float(np.abs(out2 - out).max()))
e4, o4, v4 = W.config4(3, first_utt=7)
e4 = e4.copy(); e4["length"] = 0.04
out4, _ = ctx.synthesize_batch(e4, o4, v4)
print("cfg4", len(out4))
st = ctx.stream(vp[0]); st.push(elems[:3]); st.finish()
n = 0
while True:
    x = st.pull(777)
    if len(x) == 0: break
    n += len(x)
print("stream", n)
# phoneme-level input (k_select) and an interleaved group of 32 equally long utterances
v = g.voices.generic()
lists = [[0, 3, 4]] * 33 + [[3]]
ids = np.concatenate([np.asarray(p, np.uint8) for p in lists])
poffs = np.concatenate([[0], np.cumsum([len(p) for p in lists])]).astype(np.uint32)
pvp = np.zeros(len(lists), g.VOICE_DT); pvp[:] = v.params(0); pvp["jitter_seed"] = np.arange(len(lists))
ctx.set_option("pscan_min_samples", 1 << 18); ctx.set_option("pscan_cost_model", 1)
ctx.set_option("min_chunk", 2048); ctx.set_option("target_lanes", 1 << 20)
plan = ctx.plan_phonemes(ids, poffs, v.storage(), pvp, center_frequency=np.full(len(lists), v.center_frequency, np.float32))
plan.launch(); o = plan.read_output(); print("phonemes", len(o), float(np.abs(o).sum())); plan.close()
# a long utterance whose chunk records are scanned by a CTA of 32 warps (>= 2048 phase chunks of 256 samples)
ctx.set_option("min_chunk", 2048); ctx.set_option("target_lanes", 0); ctx.set_option("phase_chunk", 256)
el, ol, vl = W.from_phonemes([W.config3_phonemes(28)], g.voices.generic(), [5])
plan = ctx.plan(el, ol, vl)
plan.launch(); o = plan.read_output(); print("long form", len(o), plan.phase_stats()); plan.close()
ctx.set_option("phase_chunk", 0)
