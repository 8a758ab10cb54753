"""pull latency of a grail_cuda_stream (the interactive.rs use: an audio callback pulling fixed windows)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
ctx = g.Context(0)
v = g.voices.generic()
elems, offs, vp = W.from_phonemes([[0, 3, 4, 3, 4, 3, 3, 4, 3, 4] * 4], v, [1])
for window in (441, 4410, 44100):
    st = ctx.stream(vp[0])
    st.push(elems)
    st.finish()
    lat = []
    total = 0
    while True:
        t0 = time.perf_counter()
        x = st.pull(window)
        lat.append(time.perf_counter() - t0)
        if len(x) == 0:
            break
        total += len(x)
    lat = np.array(lat[1:-1]) * 1e6
    print(f"window {window:6d} samples ({window/44.1:.0f} ms of audio): {len(lat)} pulls, median {np.median(lat):.0f} us, p99 {np.percentile(lat, 99):.0f} us, "
          f"real-time factor {window/44100/np.median(lat)*1e6:.0f}")
    st.close() if hasattr(st, "close") else None

# batched pulls (grail_cuda_streams_pull): N concurrent streams, one plan and one launch per kernel per tick
print("batched pulls of a 10 ms window (441 samples per stream and tick):")
for n in (1, 8, 64, 256, 1024):
    streams = []
    for k in range(n):
        p = vp[0].copy()
        p["jitter_seed"] = k
        st = ctx.stream(p)
        st.push(elems)
        st.finish()
        streams.append(st)
    lat = []
    for tick in range(60):
        t0 = time.perf_counter()
        xs = g.pull_streams(streams, 441)
        lat.append(time.perf_counter() - t0)
        if len(xs[0]) == 0:
            break
    lat = np.array(lat[2:]) * 1e6
    print(f"  {n:5d} streams: median {np.median(lat):8.0f} us per tick = {np.median(lat) / n:7.1f} us per stream, "
          f"{n * 441 / np.median(lat) * 1e6:.3e} samples/s aggregate, {n * 0.01 / np.median(lat) * 1e6:.0f} real-time streams per GPU")
    for st in streams:
        st.close()
