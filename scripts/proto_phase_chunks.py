"""CPU prototype of the chunk-parallel exact carrier phase (k_phase_a / k_phase_scan / k_phase_b).

One "lane" per chunk of PC samples.  Round A walks every chunk from a GUESSED start phase (double prefix sum of
F_t, which ignores the f32 rounding of the accumulator and is therefore off by tens of 2^-23 units) to learn how
far the guess is from the truth; after the first carrier wrap inside the chunk every phase is a multiple of 2^-23
plus exact low-order additions, so the guessed trajectory and the true one differ by a constant k * 2^-23 from that
wrap on (translation invariance), and k is found by an integer prefix sum over the chunks of one utterance.
Round B walks every chunk again from the corrected start, which yields the outputs AND the proof: the end of
chunk c must equal the start of chunk c+1 bit for bit; by induction from the exact phase at sample 0 the whole
trajectory is then the reference's.  A mismatch (a "fluke": the translated trajectory crossed a binade boundary
or wrapped one step earlier or later than the true one) is repaired by shifting the downstream starts and walking
those chunks again.

This script measures how often that happens on the oracle's F_t (numpy float32 = strict IEEE f32).
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W

f32 = np.float32
ONE = f32(1.0)
U23 = f32(2.0 ** -23)


def walk(F, n, n0, p0, n1, out=None):
    """vectorised over chunks: literal f32 steps from sample n0[c] (phase p0[c]) to n1[c]; returns the phase at n1,
    index/value of the first wrap (index = sample whose step wrapped, value after it; -1 if none)"""
    C = len(n0)
    p = p0.astype(f32).copy()
    first_i = np.full(C, -1, np.int64)
    first_v = np.zeros(C, f32)
    t = n0.astype(np.int64).copy()
    steps = int((n1 - n0).max()) if C else 0
    for _ in range(steps):
        act = t < n1
        idx = np.minimum(t, n - 1)
        q = (p + F[idx]).astype(f32)
        wr = q >= ONE
        q = np.where(wr, (q - ONE).astype(f32), q)
        newly = act & wr & (first_i < 0)
        first_i = np.where(newly, t, first_i)
        first_v = np.where(newly, q, first_v)
        if out is not None:
            out[idx[act]] = p[act]
        p = np.where(act, q, p)
        t = t + act
    return p, first_i, first_v


def run_utt(F, phase_true, PC=2048, p_init=0.0, verbose=False):
    n = len(F)
    C = (n + PC - 1) // PC
    n0 = np.arange(C, dtype=np.int64) * PC
    n1 = np.minimum(n0 + PC, n)
    # guesses: double prefix sums (exact: every F is a multiple of 2^-40 or so)
    cs = np.concatenate([[0.0], np.cumsum(F.astype(np.float64))])
    G = ((p_init + cs[n0]) % 1.0).astype(f32)
    G[0] = f32(p_init)
    # ---- round A: [n0 -> n1] from the guess, then on to the first wrap beyond n1 (inside the next chunk)
    E, a, R = walk(F, n, n0, G, n1)
    nn1 = np.minimum(n1 + PC, n)
    _, a2, R2 = walk(F, n, n1, E, nn1)
    # ---- scan
    k = np.zeros(C, np.int64)
    s = G.copy()
    bad_anchor = 0
    for c in range(C - 1):
        if a2[c] >= 0 and a2[c] == a[c + 1]:
            d = (np.float64(R2[c]) - np.float64(R[c + 1])) * 2.0 ** 23
            assert d == np.round(d), d
            k[c + 1] = k[c] + int(d)
        else:
            bad_anchor += 1
            k[c + 1] = k[c]
        s[c + 1] = f32(np.float64(E[c]) + k[c] * 2.0 ** -23)
    # chunk 0 has no anchor issue: its start is exact, so "k" relative to its own trajectory is 0 -- but the
    # formula above used k[0] = 0 together with R[0] from the exact trajectory: consistent.
    # ---- round B + repairs
    rounds = 0
    dirty = np.ones(C, bool)
    e = np.zeros(C, f32)
    ph = np.zeros(n, f32)
    evals = 0
    while True:
        rounds += 1
        idx = np.flatnonzero(dirty)
        evals += len(idx)
        ee, _, _ = walk(F, n, n0[idx], s[idx], n1[idx], out=ph)
        e[idx] = ee
        phi = (e[:-1].astype(np.float64) - s[1:].astype(np.float64))
        if not np.any(phi != 0.0):
            break
        # translation by a multiple of 2^-23 carries through a chunk; anything else does not: that chunk is walked
        # again from its new start and the starts after it are left alone (they are re-checked next round)
        phi = (phi + 0.5) % 1.0 - 0.5
        ns = s.copy()
        delta = 0.0
        for c in range(C - 1):
            lat = (delta * 2.0 ** 23) == np.round(delta * 2.0 ** 23)
            delta = (phi[c] + delta) if lat else 0.0
            ns[c + 1] = f32((np.float64(s[c + 1]) + delta) % 1.0)
            delta = np.float64(ns[c + 1]) - np.float64(s[c + 1])
            delta = (delta + 0.5) % 1.0 - 0.5
        dirty = ns != s
        s = ns
        if rounds > 50:
            raise RuntimeError("no convergence")
    ok = np.array_equal(ph.view(np.uint32), phase_true.view(np.uint32))
    if verbose:
        print(f"  n={n} chunks={C} kmax={np.abs(k).max()} bad_anchor={bad_anchor} rounds={rounds} "
              f"evals={evals} ({evals / C:.2f}x) exact={ok}")
    return rounds, evals / C, ok, int(np.abs(k).max())


def main():
    v = g.voices.generic()
    stats = []
    # config 2
    elems, offs, vp = W.config2(24, 10, 44100.0)
    for u in range(24):
        _, tr, _ = O.synthesize(elems[offs[u]:offs[u + 1]], vp[u], trace=True)
        stats.append(run_utt(tr["frequency"], tr["carrier_phase"], 2048, verbose=(u < 4)))
    print("config2: rounds", [s[0] for s in stats], "all exact:", all(s[2] for s in stats),
          "mean evals/chunk %.3f" % np.mean([s[1] for s in stats]))
    # config 4 voices at several rates
    for rate in (16000.0, 22050.0, 44100.0, 48000.0):
        st = []
        elems, offs, vp = W.config4(24, rate)
        for u in range(24):
            _, tr, _ = O.synthesize(elems[offs[u]:offs[u + 1]], vp[u], trace=True)
            st.append(run_utt(tr["frequency"], tr["carrier_phase"], 2048, verbose=(u < 2)))
        print(f"config4 @{rate}: rounds", [s[0] for s in st], "all exact:", all(s[2] for s in st),
              "mean evals/chunk %.3f" % np.mean([s[1] for s in st]))
    # silence-heavy
    elems, offs, vp = W.from_phonemes([[0, 0, 0, 3, 0, 0, 4, 0]], v, [5])
    _, tr, _ = O.synthesize(elems, vp[0], trace=True)
    print("silence-heavy:", run_utt(tr["frequency"], tr["carrier_phase"], 2048, verbose=True))


if __name__ == "__main__":
    main()
