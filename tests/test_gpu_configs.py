"""BASELINE.json configs 3, 4 and 5 on the GPU at (or near) full size: exact counts, bit-exact F_t / phase taps where
the oracle finishes in seconds, spot-checked audio parity, and size-independent properties."""
import os

import numpy as np
import pytest

import grail_rs_b200 as g
from grail_rs_b200 import workloads as W

pytestmark = pytest.mark.gpu
MAX_ABS, MIN_SNR_DB = 1e-4, 90.0     # north_star tolerance


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


_CONFIG3 = {}


def _config3_oracle(oracle):
    if not _CONFIG3:
        elems, offs, vp = W.config3(1200)
        want, tr, fin = oracle.synthesize(elems, vp[0], trace=True)
        _CONFIG3.update(elems=elems, offs=offs, vp=vp, want=want, tr=tr)
    return _CONFIG3


@pytest.mark.parametrize("phase_mode", [1, 3])
def test_config3_long_form(ctx, oracle, phase_mode):
    """one 10-minute utterance: 26 457 161 samples, one chain, 1 200 phonemes.  phase_mode 1 (the
    default): the chunk-parallel exact phase with CTA-wide scans over the utterance's ~13 000 chunks; phase_mode 3: the
    fixed-point phase scan.  Two unrelated exact algorithms, the same bits."""
    c3 = _config3_oracle(oracle)
    elems, offs, vp, want, tr = c3["elems"], c3["offs"], c3["vp"], c3["want"], c3["tr"]
    ctx.set_option("phase_mode", phase_mode)
    try:
        plan = ctx.plan(elems, offs, vp)
        assert plan.total_samples == 26457161
        plan.launch()
        out = plan.read_output()
        if phase_mode == 3:
            ps = plan.phase_scan_stats()
            print(plan.timings(), ps)
            assert ps["scans"] == 1 and ps["converged"] == 1 and ps["refused"] == 0      # parallel-in-time phase scan, no serial chain
        else:
            ps = plan.phase_stats()
            print(plan.timings(), ps)
            assert ps["chunks"] >= 2048 and ps["unproven_utterances"] == 0, ps             # proven chunk by chunk, no serial chain
        f, ph, saw = plan.read_intermediates()
        assert np.array_equal(f.view(np.uint32), tr["frequency"].view(np.uint32))          # bit-exact fundamental
        assert np.array_equal(ph.view(np.uint32), tr["carrier_phase"].view(np.uint32))     # bit-exact carrier phase
        st = W.parity_stats(out, want)
        print(st)
        assert st["max_abs"] <= MAX_ABS and st["snr_db"] >= MIN_SNR_DB, st
        plan.close()
    finally:
        ctx.set_option("phase_mode", 1)


def test_long_form_chunk_parallel_phase(ctx, oracle):
    """the chunk-parallel walk on one long utterance below the CTA-scan threshold (120 phonemes, 2.6 M samples, ~1 300
    chunks scanned by a single warp): carrier phase bit-exact, every utterance proven"""
    elems, offs, vp = W.from_phonemes([W.config3_phonemes(120)], g.voices.generic(), [0])
    ctx.set_option("phase_mode", 2)
    try:
        plan = ctx.plan(elems, offs, vp)
        plan.launch()
        ps = plan.phase_stats()
        assert ps["chunks"] > 500 and ps["unproven_utterances"] == 0, ps
        f, ph, saw = plan.read_intermediates()
        want, tr, _ = oracle.synthesize(elems, vp[0], trace=True)
        assert np.array_equal(ph.view(np.uint32), tr["carrier_phase"].view(np.uint32))
        plan.close()
    finally:
        ctx.set_option("phase_mode", 1)


def test_config4_random_voices_sharded_shape(ctx, oracle):
    """short utterances with per-utterance random voices (8 active formants, no time chunking needed);
    4 096 of the 65 536, the slice one of 16 ranks would own"""
    elems, offs, vp = W.config4(4096, first_utt=8192)
    plan = ctx.plan(elems, offs, vp)
    counts = g.count_samples(elems, offs, vp)
    assert plan.total_samples == int(counts.sum())
    plan.launch()
    out = plan.read_output()
    oo = plan.out_offsets
    print(plan.timings(), plan.total_samples)
    assert np.isfinite(out).all()
    # every one of the 4 096 utterances against the oracle (all host threads)
    want, woo, wcounts = oracle.synthesize_batch(elems, offs, vp, n_threads=os.cpu_count() or 1)
    assert np.array_equal(woo, oo) and np.array_equal(wcounts, counts)
    worst = W.parity_batch(out, want, oo)
    print(worst)
    assert worst["max_abs"] <= MAX_ABS and worst["snr_db"] >= MIN_SNR_DB, worst
    assert plan.phase_stats()["unproven_utterances"] == 0
    plan.close()


@pytest.mark.parametrize("rate,count", [(16000.0, 80003), (22050.0, 110244), (44100.0, 220476), (48000.0, 240016)])
def test_config5_sample_rate_sweep(ctx, oracle, rate, count):
    """config-2 shape at four sample rates (the Intonator 'lookahead' of the config name is a no-op in the reference)"""
    elems, offs, vp = W.config2(64, 10, sample_rate=rate)
    plan = ctx.plan(elems, offs, vp)
    assert plan.total_samples == 64 * count
    plan.launch()
    out = plan.read_output()
    oo = plan.out_offsets
    want, woo, _ = oracle.synthesize_batch(elems, offs, vp, n_threads=os.cpu_count() or 1)
    assert np.array_equal(woo, oo)
    worst = W.parity_batch(out, want, oo)                       # all 64 utterances
    assert worst["max_abs"] <= MAX_ABS and worst["snr_db"] >= MIN_SNR_DB, (rate, worst)
    f, ph, saw = plan.read_intermediates()                      # bit-exact taps on three of them
    for u in (0, 31, 63):
        _, tr, _ = oracle.synthesize(elems[offs[u]:offs[u + 1]], vp[u], trace=True)
        assert np.array_equal(f[oo[u]:oo[u + 1]].view(np.uint32), tr["frequency"].view(np.uint32)), (rate, u)
        assert np.array_equal(ph[oo[u]:oo[u + 1]].view(np.uint32), tr["carrier_phase"].view(np.uint32)), (rate, u)
    plan.close()


def test_parallel_phase_scan_matches_chain(ctx, oracle):
    """the exact parallel phase scan (forced on short inputs) against the oracle's f32 chain, bit for bit: voiced,
    silence-heavy (frequency 0.25: a wrap every 4 samples, exact ties on wrap steps) and random-pitch inputs"""
    ctx.set_option("phase_mode", 0)              # (the default is the chunk-parallel walk, which needs no scan)
    ctx.set_option("pscan_min_samples", 1)
    ctx.set_option("pscan_cost_model", 0)
    try:
        v = g.voices.generic()
        cases = [W.from_phonemes([[0, 4, 3, 3, 4, 3]], v, [3]), W.from_phonemes([[0, 0, 0, 3, 0, 0, 4, 0]], v, [1]),
                 W.config4(3, first_utt=500), W.config4(2, sample_rate=16000.0, first_utt=900)]
        for elems, offs, vp in cases:
            plan = ctx.plan(elems, offs, vp)
            plan.launch()
            out = plan.read_output()
            ps = plan.phase_scan_stats()
            assert ps["scans"] == len(offs) - 1 and ps["converged"] == ps["scans"], ps
            f, ph, saw = plan.read_intermediates()
            oo = plan.out_offsets
            for u in range(len(offs) - 1):
                want, tr, _ = oracle.synthesize(elems[offs[u]:offs[u + 1]], vp[u], trace=True)
                assert np.array_equal(ph[oo[u]:oo[u + 1]].view(np.uint32), tr["carrier_phase"].view(np.uint32)), (u, ps)
                st = W.parity_stats(out[oo[u]:oo[u + 1]], want)
                assert st["max_abs"] <= MAX_ABS and st["snr_db"] >= MIN_SNR_DB, st
            print(ps)
            plan.close()
    finally:
        ctx.set_option("pscan_min_samples", 1 << 18)
        ctx.set_option("pscan_cost_model", 1)
        ctx.set_option("phase_mode", 1)


def test_pipelined_launches_match_in_order_launches(ctx):
    """option "pipeline": consecutive launches of a plan overlap (the next launch's frequency / phase kernels run on a
    second stream and scratch set under the current formant kernel).  Same bits as launching in order, whether the
    launches share one output buffer or alternate between two."""
    import torch
    elems, offs, vp = W.config2(96, 4)
    plan = ctx.plan(elems, offs, vp)
    plan.launch()
    want = plan.read_output().copy()
    plan.close()
    ctx.set_option("pipeline", 1)
    try:
        plan = ctx.plan(elems, offs, vp)
        a = torch.zeros(plan.total_samples, dtype=torch.float32, device="cuda")
        b = torch.zeros_like(a)
        for i in range(5):
            plan.launch(a.data_ptr() if i % 2 == 0 else b.data_ptr())
        plan.join()
        ctx.synchronize()
        assert np.array_equal(a.cpu().numpy().view(np.uint32), want.view(np.uint32))
        assert np.array_equal(b.cpu().numpy().view(np.uint32), want.view(np.uint32))
        t = plan.timings()
        assert t["n_launches"] >= 3
        plan.close()
    finally:
        ctx.set_option("pipeline", 0)


def test_interleaved_planner_against_plain_order(ctx, oracle):
    """100 utterances of three lengths in random order: groups of 32 equally long ones are interleaved chunk by chunk
    (with empty padding items between groups), the rest stay consecutive.  Every utterance must come out where
    out_offsets says, equal to the plain-order plan at rounding level and to the oracle within the tolerance."""
    rng = np.random.default_rng(11)
    v = g.voices.generic()
    lists = [[0] + [int(x) for x in rng.integers(3, 5, int(n))] for n in rng.integers(1, 4, 100)]
    elems, offs, vp = W.from_phonemes(lists, v, list(range(100)))
    outs = {}
    for inter in (1, 0):
        ctx.set_option("interleave", inter)
        ctx.set_option("min_chunk", 4096)
        ctx.set_option("target_lanes", 100000)          # several chunks per utterance
        try:
            plan = ctx.plan(elems, offs, vp)
            plan.launch()
            outs[inter] = (plan.read_output().copy(), plan.out_offsets.copy())
            plan.close()
        finally:
            ctx.set_option("interleave", 1)
            ctx.set_option("min_chunk", 2048)
            ctx.set_option("target_lanes", 0)
    a, oo = outs[1]
    b, oo_b = outs[0]
    assert np.array_equal(oo, oo_b)
    assert float(np.abs(a - b).max()) < 2e-6
    for u in rng.choice(100, 10, replace=False):
        want, _, _ = oracle.synthesize(elems[offs[u]:offs[u + 1]], vp[u])
        st = W.parity_stats(a[oo[u]:oo[u + 1]], want)
        assert st["max_abs"] <= MAX_ABS and st["snr_db"] >= MIN_SNR_DB, (u, st)
