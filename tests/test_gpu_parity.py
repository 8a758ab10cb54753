"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): sample counts, LCG/jitter stream, Sequencer clock, F_t, carrier phase and saw
bit-exact; audio within max-abs 1e-4 and >= 90 dB SNR of the oracle's f32 output."""
import json
import os

import numpy as np
import pytest

import grail_rs_b200 as g
from grail_rs_b200 import workloads as W

pytestmark = pytest.mark.gpu

MAX_ABS = 1e-4      # north_star tolerance
MIN_SNR_DB = 90.0   # north_star tolerance
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_probe_kat.json")))


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


def check_batch(ctx, oracle, elems, offs, vp, exact_taps=True, label=""):
    """run the batch on the device and compare every utterance with the oracle; returns worst-case stats"""
    plan = ctx.plan(elems, offs, vp)
    plan.launch()
    out = plan.read_output()
    oo = plan.out_offsets
    taps = plan.read_intermediates() if exact_taps else None
    worst = {"max_abs": 0.0, "snr_db": float("inf")}
    for u in range(len(offs) - 1):
        e = elems[offs[u]:offs[u + 1]]
        want, tr, _ = oracle.synthesize(e, vp[u], trace=exact_taps)
        n = int(oo[u + 1] - oo[u])
        assert n == len(want), f"{label} utt {u}: sample count {n} != oracle {len(want)}"
        got = out[oo[u]:oo[u + 1]]
        if exact_taps and n:
            f, ph, saw = (t[oo[u]:oo[u + 1]] for t in taps)
            assert np.array_equal(f.view(np.uint32), tr["frequency"].view(np.uint32)), f"{label} utt {u}: F_t not bit-exact"
            assert np.array_equal(ph.view(np.uint32), tr["carrier_phase"].view(np.uint32)), f"{label} utt {u}: phase not bit-exact"
        if n and np.any(want):
            st = W.parity_stats(got, want)
            assert st["max_abs"] <= MAX_ABS and st["snr_db"] >= MIN_SNR_DB, f"{label} utt {u}: {st}"
            worst["max_abs"] = max(worst["max_abs"], st["max_abs"])
            worst["snr_db"] = min(worst["snr_db"], st["snr_db"])
        elif n:
            assert not np.any(got), f"{label} utt {u}: oracle is silent, device is not"
    plan.close()
    return worst, out, oo


@pytest.mark.parametrize("kat", [k for k in KAT["utterances"] if not k.get("slow")], ids=lambda k: k["name"])
def test_kat_utterances(ctx, oracle, kat):
    elems, offs, vp = W.from_phonemes([kat["phonemes"]], g.voices.generic(), [kat["jitter_seed"]])
    worst, out, oo = check_batch(ctx, oracle, elems, offs, vp, label=kat["name"])
    assert len(out) == kat["n"]
    print(kat["name"], worst)


@pytest.mark.parametrize("min_chunk,target", [(32, 1 << 20), (256, 1 << 16), (2048, 0), (1 << 22, 1)])
def test_chunking_is_transparent(ctx, oracle, min_chunk, target):
    """any time-chunking (from 32-sample chunks to one chunk per utterance) gives the same audio"""
    ctx.set_option("min_chunk", min_chunk)
    ctx.set_option("target_lanes", target)
    try:
        elems, offs, vp = W.from_phonemes([[0, 4, 3], [3, 0, 4, 4], [4]], g.voices.generic(), [1, 2, 3])
        worst, _, _ = check_batch(ctx, oracle, elems, offs, vp, label=f"chunk{min_chunk}")
        print(min_chunk, worst)
    finally:
        ctx.set_option("min_chunk", 2048)
        ctx.set_option("target_lanes", 0)


def test_edge_cases(ctx, oracle):
    v = g.voices.generic()
    lists = [[], [3], [0], [0, 1, 2], [0, 0, 0, 3], [3, 0, 0, 4], [4, 3]]
    elems, offs, vp = W.from_phonemes(lists, v, list(range(len(lists))))
    elems = elems.copy()
    # ragged lengths, a phoneme shorter than one sample, zero blend length
    elems["length"][offs[5]:offs[6]] = np.array([0.3, 1e-6, 0.011, 0.25], np.float32)
    elems["blend_length"][offs[6]] = 0.0
    elems["blend_length"][offs[4] + 3] = 0.05
    worst, out, oo = check_batch(ctx, oracle, elems, offs, vp, label="edge")
    assert oo[1] == 0                      # empty utterance yields nothing
    print(worst)


def test_empty_batch(ctx):
    out, oo = ctx.synthesize_batch(np.zeros(0, g.SEQ_ELEM_DT), np.zeros(1, np.uint32), np.zeros(0, g.VOICE_DT))
    assert len(out) == 0 and oo.tolist() == [0]


@pytest.mark.parametrize("rate", [16000.0, 22050.0, 48000.0])
def test_sample_rate_sweep(ctx, oracle, rate):
    """config 5 shape at small size: the voice rebuilt per rate; counts per SURVEY 8d"""
    elems, offs, vp = W.config2(3, 10, sample_rate=rate)
    worst, out, oo = check_batch(ctx, oracle, elems, offs, vp, label=f"rate{rate}")
    assert int(oo[1]) == KAT["sample_rate_counts"][str(int(rate))][2]
    print(rate, worst)


def test_random_voices(ctx, oracle):
    """config 4 distribution at small size: 8 active formants, random pitch / bandwidths / jitter"""
    for rate in (44100.0, 16000.0):
        elems, offs, vp = W.config4(24, sample_rate=rate)
        worst, _, _ = check_batch(ctx, oracle, elems, offs, vp, label=f"cfg4@{rate}")
        print(rate, worst)


def test_random_voices_chunked(ctx, oracle):
    """the same with forced short chunks: exercises the warm-up bound on high-Q random resonators"""
    ctx.set_option("min_chunk", 1024)
    ctx.set_option("target_lanes", 1 << 20)
    try:
        elems, offs, vp = W.config4(16, sample_rate=44100.0, first_utt=100)
        worst, _, _ = check_batch(ctx, oracle, elems, offs, vp, label="cfg4-chunked")
        print(worst)
    finally:
        ctx.set_option("min_chunk", 2048)
        ctx.set_option("target_lanes", 0)


def test_iterator_chain_drop_in(ctx, oracle):
    """the reference's own call shape (examples/cli.rs:175-184) through the mirrored verbs"""
    v = g.voices.generic()
    lang = g.languages.generic()
    audio = (g.transcribe("a pie i oui e a", lang).intonate(lang, v).select(v)
             .sequence(v).jitter(0, v).synthesize(ctx).collect())
    ph = [int(p) for p in g.transcribe("a pie i oui e a", lang)]
    want, _, _ = oracle.synthesize(oracle.select(ph, oracle.generic_voice()), oracle.voice_params(oracle.generic_voice(), 0))
    assert len(audio) == len(want)
    st = W.parity_stats(audio, want)
    assert st["max_abs"] <= MAX_ABS and st["snr_db"] >= MIN_SNR_DB, st
    it = g.sequence([g.SequenceElem.new(v.phonemes.a.copy_with_frequency(v.center_frequency), 0.01, 0.01)], v).jitter(3, v).synthesize(ctx)
    first = [next(it) for _ in range(5)]
    assert len(first) == 5 and len(first + list(it)) in (440, 441)


def test_i16_output(ctx, oracle):
    elems, offs, vp = W.from_phonemes([[0, 3, 4]], g.voices.generic(), [5])
    plan = ctx.plan(elems, offs, vp)
    plan.launch(fmt=g.I16)
    pcm = plan.read_output(fmt=g.I16)
    want, _, _ = oracle.synthesize(elems, vp[0])
    ref = np.trunc(want.astype(np.float32) * np.float32(32767.0)).astype(np.int16)   # examples/cli.rs:50
    assert len(pcm) == len(ref) and np.abs(pcm.astype(np.int32) - ref.astype(np.int32)).max() <= 4
    plan.close()
    # the one-shot i16 entry point gives the very same PCM (pageable destination, non-zero first offset)
    oo = np.array([3, 3 + len(ref)], np.uint64)
    buf = np.full(len(ref) + 5, -7, np.int16)
    got, _ = ctx.synthesize_batch(elems, offs, vp, out=buf, out_offsets=oo, fmt=g.I16)
    assert np.array_equal(got[3:3 + len(ref)], pcm) and (got[:3] == -7).all() and (got[-2:] == -7).all()


def test_count_mismatch_is_reported(ctx):
    elems, offs, vp = W.from_phonemes([[0, 3]], g.voices.generic(), [0])
    bad = np.array([0, 44100], np.uint64)
    with pytest.raises(g.GrailError) as ei:
        ctx.synthesize_batch(elems, offs, vp, out=np.zeros(44100, np.float32), out_offsets=bad)
    assert ei.value.status == g._ffi.ERR_COUNT_MISMATCH


def test_relaunch_is_deterministic(ctx):
    elems, offs, vp = W.config2(8, 4)
    plan = ctx.plan(elems, offs, vp)
    plan.launch()
    a = plan.read_output().copy()
    plan.launch()
    b = plan.read_output()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    t = plan.timings()
    assert t["n_launches"] >= 3 and t["total_ms"] > 0
    plan.close()


def test_properties_the_author_left_empty(ctx, oracle):
    """synthesize_normalized / jitter_within_bounds (src/lib.rs:603-608, 804-805) as real property tests"""
    elems, offs, vp = W.config2(16, 6)
    out, oo = ctx.synthesize_batch(elems, offs, vp)
    assert np.isfinite(out).all() and np.abs(out).max() <= 1.0        # peak values don't exceed 1.0
    plan = ctx.plan(elems, offs, vp)
    plan.launch()
    f, ph, saw = plan.read_intermediates()
    cf, dj = float(g.voices.generic().center_frequency), float(g.voices.generic().jitter_delta_frequency)
    voiced = f[f < 0.2]
    assert voiced.min() >= cf - dj * 1.0001 and voiced.max() <= cf + dj * 1.0001   # jitter stays inside its bounds
    assert ph.min() >= 0.0 and ph.max() < 1.0
    plan.close()


def test_full_size_config2_properties(ctx, oracle):
    """BASELINE config 2 at full size: exact counts, a spot-check of utterances against the oracle, and
    size-independent properties (same phonemes + same seed => identical audio; different seed => different)"""
    elems, offs, vp = W.config2(1024, 10)
    vp = vp.copy()
    ph = W.config2_phonemes(1024, 10)
    twin = next(u for u in range(1, 1024) if ph[u] == ph[0])
    vp["jitter_seed"][twin] = vp["jitter_seed"][0]
    plan = ctx.plan(elems, offs, vp)
    assert plan.total_samples == 1024 * 220476
    plan.launch()
    out = plan.read_output()
    oo = plan.out_offsets
    assert np.isfinite(out).all()
    # (bit-reproducible for a fixed batch; across lanes the exact/interpolated block choice is made per warp, so twins
    #  agree to rounding level, like two different chunkings of the same utterance)
    assert np.abs(out[oo[0]:oo[1]] - out[oo[twin]:oo[twin + 1]]).max() < 5e-6
    # ... and a different seed on the same phonemes gives different audio
    vp2 = vp[[0, twin]].copy()
    vp2["jitter_seed"][1] = 12345
    e2 = np.concatenate([elems[offs[0]:offs[1]], elems[offs[twin]:offs[twin + 1]]])
    o2, oo2 = ctx.synthesize_batch(e2, np.array([0, 10, 20], np.uint32), vp2)
    # (a different batch is chunked differently, so equality is at warm-up precision, not bit level)
    assert np.abs(o2[:oo2[1]] - out[oo[0]:oo[1]]).max() < 5e-6
    assert np.abs(o2[oo2[1]:] - o2[:oo2[1]]).max() > 1e-3
    # all 1 024 utterances against the oracle (multi-threaded, about a second per host core-dozen)
    import os
    want, woo, _ = oracle.synthesize_batch(elems, offs, vp, n_threads=os.cpu_count() or 1)
    assert np.array_equal(woo, oo)
    worst = W.parity_batch(out, want, oo)
    print(worst)
    assert worst["max_abs"] <= MAX_ABS and worst["snr_db"] >= MIN_SNR_DB, worst
    assert plan.phase_stats()["unproven_utterances"] == 0
    print(plan.timings(), plan.phase_stats())
    plan.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_randomized_batches(ctx, oracle, seed):
    """random phoneme lists (incl. Stop/Glide), ragged random lengths and blend lengths, mixed sample rates and
    random-voice elements in one batch, random jitter seeds and chunkings"""
    rng = np.random.default_rng(seed)
    v441 = g.voices.generic()
    parts, offs, vps = [], [0], []
    for u in range(10):
        rate = float(rng.choice([16000.0, 22050.0, 44100.0, 48000.0]))
        if rng.random() < 0.4:
            e, o, vp = W.config4(1, sample_rate=rate, first_utt=int(rng.integers(0, 60000)))
            e = e.copy()
        else:
            voice = v441 if rate == 44100.0 else g.voices.at_sample_rate(v441, rate)
            n = int(rng.integers(1, 6))
            e, o, vp = W.from_phonemes([[int(x) for x in rng.integers(0, 5, n)]], voice, [int(rng.integers(0, 2**32))])
            e = e.copy()
        e["length"] = rng.uniform(0.03, 0.6, len(e)).astype(np.float32)
        e["blend_length"] = np.where(rng.random(len(e)) < 0.5, e["length"], rng.uniform(0.01, 0.7, len(e))).astype(np.float32)
        parts.append(e)
        offs.append(offs[-1] + len(e))
        vps.append(vp[0])
    elems = np.concatenate(parts)
    ctx.set_option("min_chunk", int(rng.choice([256, 1024, 2048, 8192])))
    ctx.set_option("target_lanes", int(rng.choice([0, 1 << 20, 64])))
    try:
        worst, _, _ = check_batch(ctx, oracle, elems, np.array(offs, np.uint32), np.array(vps), label=f"rand{seed}")
        print(worst)
    finally:
        ctx.set_option("min_chunk", 2048)
        ctx.set_option("target_lanes", 0)


def test_unaligned_device_output(ctx, oracle):
    """a caller-provided device buffer that is only 4-byte aligned (the kernel's 128-bit stores must fall back)"""
    import torch
    elems, offs, vp = W.from_phonemes([[0, 3], [4]], g.voices.generic(), [1, 2])
    plan = ctx.plan(elems, offs, vp)
    buf = torch.zeros(plan.total_samples + 8, dtype=torch.float32, device="cuda")
    for shift in (1, 2, 3):
        plan.launch(buf.data_ptr() + 4 * shift)
        ctx.synchronize()
        got = buf[shift:shift + plan.total_samples].cpu().numpy()
        oo = plan.out_offsets
        for u in range(2):
            want, _, _ = oracle.synthesize(elems[offs[u]:offs[u + 1]], vp[u])
            st = W.parity_stats(got[oo[u]:oo[u + 1]], want)
            assert st["max_abs"] <= MAX_ABS and st["snr_db"] >= MIN_SNR_DB, (shift, u, st)
    plan.close()


def test_interleaved_channels_and_wav(ctx, oracle, tmp_path):
    """f2 of SURVEY 8f: channel duplication (examples/cli.rs:229) and 16-bit PCM on the device, RIFF header on the host"""
    import torch
    elems, offs, vp = W.from_phonemes([[0, 3, 4], [4]], g.voices.generic(), [5, 6])
    plan = ctx.plan(elems, offs, vp)
    plan.launch()
    mono = plan.read_output().copy()
    for ch in (2, 3):
        buf = torch.zeros(plan.total_samples * ch, dtype=torch.float32, device="cuda")
        plan.launch_interleaved(buf.data_ptr(), ch)
        ctx.synchronize()
        got = buf.cpu().numpy().reshape(-1, ch)
        assert np.array_equal(got[:, 0].view(np.uint32), mono.view(np.uint32))
        assert all(np.array_equal(got[:, 0], got[:, c]) for c in range(1, ch))
    pcm = torch.zeros(plan.total_samples * 2, dtype=torch.int16, device="cuda")
    plan.launch_interleaved(pcm.data_ptr(), 2, fmt=g.I16)
    ctx.synchronize()
    pcm = pcm.cpu().numpy()
    ref = np.trunc(mono * np.float32(32767.0)).astype(np.int16)
    assert np.array_equal(pcm[0::2], ref) and np.array_equal(pcm[1::2], ref)
    path = str(tmp_path / "out.wav")
    g.save_wav(path, pcm, 44100, channels=2)
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and len(raw) == 44 + 2 * len(pcm)
    plan.close()


def test_synthesize_resampled_property(ctx, oracle):
    """the author's empty `synthesize_resampled` (src/lib.rs:606-608): resampling gives a similar output.
    Same phonemes at 44.1 kHz and 22.05 kHz: same duration, similar level, similar spectral envelope."""
    ph = [[0, 3, 4, 3]]
    v = g.voices.generic()
    outs = {}
    for rate in (44100.0, 22050.0):
        voice = v if rate == 44100.0 else g.voices.at_sample_rate(v, rate)
        e, o, p = W.from_phonemes(ph, voice, [3])
        outs[rate], _ = ctx.synthesize_batch(e, o, p)
    a, b = outs[44100.0], outs[22050.0]
    assert abs(len(a) / 44100.0 - len(b) / 22050.0) < 1e-3
    ra, rb = np.sqrt(np.mean(a.astype(np.float64) ** 2)), np.sqrt(np.mean(b.astype(np.float64) ** 2))
    assert 0.5 < ra / rb < 2.0
    # band energies below 5 kHz on a common frequency grid
    def bands(x, rate):
        spec = np.abs(np.fft.rfft(x.astype(np.float64) * np.hanning(len(x)))) ** 2
        freqs = np.fft.rfftfreq(len(x), 1.0 / rate)
        edges = np.linspace(100, 5000, 15)
        e = np.array([spec[(freqs >= lo) & (freqs < hi)].sum() for lo, hi in zip(edges[:-1], edges[1:])])
        return np.log10(e / e.sum())
    assert np.abs(bands(a, 44100.0) - bands(b, 22050.0)).max() < 1.0      # within a decade in every band


def test_division_by_segment_constant_is_ieee_exact(ctx):
    """k_frequency divides the Sequencer clock by the segment's blend length through RN(1/b) and one FMA correction
    (Markstein): 2^31 pseudo-random pairs across 60 binades each must equal the IEEE quotient bit for bit"""
    import ctypes as C
    bad = C.c_uint64(123)
    rc = ctx._L.grail_cuda_debug_div_check(ctx._h, 20261017, 1 << 31, C.byref(bad))
    assert rc == 0 and bad.value == 0, (rc, bad.value)


def test_jitter_frequency_extremes(ctx, oracle):
    """value-noise increments from 0 to the largest the path accepts (0.25: a wrap every 4 samples), around the 1/17 and
    1/9 thresholds where k_formant's blocks stop being split / interpolated: counts, F_t and phase bit-exact, audio in
    tolerance -- in ONE batch, so that the 32 lanes of a warp disagree about which path a block takes"""
    incs = [0.0, 1e-6, 1.0 / 2756.0, 0.01, 1.0 / 17.0, 0.0588235, 0.06, 0.1, 1.0 / 9.0, 0.12, 0.2, 0.25]
    lists = [[0, 3, 4][: 1 + (k % 3)] + [4, 3] for k in range(len(incs))]
    elems, offs, vp = W.from_phonemes(lists, g.voices.generic(), list(range(40, 40 + len(incs))))
    elems = elems.copy()
    elems["length"] = np.float32(0.05)
    elems["blend_length"] = np.float32(0.03)
    vp = vp.copy()
    vp["jitter_frequency"] = np.asarray(incs, np.float32)
    worst, out, oo = check_batch(ctx, oracle, elems, offs, vp, label="jitter extremes")
    print(worst)
