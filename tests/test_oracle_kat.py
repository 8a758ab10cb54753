"""Pins the CPU oracle against the survey probe's known answers.

The reference has no golden vectors for this path (parity unpinned); two independent strict-f32
restatements (the survey probe and oracle/grail_oracle.c) agreeing bit-for-bit is the anchor.
"""
import json
import os

import numpy as np
import pytest

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_probe_kat.json")))


def bits(x):
    return "%08x" % int(np.array(x, np.float32).view(np.uint32))


def test_lcg_stream(oracle):
    vals, states = oracle.random_f32_seq(0, 4)
    assert ["%08x" % s for s in states] == KAT["lcg_states_from_0"]
    np.testing.assert_allclose(vals, KAT["lcg_floats_from_0"], rtol=0, atol=5e-9)
    # the float map is exact: ((s>>9)|0x3F800000) - 1.5) * 2 has no rounding (src/lib.rs:50-54)
    for v, s in zip(vals, states):
        f = np.array((int(s) >> 9) | 0x3F800000, np.uint32).view(np.float32)
        assert float(v) == (float(f) - 1.5) * 2.0


def test_jitter_construction(oracle):
    st, cn = oracle.jitter_init_states(0)
    assert ["%08x" % s for s in st] == KAT["jitter_init_states_seed0"]
    np.testing.assert_allclose(cn, KAT["jitter_freq_noise_cur_next_seed0"], atol=5e-9)
    # Appendix C: the three private states are the shared stream after 2, 18 and 34 draws
    s = oracle.lcg_states(0, 34)
    assert [int(s[1]), int(s[17]), int(s[33])] == [int(x) for x in st]


def test_voice_table_bits(oracle, voice):
    a, e = voice["phonemes"]["a"], voice["phonemes"]["e"]
    vb = KAT["voice_bits"]
    assert bits(a["formant_freq"][0]) == vb["a.formant_freq[0]"]
    assert bits(a["formant_freq"][1]) == vb["a.formant_freq[1]"]
    assert bits(a["formant_bw"][0]) == vb["a.formant_bw[0]"]
    assert bits(a["formant_smooth"][0]) == vb["a.formant_smooth[0]"]
    assert [bits(x) for x in a["formant_amp"][:4]] == vb["a.formant_amp[0..3]"]
    assert [bits(x) for x in e["formant_amp"][:4]] == vb["e.formant_amp[0..3]"]
    assert bits(voice["center_frequency"]) == vb["center_frequency"]
    assert bits(voice["jitter_frequency"]) == vb["jitter_frequency"]
    assert bits(voice["jitter_delta_frequency"]) == vb["jitter_delta_frequency"]
    assert bits(np.float32(1.0) / np.float32(44100.0)) == vb["dt"]


def test_math_kernels(oracle, voice):
    a = voice["phonemes"]["a"]
    mb = KAT["math_bits"]
    assert bits(oracle.tan_approx(a["formant_freq"][0])) == mb["tan_approx(a.formant_freq[0])"]
    assert bits(oracle.exp_approx(a["formant_smooth"][0])) == mb["exp_approx(a.formant_smooth[0])"]


@pytest.mark.parametrize("kat", [k for k in KAT["utterances"] if not k.get("slow")], ids=lambda k: k["name"])
def test_utterance_kat(oracle, voice, kat):
    elems = oracle.select(kat["phonemes"], voice)
    out, tr, _ = oracle.synthesize(elems, oracle.voice_params(voice, kat["jitter_seed"]), trace=True)
    assert len(out) == kat["n"]
    assert "%08x" % oracle.fnv(out) == kat["fnv"]
    for idx, b in kat.get("samples", {}).items():
        assert bits(out[int(idx)]) == b, idx
    if "boundaries" in kat:
        pi = tr["phoneme_index"]
        assert [0] + list(np.flatnonzero(np.diff(pi)) + 1) == kat["boundaries"]
    if "jitter_wraps" in kat:
        assert int((np.diff(tr["jitter_phase"]) < 0).sum()) == kat["jitter_wraps"]
    if "carrier_wraps" in kat:
        assert int((np.diff(tr["carrier_phase"]) < 0).sum()) == kat["carrier_wraps"]
    if "peak" in kat:
        assert abs(float(np.abs(out).max()) - kat["peak"]) < 1e-3
    if "pitch_range" in kat:
        np.testing.assert_allclose([tr["frequency"].min(), tr["frequency"].max()], kat["pitch_range"], rtol=1e-5)


@pytest.mark.parametrize("rate", ["16000", "22050", "44100", "48000"])
def test_sample_rate_counts(oracle, rate):
    lo, hi, total = KAT["sample_rate_counts"][rate]
    v = oracle.generic_voice(float(rate))
    ph = [0, 4, 3, 3, 4, 3, 3, 4, 3, 3]
    elems = oracle.select(ph, v)
    assert oracle.count_samples(elems, float(rate)) == total
    per = [oracle.count_samples(elems[:i + 1], float(rate)) for i in range(len(ph))]
    lens = set(np.diff([0] + per).tolist())
    assert lens <= {lo, hi}


def test_o3_build_is_bit_identical(oracle, voice):
    elems = oracle.select([0, 4, 3, 3, 4], voice)
    vp = oracle.voice_params(voice, 7)
    a, _, fa = oracle.synthesize(elems, vp)
    b, _, fb = oracle.synthesize(elems, vp, o3=True)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(fa, fb)


def test_edge_cases(oracle, voice):
    vp = oracle.voice_params(voice, 0)
    # empty input: the Sequencer yields nothing (src/lib.rs:876-886, 930)
    assert oracle.count_samples(oracle.select([], voice), 44100.0) == 0
    # a single phoneme is played and faded to silence
    out, _, _ = oracle.synthesize(oracle.select([3], voice), vp)
    assert len(out) in (22047, 22048) and np.abs(out).max() > 0.01
    # Glide/Stop behave like Silence (src/lib.rs:666): all-silent input gives exact zeros
    out, _, _ = oracle.synthesize(oracle.select([0, 1, 2], voice), vp)
    assert len(out) > 60000 and not out.any()


@pytest.mark.slow
def test_long_form_kat(oracle, voice):
    kat = [k for k in KAT["utterances"] if k["name"] == "long1200"][0]
    ph = [0] + [3 + ((i * 7 + i // 3) & 1) for i in range(1, kat["n_phonemes"])]
    out, _, fin = oracle.synthesize(oracle.select(ph, voice), oracle.voice_params(voice, 0))
    assert len(out) == kat["n"]
    assert "%08x" % oracle.fnv(out) == kat["fnv"]
    assert "%08x" % int(fin[27]) == KAT["lcg_state_after_26457161"]
