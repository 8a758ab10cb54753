"""Synthetic phoneme-sequence workloads of BASELINE.json `configs`.  Host-side input
builders only: they produce grail_seq_elem / grail_voice_params records for the C ABI (and, in tests and
bench.py, the very same records are handed to the oracle)."""
from __future__ import annotations

import numpy as np

from ._ffi import SEQ_ELEM_DT, VOICE_DT
from .synth import SequenceElem, Voice, f32, pack_sequence
from . import voices as _voices

SILENCE, STOP, GLIDE, A, E = range(5)


def _xorshift32(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= (x << 13) & 0xFFFFFFFF
    x ^= x >> 17
    x ^= (x << 5) & 0xFFFFFFFF
    return x & 0xFFFFFFFF


def phoneme_records(voice: Voice, length=0.5, blend_length=0.5) -> dict:
    """one packed SequenceElem per phoneme id, as Intonator + Selector would emit it (src/lib.rs:1068-1073, 990-1005)"""
    recs = {}
    for pid in range(5):
        elem = None
        if pid == A:
            elem = voice.phonemes.a.copy_with_frequency(voice.center_frequency)
        elif pid == E:
            elem = voice.phonemes.e.copy_with_frequency(voice.center_frequency)
        recs[pid] = pack_sequence([SequenceElem.new(elem, length, blend_length)])[0]
    return recs


def from_phonemes(phoneme_lists, voice: Voice, jitter_seeds=None):
    """utterances given as lists of phoneme ids, all with `voice`"""
    recs = phoneme_records(voice)
    table = np.stack([recs[i] for i in range(5)])
    flat = np.concatenate([np.asarray(p, np.int64) for p in phoneme_lists]) if len(phoneme_lists) else np.zeros(0, np.int64)
    elems = table[flat] if len(flat) else np.zeros(0, SEQ_ELEM_DT)
    offs = np.concatenate([[0], np.cumsum([len(p) for p in phoneme_lists])]).astype(np.uint32)
    n = len(phoneme_lists)
    if jitter_seeds is None:
        jitter_seeds = np.arange(n)
    vp = np.zeros(n, VOICE_DT)
    vp[:] = voice.params(0)
    vp["jitter_seed"] = np.asarray(jitter_seeds, np.uint64).astype(np.uint32)
    return np.ascontiguousarray(elems), offs, vp


def config2_phonemes(n_utts=1024, n_phonemes=10):
    out = []
    for utt in range(n_utts):
        x = _xorshift32(0x9E3779B9 ^ utt) or 1
        ph = [SILENCE]
        for _ in range(n_phonemes - 1):
            x = _xorshift32(x)
            ph.append(A if (x >> 7) & 1 else E)
        out.append(ph)
    return out


def config2(n_utts=1024, n_phonemes=10, sample_rate=44100.0):
    """batch of ~5 s utterances [Silence, p1..p9], p_j in {A, E}, default voice, jitter_seed = utt"""
    v = _voices.generic()
    if float(sample_rate) != 44100.0:
        v = _voices.at_sample_rate(v, sample_rate)
    return from_phonemes(config2_phonemes(n_utts, n_phonemes), v)


def config3_phonemes(n_phonemes=1200):
    return [SILENCE] + [3 + ((i * 7 + i // 3) & 1) for i in range(1, n_phonemes)]


def config3(n_phonemes=1200):
    """one long-form utterance (26 457 161 samples at 1 200 phonemes)"""
    return from_phonemes([config3_phonemes(n_phonemes)], _voices.generic(), [0])


def config4(n_utts=65536, sample_rate=44100.0, first_utt=0):
    """short utterances, (2 + utt mod 3) phonemes, per-utterance random voice (pitch, 8 formants, jitter)"""
    R = f32(sample_rate)
    n_ph = 2 + (np.arange(first_utt, first_utt + n_utts) % 3)
    offs = np.concatenate([[0], np.cumsum(n_ph)]).astype(np.uint32)
    elems = np.zeros(int(offs[-1]), SEQ_ELEM_DT)
    vp = np.zeros(n_utts, VOICE_DT)
    for k in range(n_utts):
        utt = first_utt + k
        rng = np.random.default_rng(_xorshift32(utt + 1))
        pitch = f32(rng.uniform(80.0, 300.0)) / R
        nyq = 0.45 * float(R)
        edges = np.linspace(200.0, min(nyq, 6000.0), 9)
        vp[k] = (R, f32(rng.uniform(8.0, 32.0)) / R, f32(rng.uniform(0.0, 12.0)) / R, f32(rng.uniform(0.0, 12.0)) / R,
                 f32(rng.uniform(0.0, 0.4)), utt & 0xFFFFFFFF, 0)
        for j in range(int(n_ph[k])):
            rec = elems[offs[k] + j]
            rec["length"] = f32(0.5)
            rec["blend_length"] = f32(0.5)
            if j == 0 and (utt & 1):
                continue  # some utterances open with a silence
            rec["has_elem"] = 1
            amp = rng.uniform(0.0, 1.0, 8).astype(np.float32)
            s = f32(0.0)
            for a in amp:
                s = f32(s + a)
            rec["elem"]["frequency"] = min(pitch, f32(0.5))
            rec["elem"]["formant_freq"] = (rng.uniform(edges[:-1], edges[1:]).astype(np.float32) / R).astype(np.float32)
            rec["elem"]["formant_bw"] = (rng.uniform(40.0, 250.0, 8).astype(np.float32) / R).astype(np.float32)
            rec["elem"]["formant_smooth"] = (rng.uniform(800.0, 3000.0, 8).astype(np.float32) / R).astype(np.float32)
            rec["elem"]["formant_breath"] = rng.uniform(0.0, 1.0, 8).astype(np.float32)
            rec["elem"]["formant_turb"] = rng.uniform(0.0, 1.0, 8).astype(np.float32)
            rec["elem"]["formant_amp"] = (amp / s).astype(np.float32)
    return elems, offs, vp


def parity_stats(got: np.ndarray, want: np.ndarray) -> dict:
    """max-abs error and SNR (dB) of `got` against the reference waveform `want`"""
    g64, w64 = got.astype(np.float64), want.astype(np.float64)
    err = g64 - w64
    num = float(np.sum(w64 * w64))
    den = float(np.sum(err * err))
    snr = float("inf") if den == 0.0 else (-float("inf") if num == 0.0 else 10.0 * np.log10(num / den))
    return {"max_abs": float(np.max(np.abs(err))) if len(err) else 0.0, "snr_db": snr}


def parity_batch(got: np.ndarray, want: np.ndarray, out_offsets: np.ndarray) -> dict:
    """worst per-utterance max-abs error and SNR of a packed batch against the reference waveforms (same packing)"""
    worst = {"max_abs": 0.0, "snr_db": float("inf"), "worst_utt": -1, "n_utts": len(out_offsets) - 1}
    for u in range(len(out_offsets) - 1):
        a, b = int(out_offsets[u]), int(out_offsets[u + 1])
        if a == b:
            continue
        st = parity_stats(got[a:b], want[a:b])
        if st["max_abs"] > worst["max_abs"] or st["snr_db"] < worst["snr_db"]:
            worst["worst_utt"] = u
        worst["max_abs"] = max(worst["max_abs"], st["max_abs"])
        worst["snr_db"] = min(worst["snr_db"], st["snr_db"])
    return worst
