"""Second, independent restatement of the reference's waveform path, in NumPy float32.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): nothing in the product imports it.

PARITY UNPINNED: the reference (Rust, /root/reference/src/lib.rs) cannot be compiled in this image and ships no
golden audio, so nothing here is checked against the reference's own output.  What this file adds is a second
witness: it was written from src/lib.rs alone, as the same chain of pull iterators the reference is
(Sequencer -> Jitter -> Synthesize, one element per next()), not from oracle/grail_oracle.c, which is a flat
per-sample loop in C.  tests/test_restatement2.py and scripts/cross_check_restatements.py require the two to agree
BIT FOR BIT (audio, F_t, carrier phase, sample counts) on the survey's known-answer inputs and on random-voice
utterances at 16 / 22.05 / 48 kHz; a transcription error in either would have to be made twice, identically.

Every arithmetic step is one IEEE f32 operation: operands are np.float32 scalars or float32 arrays (NumPy never
contracts or reassociates; Python literals are weak scalars under NEP 50, so `1.0 - x` stays float32).
"""
from __future__ import annotations

import numpy as np

F = np.float32
NUM_FORMANTS = 8                      # src/lib.rs:24
DEFAULT_SAMPLE_RATE = F(44100.0)      # src/lib.rs:21


# ------------------------------------------------------------------------------------------------ helpers
class Rng:
    """`state: &mut u32` of random_f32 (src/lib.rs:36-55)"""

    def __init__(self, state: int):
        self.state = state & 0xFFFFFFFF

    def random_f32(self) -> np.float32:
        self.state = (self.state * 16807 + 1) & 0xFFFFFFFF                      # :40 wrapping_mul / wrapping_add
        res = np.uint32((self.state >> 9) | 0x3F800000)                          # :50
        return (res.view(np.float32) - F(1.5)) * F(2.0)                          # :54


def tan_approx(x):
    """src/lib.rs:63-70 (x: float32 scalar or array)"""
    num = (1.0 - x) * x * (5.0 - 4.0 * (x + 0.5) * (0.5 - x))
    den = (x + 0.5) * (5.0 - 4.0 * (1.0 - x) * x) * (0.5 - x)
    return num / den


def exp_approx(x):
    """src/lib.rs:75-82"""
    o = 1.0 - x
    o2 = o * o
    return o2 * o2 * o


def array_sum(a) -> np.float32:
    """Array::sum, src/lib.rs:123: iter().sum::<f32>() folds from 0.0 in index order"""
    s = F(0.0)
    for v in a:
        s = s + v
    return s


# ------------------------------------------------------------------------------------------------ value noise
class ValueNoise:
    """src/lib.rs:218-256"""

    def __init__(self, rng: Rng):
        self.current = rng.random_f32()
        self.next_ = rng.random_f32()
        self.phase = F(0.0)
        self.rng = Rng(rng.state)                                                # `state: *state`, a COPY (:235)

    def next(self, increment: np.float32) -> np.float32:
        self.phase = self.phase + increment                                      # :242
        if self.phase > 1.0:                                                     # :245
            self.phase = self.phase - F(1.0)
            self.current = self.next_
            self.next_ = self.rng.random_f32()
        return self.current * (1.0 - self.phase) + self.next_ * self.phase       # :254


class ArrayValueNoise:
    """src/lib.rs:261-307"""

    def __init__(self, rng: Rng):
        cur = np.zeros(NUM_FORMANTS, F)
        nxt = np.zeros(NUM_FORMANTS, F)
        for i in range(NUM_FORMANTS):                                            # interleaved draws (:275-278)
            cur[i] = rng.random_f32()
            nxt[i] = rng.random_f32()
        self.current, self.next_ = cur, nxt
        self.phase = F(0.0)
        self.rng = Rng(rng.state)                                                # a copy again (:284)

    def next(self, increment: np.float32) -> np.ndarray:
        self.phase = self.phase + increment                                      # :291
        if self.phase > 1.0:                                                     # :294
            self.phase = self.phase - F(1.0)
            self.current = self.next_
            self.next_ = np.array([self.rng.random_f32() for _ in range(NUM_FORMANTS)], F)   # :301
        return self.current * (1.0 - self.phase) + self.next_ * self.phase       # :305 (splat(1 - phase), splat(phase))


# ------------------------------------------------------------------------------------------------ SynthesisElem
class SynthesisElem:
    """src/lib.rs:316-337.  The six arrays are kept as one (6, 8) float32 block in declaration order
    (formant_freq, formant_bw, formant_smooth, formant_breath, formant_turb, formant_amp): blend and the element-wise
    updates act on every entry independently, so stacking them changes no rounding."""
    FREQ, BW, SMOOTH, BREATH, TURB, AMP = range(6)
    __slots__ = ("frequency", "arr")

    def __init__(self, frequency, arr):
        self.frequency = F(frequency)
        self.arr = arr

    @staticmethod
    def silent() -> "SynthesisElem":                                             # :367-377
        arr = np.zeros((6, NUM_FORMANTS), F)
        arr[SynthesisElem.FREQ] = 0.25
        arr[SynthesisElem.BW] = 0.25
        arr[SynthesisElem.SMOOTH] = 0.25
        return SynthesisElem(0.25, arr)

    @staticmethod
    def new_phoneme(freq, bw, smooth, turb, breath, amp) -> "SynthesisElem":     # :381-401
        amp = np.asarray(amp, F)
        arr = np.stack([np.asarray(freq, F), np.asarray(bw, F), np.asarray(smooth, F), np.asarray(breath, F),
                        np.asarray(turb, F), amp / array_sum(amp)]).astype(F)
        return SynthesisElem(0.0, arr).resample(F(1.0), DEFAULT_SAMPLE_RATE)

    def blend(self, other: "SynthesisElem", alpha) -> "SynthesisElem":           # :404-414, Array::blend :135
        return SynthesisElem(self.frequency * (1.0 - alpha) + other.frequency * alpha,
                             self.arr * (1.0 - alpha) + other.arr * alpha)

    def resample(self, old_rate, new_rate) -> "SynthesisElem":                   # :418-440
        scale = F(old_rate) / F(new_rate)
        ff = self.arr[self.FREQ] * scale
        arr = self.arr.copy()
        arr[self.FREQ] = np.fmin(ff, F(0.5))
        arr[self.BW] = self.arr[self.BW] * scale
        arr[self.SMOOTH] = self.arr[self.SMOOTH] * scale
        arr[self.AMP] = np.where(ff > 0.5, F(0.0), self.arr[self.AMP])
        return SynthesisElem(np.fmin(self.frequency * scale, F(0.5)), arr)

    def copy_with_frequency(self, frequency) -> "SynthesisElem":                 # :445-450
        return SynthesisElem(np.fmin(F(frequency), F(0.5)), self.arr)

    def copy_silent(self) -> "SynthesisElem":                                    # :454-459
        arr = self.arr.copy()
        arr[self.AMP] = 0.0
        return SynthesisElem(self.frequency, arr)


class SequenceElem:                                                              # :814-824
    __slots__ = ("elem", "length", "blend_length")

    def __init__(self, elem, length, blend_length):
        self.elem, self.length, self.blend_length = elem, F(length), F(blend_length)


class Voice:                                                                     # the scalars of :696-717 the path reads
    def __init__(self, sample_rate, jitter_frequency, jitter_delta_frequency, jitter_delta_formant_frequency,
                 jitter_delta_amplitude):
        self.sample_rate = F(sample_rate)
        self.jitter_frequency = F(jitter_frequency)
        self.jitter_delta_frequency = F(jitter_delta_frequency)
        self.jitter_delta_formant_frequency = F(jitter_delta_formant_frequency)
        self.jitter_delta_amplitude = F(jitter_delta_amplitude)


# ------------------------------------------------------------------------------------------------ the three iterators
class Sequencer:
    """src/lib.rs:828-953"""

    def __init__(self, upstream, voice: Voice):
        self.iter = iter(upstream)
        self.delta_time = F(1.0) / voice.sample_rate                             # :944
        self.cur_elem = None
        self.next_elem = None
        self.time = F(0.0)
        self.trace_time = None

    def _pull(self):
        return next(self.iter, None)

    def __iter__(self):
        return self

    def __next__(self) -> SynthesisElem:
        self.time = self.time - self.delta_time                                  # :861
        if self.time < 0.0:                                                      # :864
            if self.cur_elem is not None and self.next_elem is not None:         # :868
                a = self.next_elem
                self.cur_elem = self.next_elem
                self.next_elem = self._pull()
                self.time = self.time + a.length                                 # :873
            elif self.cur_elem is None and self.next_elem is None:               # :876
                self.cur_elem = self._pull()
                self.next_elem = self._pull()
                if self.cur_elem is not None:
                    self.time = self.time + self.cur_elem.length                 # :882
            else:
                raise StopIteration                                              # :886
        a = self.cur_elem
        if a is None:
            raise StopIteration                                                  # :930
        b = a.elem
        c = self.next_elem.elem if self.next_elem is not None else None
        if b is None and c is None:
            return SynthesisElem.silent()                                        # :924-927
        alpha = np.fmin(self.time / a.blend_length, F(1.0))                      # :899 / :908 / :917
        if b is not None and c is not None:
            return c.blend(b, alpha)                                             # :902
        if b is not None:
            return b.copy_silent().blend(b, alpha)                               # :911
        return c.blend(c.copy_silent(), alpha)                                   # :920


class Jitter:
    """src/lib.rs:724-801"""

    def __init__(self, upstream, seed: int, voice: Voice):
        self.iter = iter(upstream)
        rng = Rng(seed)                                                          # `mut seed`: ONE running state (:786)
        self.freq_noise = ValueNoise(rng)                                        # :789
        self.formant_freq_noise = ArrayValueNoise(rng)                           # :790
        self.formant_amp_noise = ArrayValueNoise(rng)                            # :791
        self.frequency = voice.jitter_frequency
        self.delta_frequency = voice.jitter_delta_frequency
        self.delta_formant_freq = voice.jitter_delta_formant_frequency
        self.delta_amplitude = voice.jitter_delta_amplitude

    def __iter__(self):
        return self

    def __next__(self) -> SynthesisElem:
        elem = next(self.iter)                                                   # :754
        freq = self.freq_noise.next(self.frequency)
        formant_freq = self.formant_freq_noise.next(self.frequency)
        formant_amp = self.formant_amp_noise.next(self.frequency)
        arr = elem.arr.copy()
        frequency = elem.frequency + freq * self.delta_frequency                 # :763
        arr[SynthesisElem.FREQ] = arr[SynthesisElem.FREQ] + formant_freq * self.delta_formant_freq      # :764
        delta = (formant_amp + F(1.0)) * (F(0.5) * self.delta_amplitude)         # :768-769
        mul = F(1.0) - delta                                                     # :772
        arr[SynthesisElem.AMP] = arr[SynthesisElem.AMP] * mul                    # :773
        return SynthesisElem(frequency, arr)


class Synthesize:
    """src/lib.rs:470-600"""

    def __init__(self, upstream):
        self.iter = iter(upstream)
        self.phase = F(0.0)
        self.filter_state_a = np.zeros(NUM_FORMANTS, F)
        self.filter_state_b = np.zeros(NUM_FORMANTS, F)
        self.filter_state_c = np.zeros(NUM_FORMANTS, F)
        self.rng = Rng(0)                                                        # seed: 0 (:594)
        self.tap_frequency = []
        self.tap_phase = []

    def __iter__(self):
        return self

    def __next__(self) -> np.float32:
        elem = next(self.iter)                                                   # :499
        f = elem.frequency
        self.tap_frequency.append(f)
        self.tap_phase.append(self.phase)
        if self.phase < f:                                                       # :503
            t = self.phase / f
            polyblep = F(2.0) * t - (t * t) - F(1.0)
        elif self.phase > (F(1.0) - f):                                          # :507
            t = (self.phase - F(1.0)) / f
            polyblep = (t * t) + F(2.0) * t + F(1.0)
        else:
            polyblep = F(0.0)
        saw = (F(2.0) * self.phase - F(1.0)) - polyblep                          # :517
        self.phase = self.phase + f                                              # :520
        if self.phase >= 1.0:                                                    # :523
            self.phase = self.phase - F(1.0)
        noise = self.rng.random_f32()                                            # :528
        A = elem.arr
        breath, turb = A[SynthesisElem.BREATH], A[SynthesisElem.TURB]
        # blend_multiple(self, other, alpha) = self * (1 - alpha) + other * alpha   (:141-143)
        noise_wave = saw * (F(1.0) - breath) + noise * breath                    # :531
        alpha = exp_approx(A[SynthesisElem.SMOOTH])                              # :535
        self.filter_state_a = self.filter_state_a + (F(1.0) - alpha) * (noise_wave - self.filter_state_a)   # :538
        glottal = self.filter_state_a
        turbulence = glottal * (F(1.0) * (F(1.0) - turb) + noise * turb)         # :544-545
        v0 = turbulence * A[SynthesisElem.AMP]                                   # :550
        g = tan_approx(A[SynthesisElem.FREQ])                                    # :555
        k = A[SynthesisElem.BW] / A[SynthesisElem.FREQ]                          # :558
        a1 = F(1.0) / (F(1.0) + g * (g + k))                                     # :560
        a2 = g * a1
        a3 = g * a2
        v3 = v0 - self.filter_state_c                                            # :565
        v1 = a1 * self.filter_state_b + a2 * v3
        v2 = self.filter_state_c + a2 * self.filter_state_b + a3 * v3
        self.filter_state_b = F(2.0) * v1 - self.filter_state_b                  # :570
        self.filter_state_c = F(2.0) * v2 - self.filter_state_c
        return array_sum(v1) * F(0.5)                                            # :574


# ------------------------------------------------------------------------------------------------ voices::generic()
_GENERIC = {   # src/voices/generic.rs:9-32; MKPHON order: freq, bw, smooth, turb, breath, amp (src/voices/mod.rs:7-14)
    "a": ([910.0, 1271.0, 2851.0, 3213.0, 1200.0, 2000.0, 3000.0, 4000.0], [60.0, 160.0, 180.0, 200.0, 100.0, 100.0, 100.0, 100.0],
          [1600.0] * 8, [0.2, 0.2, 0.1, 0.0, 0.0, 0.0, 0.0, 0.0], [0.5, 0.2, 0.05, 0.0, 0.0, 0.0, 0.0, 0.0],
          [0.3, 0.3, 0.2, 0.1, 0.0, 0.0, 0.0, 0.0]),
    "e": ([910.0, 1871.0, 2851.0, 3213.0, 1200.0, 2000.0, 3000.0, 4000.0], [80.0, 180.0, 180.0, 200.0, 100.0, 100.0, 100.0, 100.0],
          [1600.0] * 8, [0.2, 0.4, 0.4, 0.4, 0.4, 0.4, 0.4, 0.4], [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.1, 0.1],
          [0.5, 0.4, 0.3, 0.2, 0.0, 0.0, 0.0, 0.0]),
}


def generic_voice():
    """(Voice, {"a": SynthesisElem, "e": SynthesisElem}, center_frequency), src/voices/generic.rs:5-40"""
    ph = {k: SynthesisElem.new_phoneme(*v) for k, v in _GENERIC.items()}
    R = DEFAULT_SAMPLE_RATE
    return Voice(R, F(16.0) / R, F(6.0) / R, F(6.0) / R, F(0.2)), ph, F(120.0) / R


def phoneme_sequence(ids, phonemes, center_frequency):
    """Intonator (:1057-1075: length 0.5, blend 0.5, centre frequency) + Selector (:987-1005) for phoneme ids
    0 Silence, 1 Stop, 2 Glide (no sound), 3 A, 4 E"""
    for pid in ids:
        elem = None
        if pid == 3:
            elem = phonemes["a"].copy_with_frequency(center_frequency)
        elif pid == 4:
            elem = phonemes["e"].copy_with_frequency(center_frequency)
        yield SequenceElem(elem, 0.5, 0.5)


# ------------------------------------------------------------------------------------------------ record-level entry
def sequence_from_records(records):
    """SequenceElems from packed grail_seq_elem records (include/grail_cuda.h; numpy structured array)"""
    for r in records:
        elem = None
        if int(r["has_elem"]):
            e = r["elem"]
            arr = np.stack([e["formant_freq"], e["formant_bw"], e["formant_smooth"], e["formant_breath"],
                            e["formant_turb"], e["formant_amp"]]).astype(F)
            elem = SynthesisElem(e["frequency"], arr)
        yield SequenceElem(elem, r["length"], r["blend_length"])


def synthesize(sequence, voice: Voice, jitter_seed: int, trace: bool = False):
    """`sequence.sequence(voice).jitter(seed, voice).synthesize()` drained (examples/cli.rs:175-184)"""
    syn = Synthesize(Jitter(Sequencer(sequence, voice), int(jitter_seed), voice))
    out = np.array(list(syn), F)
    if trace:
        return out, {"frequency": np.array(syn.tap_frequency, F), "carrier_phase": np.array(syn.tap_phase, F)}
    return out


def synthesize_records(records, voice_params, trace: bool = False):
    """same, from the C ABI's records (one utterance)"""
    vp = voice_params
    voice = Voice(vp["sample_rate"], vp["jitter_frequency"], vp["jitter_delta_frequency"],
                  vp["jitter_delta_formant_frequency"], vp["jitter_delta_amplitude"])
    assert int(vp["synth_seed"]) == 0, "the reference hard-codes the Synthesize noise seed to 0 (src/lib.rs:594)"
    return synthesize(sequence_from_records(records), voice, int(vp["jitter_seed"]), trace)


def fnv(samples: np.ndarray) -> int:
    """word-wise FNV-1a over the f32 bit patterns"""
    h = 2166136261
    for w in np.ascontiguousarray(samples, F).view(np.uint32).tolist():
        h = ((h ^ w) * 16777619) & 0xFFFFFFFF
    return h
