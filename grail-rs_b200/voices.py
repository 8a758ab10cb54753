"""voices::generic() (reference src/voices/generic.rs:5-40, src/voices/mod.rs:7-14): table data and the
per-phoneme constructors run on the host; only the resulting numbers cross the C ABI."""
from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np

from .synth import DEFAULT_SAMPLE_RATE, SynthesisElem, Voice, f32

MKPHON = SynthesisElem.new_phoneme  # src/voices/mod.rs:7-14: (freq, bw, smooth, turb, breath, amp)


@dataclass
class VoiceStorage:
    """make_phonemes!(A a test, E e test) (src/lib.rs:653-689)"""
    a: SynthesisElem
    e: SynthesisElem

    def get(self, phoneme):
        from .text import Phoneme
        if phoneme in (Phoneme.Silence, Phoneme.Stop, Phoneme.Glide):   # src/lib.rs:666
            return None
        return {Phoneme.A: self.a, Phoneme.E: self.e}[phoneme]

    def for_all(self, func):
        from .text import Phoneme
        func(Phoneme.A, self.a)
        func(Phoneme.E, self.e)


def generic() -> Voice:
    return Voice(
        sample_rate=DEFAULT_SAMPLE_RATE,
        phonemes=VoiceStorage(
            a=MKPHON([910.0, 1271.0, 2851.0, 3213.0, 1200.0, 2000.0, 3000.0, 4000.0],
                     [60.0, 160.0, 180.0, 200.0, 100.0, 100.0, 100.0, 100.0],
                     [1600.0] * 8,
                     [0.2, 0.2, 0.1, 0.0, 0.0, 0.0, 0.0, 0.0],
                     [0.5, 0.2, 0.05, 0.0, 0.0, 0.0, 0.0, 0.0],
                     [0.3, 0.3, 0.2, 0.1, 0.0, 0.0, 0.0, 0.0]),
            e=MKPHON([910.0, 1871.0, 2851.0, 3213.0, 1200.0, 2000.0, 3000.0, 4000.0],
                     [80.0, 180.0, 180.0, 200.0, 100.0, 100.0, 100.0, 100.0],
                     [1600.0] * 8,
                     [0.2, 0.4, 0.4, 0.4, 0.4, 0.4, 0.4, 0.4],
                     [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.1, 0.1],
                     [0.5, 0.4, 0.3, 0.2, 0.0, 0.0, 0.0, 0.0]),
        ),
        center_frequency=f32(f32(120.0) / DEFAULT_SAMPLE_RATE),
        jitter_frequency=f32(f32(16.0) / DEFAULT_SAMPLE_RATE),
        jitter_delta_frequency=f32(f32(6.0) / DEFAULT_SAMPLE_RATE),
        jitter_delta_formant_frequency=f32(f32(6.0) / DEFAULT_SAMPLE_RATE),
        jitter_delta_amplitude=f32(0.2),
    )


def at_sample_rate(voice: Voice, sample_rate: float) -> Voice:
    """The reference has no whole-voice resample; a voice at rate R is assembled by hand:
    each element `.resample(old, R)`, scalars rescaled by old/R."""
    R = f32(sample_rate)
    old = voice.sample_rate
    k = f32(old / R)
    ph = VoiceStorage(a=voice.phonemes.a.resample(old, R), e=voice.phonemes.e.resample(old, R))
    return replace(voice, sample_rate=R, phonemes=ph,
                   center_frequency=f32(f32(voice.center_frequency * old) / R),
                   jitter_frequency=f32(f32(voice.jitter_frequency * old) / R),
                   jitter_delta_frequency=f32(f32(voice.jitter_delta_frequency * old) / R),
                   jitter_delta_formant_frequency=f32(f32(voice.jitter_delta_formant_frequency * old) / R)) if k else voice
