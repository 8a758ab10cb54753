"""Host-side front-end, per character / per phoneme (stays on the host per north_star): Phoneme (reference
src/lib.rs:632-649), Transcriber (:1098-1207), Intonator (:1049-1089), Selector (:979-1022), and the generic
language (src/languages/mod.rs:4-34).  Small and sequential by nature; mirrored so the chain reads like the
reference's examples/cli.rs:175-184."""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Sequence

import numpy as np

from .synth import SequenceElem, Sequencer, Voice, f32


class Phoneme(enum.IntEnum):
    Silence = 0
    Stop = 1
    Glide = 2
    A = 3
    E = 4


@dataclass(frozen=True)
class TranscriptionRule:
    string: str
    phonemes: Sequence[Phoneme]


@dataclass(frozen=True)
class Language:
    rules: Sequence[TranscriptionRule]   # sorted by string
    case_sensitive: bool


def generic_language() -> Language:
    P = Phoneme
    return Language(case_sensitive=False, rules=(
        TranscriptionRule("a", (P.A,)), TranscriptionRule("e", (P.E,)), TranscriptionRule("i", (P.A,)),
        TranscriptionRule("ii", (P.E, P.A)), TranscriptionRule("oui", (P.A, P.E, P.A)),
        TranscriptionRule("p", (P.Silence,)),
    ))


def _partition_point(seq, pred) -> int:
    lo, hi = 0, len(seq)
    while lo < hi:
        mid = (lo + hi) // 2
        if pred(seq[mid]):
            lo = mid + 1
        else:
            hi = mid
    return lo


class _Peekable:
    def __init__(self, it):
        self._it = iter(it)
        self._buf = []

    def peek(self):
        if not self._buf:
            try:
                self._buf.append(next(self._it))
            except StopIteration:
                return None
        return self._buf[0]

    def next(self):
        if self._buf:
            return self._buf.pop()
        try:
            return next(self._it)
        except StopIteration:
            return None


class Transcriber:
    """longest-match find-and-replace over a sorted rule list by incremental range narrowing
    (src/lib.rs:1116-1191).  The initial buffer holds one Silence (:1201)."""

    def __init__(self, chars: Iterable[str], ruleset: Sequence[TranscriptionRule], case_sensitive: bool,
                 buffer: Sequence[Phoneme] = (Phoneme.Silence,)):
        self.iter = _Peekable(chars)
        self.ruleset = list(ruleset)
        self.case_sensitive = case_sensitive
        self.buffer: List[Phoneme] = list(buffer)

    def __iter__(self):
        return self

    def __next__(self) -> Phoneme:
        search_min, search_max, index = 0, len(self.ruleset), 0
        while not self.buffer:
            ch = self.iter.peek()
            if ch is None:
                raise StopIteration
            if not self.case_sensitive and ch.isascii():
                ch = ch.lower()
            window = self.ruleset[search_min:search_max]

            def nth(rule, i=index):
                return rule.string[i] if i < len(rule.string) else None

            new_min = _partition_point(window, lambda r: nth(r) is None or nth(r) < ch) + search_min
            new_max = _partition_point(window, lambda r: nth(r) is not None and nth(r) <= ch) + search_min
            if new_min >= new_max and len(self.ruleset[search_min].string) == index:
                self.buffer = list(self.ruleset[search_min].phonemes)
            elif new_min >= new_max:
                self.buffer = [Phoneme.Silence]
                self.iter.next()
            else:
                search_min, search_max, index = new_min, new_max, index + 1
                self.iter.next()
                if self.iter.peek() is None and len(self.ruleset[search_min].string) == index:
                    self.buffer = list(self.ruleset[search_min].phonemes)
                elif self.iter.peek() is None:
                    self.buffer = [Phoneme.Silence]
        return self.buffer.pop(0)

    def intonate(self, language: Language, voice: Voice) -> "Intonator":
        return Intonator(self, voice)


@dataclass
class PhonemeElem:
    """src/lib.rs:961-973"""
    phoneme: Phoneme
    length: np.float32
    blend_length: np.float32
    frequency: np.float32


class Intonator:
    """the reference's stub: constant length / blend / pitch (src/lib.rs:1057-1075)"""

    def __init__(self, phonemes: Iterable[Phoneme], voice: Voice):
        self.iter = iter(phonemes)
        self.center_frequency = voice.center_frequency

    def __iter__(self):
        return self

    def __next__(self) -> PhonemeElem:
        return PhonemeElem(next(self.iter), f32(0.5), f32(0.5), self.center_frequency)

    def select(self, voice: Voice) -> "Selector":
        return Selector(self, voice)


class Selector:
    """phoneme -> Option<SynthesisElem> with copy_with_frequency (src/lib.rs:987-1005)"""

    def __init__(self, phoneme_elems: Iterable[PhonemeElem], voice: Voice):
        self.iter = iter(phoneme_elems)
        self.voice_storage = voice.phonemes

    def __iter__(self):
        return self

    def __next__(self) -> SequenceElem:
        ph = next(self.iter)
        elem = self.voice_storage.get(ph.phoneme)
        return SequenceElem.new(elem.copy_with_frequency(ph.frequency) if elem is not None else None,
                                ph.length, ph.blend_length)

    def sequence(self, voice: Voice) -> Sequencer:
        return Sequencer(self, voice)


def transcribe(chars: Iterable[str], language: Language) -> Transcriber:
    """IntoTranscriber::transcribe (src/lib.rs:1193-1205)"""
    return Transcriber(chars, language.rules, language.case_sensitive)


def intonate(phonemes: Iterable[Phoneme], language: Language, voice: Voice) -> Intonator:
    return Intonator(phonemes, voice)


def select(phoneme_elems: Iterable[PhonemeElem], voice: Voice) -> Selector:
    return Selector(phoneme_elems, voice)


def transcribe_batch(texts: Sequence[str], language: Language, leading_silence: bool = True, n_threads: int = 0):
    """native, multi-threaded Transcriber over a batch of texts (grail_cuda_transcribe_batch; host only).
    Returns (phoneme ids uint8, utterance offsets uint32) in the layout Context.plan_phonemes takes."""
    import ctypes as C

    from . import _ffi

    L = _ffi.lib()
    enc = [t.encode("utf-8") for t in texts]
    n = len(enc)
    arr = (C.c_char_p * max(n, 1))(*enc)
    nbytes = (C.c_size_t * max(n, 1))(*[len(b) for b in enc])
    rules = (_ffi.TranscriptionRuleC * max(len(language.rules), 1))()
    keep = []
    for i, r in enumerate(language.rules):
        ph = (C.c_uint8 * max(len(r.phonemes), 1))(*[int(p) for p in r.phonemes])
        keep.append(ph)
        rules[i].string = r.string.encode("utf-8")
        rules[i].phonemes = C.cast(ph, C.POINTER(C.c_uint8))
        rules[i].n_phonemes = len(r.phonemes)
    offs = np.zeros(n + 1, np.uint32)
    args = (C.cast(arr, C.c_void_p), C.cast(nbytes, C.c_void_p), n, C.cast(rules, C.c_void_p), len(language.rules),
            int(language.case_sensitive), int(leading_silence))
    rc = L.grail_cuda_transcribe_batch(*args, None, 0, _ffi.ptr(offs), n_threads)
    if rc:
        raise _ffi.GrailError(rc)
    ids = np.zeros(int(offs[-1]), np.uint8)
    rc = L.grail_cuda_transcribe_batch(*args, _ffi.ptr(ids), ids.size, _ffi.ptr(offs), n_threads)
    if rc:
        raise _ffi.GrailError(rc)
    return ids, offs
