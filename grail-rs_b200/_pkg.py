"""public names of the package (see grail_rs_b200/__init__.py for why this indirection exists)"""
from . import _ffi, languages, voices  # noqa: F401
from ._ffi import ELEM_DT, F32, I16, PHONEME_ELEM_DT, SEQ_ELEM_DT, VOICE_DT, GrailError
from .synth import (DEFAULT_SAMPLE_RATE, NUM_FORMANTS, Context, Jitter, Plan, SequenceElem, Sequencer, Stream, Synthesize,
                    SynthesisElem, Voice, count_samples, default_context, pack_sequence, pull_streams, save_wav, sequence)
from .text import (Intonator, Language, Phoneme, PhonemeElem, Selector, Transcriber, TranscriptionRule, intonate, select,
                   transcribe)

__all__ = [
    "ELEM_DT", "SEQ_ELEM_DT", "PHONEME_ELEM_DT", "VOICE_DT", "F32", "I16", "GrailError", "DEFAULT_SAMPLE_RATE", "NUM_FORMANTS", "Context",
    "Plan", "Stream", "Jitter", "Sequencer", "Synthesize", "SequenceElem", "SynthesisElem", "Voice", "count_samples",
    "default_context", "pack_sequence", "pull_streams", "save_wav", "sequence", "Intonator", "Language", "Phoneme", "PhonemeElem", "Selector",
    "Transcriber", "TranscriptionRule", "intonate", "select", "transcribe", "voices", "languages",
]
