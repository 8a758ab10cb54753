"""Host-side mirror of the reference interface: voice tables bit-identical to the oracle's restatement,
and the reference's own six Transcriber tests (src/lib.rs:1210-1358) re-run against the mirrored Transcriber."""
import numpy as np

import grail_rs_b200 as g
from grail_rs_b200 import Phoneme as P
from grail_rs_b200 import Transcriber, TranscriptionRule as R


def test_generic_voice_matches_oracle_bits(oracle):
    ov = oracle.generic_voice()
    v = g.voices.generic()
    for name in ("a", "e"):
        rec = getattr(v.phonemes, name).to_record()
        assert rec.tobytes() == ov["phonemes"][name].tobytes(), name
    vp = v.params(5)
    assert vp.tobytes() == oracle.voice_params(ov, 5).tobytes()
    assert v.center_frequency.view(np.uint32) == ov["center_frequency"].view(np.uint32)


def test_resampled_voice_matches_oracle_bits(oracle):
    for rate in (16000.0, 22050.0, 48000.0):
        ov = oracle.generic_voice(rate)
        v = g.voices.at_sample_rate(g.voices.generic(), rate)
        for name in ("a", "e"):
            assert getattr(v.phonemes, name).to_record().tobytes() == ov["phonemes"][name].tobytes(), (rate, name)
        assert v.params(1).tobytes() == oracle.voice_params(ov, 1).tobytes()
        assert v.center_frequency.view(np.uint32) == ov["center_frequency"].view(np.uint32)


def test_selector_chain_matches_oracle_records(oracle):
    v = g.voices.generic()
    lang = g.languages.generic()
    seq = g.pack_sequence(g.transcribe("a e", lang).intonate(lang, v).select(v))
    want = oracle.select([0, 3, 0, 4], oracle.generic_voice())
    assert seq.tobytes() == want.tobytes()


def test_elem_helpers():
    v = g.voices.generic()
    a, e = v.phonemes.a, v.phonemes.e
    s = g.SynthesisElem.silent()
    assert s.frequency == 0.25 and (s.formant_amp == 0).all() and (s.formant_bw == 0.25).all()
    assert (a.copy_silent().formant_amp == 0).all() and (a.copy_silent().formant_freq == a.formant_freq).all()
    assert a.copy_with_frequency(0.9).frequency == 0.5
    b = a.blend(e, 0.25)
    np.testing.assert_allclose(b.formant_freq, a.formant_freq * 0.75 + e.formant_freq * 0.25, rtol=1e-6)
    # resampling to a lower rate zeroes formants above Nyquist (src/lib.rs:433-435)
    lo = a.resample(44100.0, 6000.0)
    assert lo.formant_amp[3] == 0 and lo.formant_freq[3] == 0.5 and lo.formant_amp[0] > 0


def _t(text, rules):
    return Transcriber(text, rules, False, buffer=())


def test_transcribe_unique():
    assert list(_t("abc", [R("ab", (P.A,)), R("c", (P.E,))])) == [P.A, P.E]


def test_transcribe_same_start():
    assert list(_t("abacab", [R("ab", (P.A,)), R("ac", (P.E,))])) == [P.A, P.E, P.A]


def test_transcribe_same_char_different_length():
    assert list(_t("aaa", [R("a", (P.A,)), R("aa", (P.E,))])) == [P.E, P.A]


def test_transcribe_same_char_different_length_cutoff():
    assert list(_t("ae", [R("a", (P.A,)), R("aa", (P.E,)), R("e", (P.E,))])) == [P.A, P.E]


def test_transcribe_skip_no_matches():
    assert list(_t("abuac", [R("ab", (P.A,)), R("ac", (P.E,))])) == [P.A, P.Silence, P.E]


def test_transcribe_skip_partial_match_at_end():
    assert list(_t("abaca", [R("ab", (P.A,)), R("ac", (P.E,))])) == [P.A, P.E, P.Silence]


def test_transcribe_starts_with_silence():
    # IntoTranscriber::transcribe seeds the buffer with one Silence (src/lib.rs:1201)
    assert list(g.transcribe("a", g.languages.generic())) == [P.Silence, P.A]
