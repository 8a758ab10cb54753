"""grail_cuda_transcribe_batch (host-only C++) against (1) the six asserting tests the reference holds
for its Transcriber (src/lib.rs:1210-1358; the only known answers the reference ships), restated here input for
input, and (2) the Python mirror of the same algorithm on random texts and rule sets."""
import ctypes as C

import numpy as np
import pytest

from grail_rs_b200 import _ffi
from grail_rs_b200 import text as T

P = T.Phoneme
R = T.TranscriptionRule


def lang(*rules, case_sensitive=False):
    return T.Language(rules=tuple(rules), case_sensitive=case_sensitive)


# (text, rules, expected) -- reference src/lib.rs:1211-1358, all with an empty initial buffer
REFERENCE_CASES = [
    ("abc", [R("ab", (P.A,)), R("c", (P.E,))], [P.A, P.E]),                                     # transcribe_unique
    ("abacab", [R("ab", (P.A,)), R("ac", (P.E,))], [P.A, P.E, P.A]),                            # transcribe_same_start
    ("aaa", [R("a", (P.A,)), R("aa", (P.E,))], [P.E, P.A]),                                     # ..._different_length
    ("ae", [R("a", (P.A,)), R("aa", (P.E,)), R("e", (P.E,))], [P.A, P.E]),                      # ..._length_cutoff
    ("abuac", [R("ab", (P.A,)), R("ac", (P.E,))], [P.A, P.Silence, P.E]),                       # transcribe_skip_no_matches
    ("abaca", [R("ab", (P.A,)), R("ac", (P.E,))], [P.A, P.E, P.Silence]),                       # ..._partial_match_at_end
]


@pytest.mark.parametrize("text,rules,want", REFERENCE_CASES)
def test_reference_transcriber_tests(text, rules, want):
    ids, offs = T.transcribe_batch([text], lang(*rules), leading_silence=False)
    assert list(offs) == [0, len(want)]
    assert [int(x) for x in ids] == [int(p) for p in want]


def test_leading_silence_and_generic_language():
    """IntoTranscriber::transcribe starts with one Silence (:1201); text "a" => [Silence, A]"""
    ids, offs = T.transcribe_batch(["a", "", "a pie i oui e a"], T.generic_language())
    assert list(ids[offs[0]:offs[1]]) == [P.Silence, P.A]
    assert list(ids[offs[1]:offs[2]]) == [P.Silence]
    assert list(ids[offs[2]:offs[3]]) == [int(p) for p in T.transcribe("a pie i oui e a", T.generic_language())]


def _random_language(rng, alphabet):
    n = int(rng.integers(1, 12))
    strings = set()
    while len(strings) < n:
        k = int(rng.integers(1, 5))
        strings.add("".join(rng.choice(alphabet, k)))
    rules = [R(s, tuple(int(x) for x in rng.integers(0, 5, int(rng.integers(1, 4))))) for s in sorted(strings)]
    return T.Language(rules=tuple(rules), case_sensitive=bool(rng.integers(0, 2)))


@pytest.mark.parametrize("seed", range(8))
def test_random_texts_match_python_mirror(seed):
    """random rule sets (shared prefixes, multi-phoneme rules, non-ASCII scalars, upper case in the text) and texts"""
    rng = np.random.default_rng(seed)
    alphabet = np.array(list("abcé") if seed % 2 else list("abAB ü√"))
    language = _random_language(rng, np.array([c for c in alphabet if not c.isupper()]))
    texts = ["".join(rng.choice(alphabet, int(rng.integers(0, 40)))) for _ in range(200)]
    for ls in (True, False):
        ids, offs = T.transcribe_batch(texts, language, leading_silence=ls, n_threads=3)
        assert len(offs) == len(texts) + 1
        for i, t in enumerate(texts):
            want = [int(p) for p in T.Transcriber(t, language.rules, language.case_sensitive,
                                                  buffer=(P.Silence,) if ls else ())]
            assert [int(x) for x in ids[offs[i]:offs[i + 1]]] == want, (seed, t, language)


def test_threads_do_not_change_the_result():
    rng = np.random.default_rng(99)
    texts = ["".join(rng.choice(list("aeiou p"), int(rng.integers(0, 200)))) for _ in range(3000)]
    a = T.transcribe_batch(texts, T.generic_language(), n_threads=1)
    b = T.transcribe_batch(texts, T.generic_language(), n_threads=0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_errors():
    L = _ffi.lib()
    offs = np.zeros(2, np.uint32)
    ph = (C.c_uint8 * 1)(3)

    def call(rules, text=b"ab", ids=None, cap=0):
        arr = (C.c_char_p * 1)(text)
        rr = (_ffi.TranscriptionRuleC * max(len(rules), 1))()
        for i, (s, n) in enumerate(rules):
            rr[i].string = s
            rr[i].phonemes = C.cast(ph, C.POINTER(C.c_uint8))
            rr[i].n_phonemes = n
        return L.grail_cuda_transcribe_batch(C.cast(arr, C.c_void_p), None, 1, C.cast(rr, C.c_void_p), len(rules), 0, 0,
                                             _ffi.ptr(ids), cap, _ffi.ptr(offs), 1)

    assert call([(b"b", 1), (b"a", 1)]) == _ffi.ERR_INVALID_ARG        # not sorted
    assert call([(b"", 1)]) == _ffi.ERR_INVALID_ARG                    # empty rule string: the reference never ends
    assert call([(b"a", 0)]) == _ffi.ERR_INVALID_ARG                   # no phonemes: the reference never ends
    assert call([(b"a", 1), (b"b", 1)]) == 0 and offs[1] == 2          # counting call
    ids = np.zeros(1, np.uint8)
    assert call([(b"a", 1), (b"b", 1)], ids=ids, cap=1) == _ffi.ERR_COUNT_MISMATCH
    assert call([], text=b"xyz") == 0 and offs[1] == 3                 # no rules: every character is a Silence
    # malformed UTF-8 is not trusted: each bad byte becomes one replacement character (=> one Silence here)
    assert call([(b"a", 1)], text=b"a\xffa\xc3") == 0 and offs[1] == 4
