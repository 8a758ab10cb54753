"""Selector (src/lib.rs:979-1005) and the Intonator stub (:1057-1075) on the device.  A plan built from
phoneme ids / PhonemeElem records must give the very same samples, bit for bit, as the plan built from the
Sequencer records the host front-end (text.py, a mirror of the reference) expands them to."""
import numpy as np
import pytest

import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from grail_rs_b200._ffi import PHONEME_ELEM_DT, GrailError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


def _run(plan):
    plan.launch()
    out = plan.read_output()
    oo = plan.out_offsets.copy()
    plan.close()
    return out, oo


def test_phoneme_ids_match_host_expansion(ctx, oracle):
    v = g.voices.generic()
    lists = [[0, 3, 4, 3], [3], [0, 0, 4], [], [1, 3, 2, 4, 0], [4, 4, 4, 3, 3, 0, 3]]
    elems, offs, vp = W.from_phonemes(lists, v)
    want, woo = _run(ctx.plan(elems, offs, vp))
    ids = np.concatenate([np.asarray(p, np.uint8) for p in lists])
    cf = np.full(len(lists), v.center_frequency, np.float32)
    got, goo = _run(ctx.plan_phonemes(ids, offs, v.storage(), vp, center_frequency=cf))
    assert np.array_equal(goo, woo)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # and the oracle agrees on one of them (the usual tolerance)
    ref, _, _ = oracle.synthesize(elems[offs[0]:offs[1]], vp[0])
    st = W.parity_stats(got[goo[0]:goo[1]], ref)
    assert st["max_abs"] <= 1e-4 and st["snr_db"] >= 90.0, st


def test_phoneme_elems_per_utterance_voices(ctx):
    """PhonemeElem records with their own lengths and pitches (incl. one above 0.5: copy_with_frequency clamps, :448)
    and two voice storages picked per utterance"""
    v1 = g.voices.generic()
    v2 = g.voices.generic()
    st = np.stack([v1.storage(), v1.storage()])
    st[1]["formant_freq"] *= np.float32(1.07)          # a second, different voice
    st[1]["formant_amp"][:, 3] = 0.0                    # with another set of active formants
    rng = np.random.default_rng(5)
    lists = [[0, 3, 4], [4, 3], [3, 0, 4, 4], [4]]
    utt_storage = np.array([0, 1, 1, 0], np.uint32)
    n = sum(len(p) for p in lists)
    ph = np.zeros(n, PHONEME_ELEM_DT)
    ph["phoneme"] = np.concatenate([np.asarray(p) for p in lists])
    ph["length"] = rng.uniform(0.05, 0.4, n).astype(np.float32)
    ph["blend_length"] = rng.uniform(0.01, 0.3, n).astype(np.float32)
    ph["frequency"] = rng.uniform(80.0, 300.0, n).astype(np.float32) / np.float32(44100.0)
    ph["frequency"][1] = np.float32(0.75)
    offs = np.concatenate([[0], np.cumsum([len(p) for p in lists])]).astype(np.uint32)
    vp = np.zeros(len(lists), g.VOICE_DT)
    vp[:] = v1.params(0)
    vp["jitter_seed"] = np.arange(len(lists))
    # host expansion: Selector::next
    elems = np.zeros(n, g.SEQ_ELEM_DT)
    for u in range(len(lists)):
        for p in range(offs[u], offs[u + 1]):
            elems[p]["length"] = ph[p]["length"]
            elems[p]["blend_length"] = ph[p]["blend_length"]
            pid = int(ph[p]["phoneme"])
            if pid >= 3:
                elems[p]["has_elem"] = 1
                elems[p]["elem"] = st[utt_storage[u]][pid - 3]
                elems[p]["elem"]["frequency"] = min(ph[p]["frequency"], np.float32(0.5))
    want, woo = _run(ctx.plan(elems, offs, vp))
    got, goo = _run(ctx.plan_phonemes(ph, offs, st, vp, utt_storage=utt_storage))
    assert np.array_equal(goo, woo)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_phoneme_input_errors(ctx):
    v = g.voices.generic()
    offs = np.array([0, 2], np.uint32)
    vp = np.zeros(1, g.VOICE_DT)
    vp[:] = v.params(0)
    cf = np.array([v.center_frequency], np.float32)
    with pytest.raises(GrailError):                      # id 5 has no entry in a two-sound storage
        ctx.plan_phonemes(np.array([3, 5], np.uint8), offs, v.storage(), vp, center_frequency=cf)
    with pytest.raises(GrailError):                      # storage index out of range
        ctx.plan_phonemes(np.array([3, 4], np.uint8), offs, v.storage(), vp, center_frequency=cf,
                          utt_storage=np.array([1], np.uint32))
    ph = np.zeros(2, PHONEME_ELEM_DT)
    ph["phoneme"] = [3, 4]
    ph["length"] = [0.5, np.inf]
    with pytest.raises(GrailError):                      # non-finite length
        ctx.plan_phonemes(ph, offs, v.storage(), vp)
    # an empty batch is fine
    p = ctx.plan_phonemes(np.zeros(0, np.uint8), np.array([0], np.uint32), v.storage(), np.zeros(0, g.VOICE_DT),
                          center_frequency=np.zeros(0, np.float32))
    assert p.total_samples == 0
    p.close()


def test_text_to_audio_without_host_records(ctx):
    """text -> (native batched Transcriber) -> phoneme ids -> (device Intonator + Selector) -> the path, against the
    reference-shaped chain text.transcribe(..).intonate(..).select(..) expanded on the host (examples/cli.rs:175-184)"""
    from grail_rs_b200 import text as T
    v = g.voices.generic()
    language = T.generic_language()
    texts = ["a pie i oui e a", "", "ee", "Oui oui", "p"]
    ids, offs = T.transcribe_batch(texts, language)
    vp = np.zeros(len(texts), g.VOICE_DT)
    vp[:] = v.params(0)
    vp["jitter_seed"] = np.arange(len(texts))
    cf = np.full(len(texts), v.center_frequency, np.float32)
    got, goo = _run(ctx.plan_phonemes(ids, offs, v.storage(), vp, center_frequency=cf))
    host = [g.pack_sequence(list(T.transcribe(t, language).intonate(language, v).select(v))) for t in texts]
    elems = np.concatenate(host)
    hoffs = np.concatenate([[0], np.cumsum([len(h) for h in host])]).astype(np.uint32)
    assert np.array_equal(hoffs, offs)
    want, woo = _run(ctx.plan(elems, hoffs, vp))
    assert np.array_equal(goo, woo)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
