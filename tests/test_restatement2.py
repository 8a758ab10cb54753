"""The second, independent CPU restatement (oracle/restatement2.py: NumPy float32, written from src/lib.rs as the
reference's chain of pull iterators) against the C oracle and the CUDA path.

tests/golden/restatement2_kat.json holds restatement2's answers for 16 inputs (the survey's known-answer phoneme
lists, config-4 random voices and config-5 rebuilt voices at 16 / 22.05 / 48 kHz); scripts/cross_check_restatements.py
regenerates it.  Parity stays UNPINNED by the reference itself (Rust, not compilable here): these tests prove that two
restatements written separately agree bit for bit, not that either equals the crate's output."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import cross_check_restatements as X   # noqa: E402

from oracle import restatement2 as R2   # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "restatement2_kat.json")))["cases"]
CASES = {name: (elems, vp) for name, elems, vp in X.cases()}


def test_case_list_matches_golden_file():
    assert sorted(CASES) == sorted(GOLD)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_c_oracle_equals_restatement2_golden(oracle, name):
    """the C oracle reproduces restatement2's committed answers: count, FNV of audio / F_t / carrier phase, samples"""
    elems, vp = CASES[name]
    audio, tr, _ = oracle.synthesize(elems, vp, trace=True)
    gold = GOLD[name]
    assert len(audio) == gold["n"]
    assert f"{oracle.fnv(audio):08x}" == gold["fnv"]
    assert f"{oracle.fnv(tr['frequency']):08x}" == gold["fnv_frequency"]
    assert f"{oracle.fnv(tr['carrier_phase']):08x}" == gold["fnv_carrier_phase"]
    for i, b in gold["samples"].items():
        assert f"{int(audio[int(i)].view(np.uint32)):08x}" == b


@pytest.mark.parametrize("name", ["config4_utt1092_16000hz", "config5_16000hz"])
def test_restatement2_live_against_c_oracle(oracle, name):
    """restatement2 itself, run here (2-3 s per case), bit for bit against the C oracle: audio, F_t, carrier phase"""
    elems, vp = CASES[name]
    got, tr2 = R2.synthesize_records(elems, vp, trace=True)
    want, tr1, _ = oracle.synthesize(elems, vp, trace=True)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(tr2["frequency"].view(np.uint32), tr1["frequency"].view(np.uint32))
    assert np.array_equal(tr2["carrier_phase"].view(np.uint32), tr1["carrier_phase"].view(np.uint32))
    assert f"{R2.fnv(got):08x}" == GOLD[name]["fnv"]


def test_restatement2_voice_table_and_leaf_math_bits():
    """SURVEY Appendix B bit patterns from restatement2's own voice builder and math kernels"""
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "survey_probe_kat.json")))
    voice, ph, cf = R2.generic_voice()
    b = lambda x: f"{int(np.float32(x).view(np.uint32)):08x}"   # noqa: E731
    vb = kat["voice_bits"]
    assert b(ph["a"].arr[0][0]) == vb["a.formant_freq[0]"] and b(ph["a"].arr[0][1]) == vb["a.formant_freq[1]"]
    assert b(ph["a"].arr[1][0]) == vb["a.formant_bw[0]"] and b(ph["a"].arr[2][0]) == vb["a.formant_smooth[0]"]
    assert [b(x) for x in ph["a"].arr[5][:4]] == vb["a.formant_amp[0..3]"]
    assert [b(x) for x in ph["e"].arr[5][:4]] == vb["e.formant_amp[0..3]"]
    assert b(cf) == vb["center_frequency"] and b(voice.jitter_frequency) == vb["jitter_frequency"]
    assert b(voice.jitter_delta_frequency) == vb["jitter_delta_frequency"]
    assert b(np.float32(1.0) / voice.sample_rate) == vb["dt"]
    mb = kat["math_bits"]
    assert b(R2.tan_approx(ph["a"].arr[0][0])) == mb["tan_approx(a.formant_freq[0])"]
    assert b(R2.exp_approx(ph["a"].arr[2][0])) == mb["exp_approx(a.formant_smooth[0])"]
    r = R2.Rng(0)
    vals = [r.random_f32() for _ in range(4)]
    assert np.allclose(vals, kat["lcg_floats_from_0"], rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_cuda_against_restatement2_golden():
    """the CUDA path against restatement2's answers with no oracle in the loop: sample counts, F_t and carrier phase
    bit-exact (FNV of the taps), a few audio samples within the tolerance"""
    import grail_rs_b200 as g
    names = sorted(GOLD)
    elems = np.concatenate([CASES[n][0] for n in names])
    offs = np.concatenate([[0], np.cumsum([len(CASES[n][0]) for n in names])]).astype(np.uint32)
    vp = np.array([CASES[n][1] for n in names])
    with g.Context(0) as ctx:
        plan = ctx.plan(elems, offs, vp)
        plan.launch()
        out = plan.read_output()
        f, ph, _ = plan.read_intermediates()
        oo = plan.out_offsets
        for u, n in enumerate(names):
            gold = GOLD[n]
            a, b = int(oo[u]), int(oo[u + 1])
            assert b - a == gold["n"], n
            assert f"{R2.fnv(f[a:b]):08x}" == gold["fnv_frequency"], n
            assert f"{R2.fnv(ph[a:b]):08x}" == gold["fnv_carrier_phase"], n
            for i, bits in gold["samples"].items():
                want = np.uint32(int(bits, 16)).view(np.float32)
                assert abs(float(out[a + int(i)]) - float(want)) <= 1e-4, (n, i)
        plan.close()
