"""Regenerates tests/golden/oracle_waveforms.npz: oracle outputs (audio, F_t, carrier phase as raw f32 bits) for a few
small inputs, so that the GPU parity tests also have a committed, oracle-independent-at-test-time anchor and the CPU
suite can detect drift of the oracle itself (compiler, flags).  Run from the repo root: python tests/golden/make_golden.py
Provenance: oracle/grail_oracle.c (strict f32 restatement of reference src/lib.rs), gcc -O2 -ffp-contract=off, x86-64."""
import os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grail_rs_b200 as g
from grail_rs_b200 import workloads as W
from oracle import oracle as O

CASES = {
    # name: (phoneme lists, sample rate, jitter seed)
    "sil_a_44100": ([[0, 3]], 44100.0, 0),            # text "a" (SURVEY 8d config 1 smallest KAT: 44 095 samples)
    "a_e_sil_a_16000": ([[3, 4, 0, 3]], 16000.0, 7),  # a resampled voice, a silence in the middle
    "e_22050_seed12345": ([[4]], 22050.0, 12345),
}


def build(name):
    lists, rate, seed = CASES[name]
    v = g.voices.generic()
    if rate != 44100.0:
        v = g.voices.at_sample_rate(v, rate)
    elems, offs, vp = W.from_phonemes(lists, v, [seed])
    return elems, offs, vp


def main():
    O.lib()
    out = {}
    for name in CASES:
        elems, offs, vp = build(name)
        audio, tr, _ = O.synthesize(elems, vp[0], trace=True)
        out[name + "/audio"] = audio.view(np.uint32)
        out[name + "/frequency"] = tr["frequency"].view(np.uint32)
        out[name + "/carrier_phase"] = tr["carrier_phase"].view(np.uint32)
        print(name, len(audio), "samples, fnv %08x" % O.fnv(audio))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_waveforms.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
