"""ctypes front-end of the CPU oracle (oracle/grail_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

PARITY UNPINNED (see the header of grail_oracle.c): the reference has no golden vectors
for the hot path; the known answers checked in tests/ come from the survey's independent
probe.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NF = 8

# numpy mirrors of the C structs (layout identical to include/grail_cuda.h, declared independently)
ELEM_DT = np.dtype([
    ("frequency", "<f4"),
    ("formant_freq", "<f4", (NF,)), ("formant_bw", "<f4", (NF,)), ("formant_smooth", "<f4", (NF,)),
    ("formant_breath", "<f4", (NF,)), ("formant_turb", "<f4", (NF,)), ("formant_amp", "<f4", (NF,)),
])
SEQ_ELEM_DT = np.dtype([("has_elem", "<u4"), ("elem", ELEM_DT), ("length", "<f4"), ("blend_length", "<f4")])
VOICE_DT = np.dtype([
    ("sample_rate", "<f4"), ("jitter_frequency", "<f4"), ("jitter_delta_frequency", "<f4"),
    ("jitter_delta_formant_frequency", "<f4"), ("jitter_delta_amplitude", "<f4"),
    ("jitter_seed", "<u4"), ("synth_seed", "<u4"),
])
assert ELEM_DT.itemsize == 196 and SEQ_ELEM_DT.itemsize == 208 and VOICE_DT.itemsize == 28


class _Trace(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "time", "alpha", "jitter_phase", "frequency", "carrier_phase", "phoneme_index",
        "state_a", "state_b", "state_c")]


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc, strict f32)."""
    if force:
        subprocess.check_call(["make", "-C", HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE, "all"], stdout=subprocess.DEVNULL)


_libs: dict[str, C.CDLL] = {}


def lib(o3: bool = False) -> C.CDLL:
    name = "libgrail_oracle_o3.so" if o3 else "libgrail_oracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(HERE, name)
    src = os.path.join(HERE, "grail_oracle.c")
    if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
        build()
    L = C.CDLL(path)
    L.grail_oracle_random_f32.restype = C.c_float
    L.grail_oracle_random_f32.argtypes = [C.POINTER(C.c_uint32)]
    L.grail_oracle_tan_approx.restype = C.c_float
    L.grail_oracle_tan_approx.argtypes = [C.c_float]
    L.grail_oracle_exp_approx.restype = C.c_float
    L.grail_oracle_exp_approx.argtypes = [C.c_float]
    L.grail_oracle_synthesize.restype = C.c_uint64
    L.grail_oracle_synthesize.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.c_void_p, C.c_void_p]
    L.grail_oracle_count_samples.restype = C.c_uint64
    L.grail_oracle_count_samples.argtypes = [C.c_void_p, C.c_uint32, C.c_float]
    L.grail_oracle_synthesize_batch.restype = C.c_int
    L.grail_oracle_synthesize_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_int]
    L.grail_oracle_fnv.restype = C.c_uint32
    L.grail_oracle_fnv.argtypes = [C.c_void_p, C.c_uint64]
    L.grail_oracle_new_phoneme.restype = None
    L.grail_oracle_new_phoneme.argtypes = [C.c_void_p] * 7
    L.grail_oracle_elem_resample.restype = None
    L.grail_oracle_elem_resample.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    L.grail_oracle_copy_with_frequency.restype = None
    L.grail_oracle_copy_with_frequency.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
    L.grail_oracle_jitter_init_states.restype = None
    L.grail_oracle_jitter_init_states.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    L.grail_oracle_sizeof_seq_elem.restype = C.c_uint32
    L.grail_oracle_sizeof_voice_params.restype = C.c_uint32
    assert L.grail_oracle_sizeof_seq_elem() == SEQ_ELEM_DT.itemsize
    assert L.grail_oracle_sizeof_voice_params() == VOICE_DT.itemsize
    _libs[name] = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------
# L0
# ---------------------------------------------------------------------------------------------
def lcg_states(seed: int, n: int) -> np.ndarray:
    """first n LCG states after `seed` (src/lib.rs:40)"""
    out = np.empty(n, np.uint32)
    s = seed & 0xFFFFFFFF
    for i in range(n):
        s = (s * 16807 + 1) & 0xFFFFFFFF
        out[i] = s
    return out


def random_f32_seq(seed: int, n: int):
    L = lib()
    st = C.c_uint32(seed)
    vals = np.empty(n, np.float32)
    states = np.empty(n, np.uint32)
    for i in range(n):
        vals[i] = L.grail_oracle_random_f32(C.byref(st))
        states[i] = st.value
    return vals, states


def tan_approx(x: float) -> np.float32:
    return np.float32(lib().grail_oracle_tan_approx(C.c_float(x)))


def exp_approx(x: float) -> np.float32:
    return np.float32(lib().grail_oracle_exp_approx(C.c_float(x)))


# ---------------------------------------------------------------------------------------------
# voice tables (src/voices/generic.rs:5-40 via src/lib.rs:381-401)
# ---------------------------------------------------------------------------------------------
_GENERIC = {
    # MKPHON argument order: freq, bw, smooth, turb, breath, amp (src/voices/mod.rs:7-14)
    "a": ([910.0, 1271.0, 2851.0, 3213.0, 1200.0, 2000.0, 3000.0, 4000.0],
          [60.0, 160.0, 180.0, 200.0, 100.0, 100.0, 100.0, 100.0],
          [1600.0] * 8,
          [0.2, 0.2, 0.1, 0.0, 0.0, 0.0, 0.0, 0.0],
          [0.5, 0.2, 0.05, 0.0, 0.0, 0.0, 0.0, 0.0],
          [0.3, 0.3, 0.2, 0.1, 0.0, 0.0, 0.0, 0.0]),
    "e": ([910.0, 1871.0, 2851.0, 3213.0, 1200.0, 2000.0, 3000.0, 4000.0],
          [80.0, 180.0, 180.0, 200.0, 100.0, 100.0, 100.0, 100.0],
          [1600.0] * 8,
          [0.2, 0.4, 0.4, 0.4, 0.4, 0.4, 0.4, 0.4],
          [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.1, 0.1],
          [0.5, 0.4, 0.3, 0.2, 0.0, 0.0, 0.0, 0.0]),
}


def new_phoneme(freq, bw, smooth, turb, breath, amp) -> np.ndarray:
    out = np.zeros((), ELEM_DT)
    arrs = [np.asarray(x, np.float32).copy() for x in (freq, bw, smooth, turb, breath, amp)]
    lib().grail_oracle_new_phoneme(*[_p(a) for a in arrs], out.ctypes.data_as(C.c_void_p))
    return out


def resample(elem: np.ndarray, old_rate: float, new_rate: float) -> np.ndarray:
    src = np.array(elem, ELEM_DT)
    out = np.zeros((), ELEM_DT)
    lib().grail_oracle_elem_resample(_p(src), C.c_float(old_rate), C.c_float(new_rate),
                                     out.ctypes.data_as(C.c_void_p))
    return out


def copy_with_frequency(elem: np.ndarray, frequency) -> np.ndarray:
    src = np.array(elem, ELEM_DT)
    out = np.zeros((), ELEM_DT)
    lib().grail_oracle_copy_with_frequency(_p(src), C.c_float(frequency), out.ctypes.data_as(C.c_void_p))
    return out


def generic_voice(sample_rate: float = 44100.0) -> dict:
    """voices::generic() (src/voices/generic.rs:5-40); for other rates the voice is rebuilt by hand
    the way SURVEY.md section 5 describes (elem.resample(44100, R), scalars divided by R)."""
    f32 = np.float32
    ph = {k: new_phoneme(*v) for k, v in _GENERIC.items()}
    R = f32(sample_rate)
    if float(R) != 44100.0:
        ph = {k: resample(v, 44100.0, float(R)) for k, v in ph.items()}
    return {
        "sample_rate": R,
        "phonemes": ph,
        "center_frequency": f32(120.0) / R,
        "jitter_frequency": f32(16.0) / R,
        "jitter_delta_frequency": f32(6.0) / R,
        "jitter_delta_formant_frequency": f32(6.0) / R,
        "jitter_delta_amplitude": f32(0.2),
    }


def voice_params(voice: dict, jitter_seed: int = 0, synth_seed: int = 0) -> np.ndarray:
    v = np.zeros((), VOICE_DT)
    for k in ("sample_rate", "jitter_frequency", "jitter_delta_frequency",
              "jitter_delta_formant_frequency", "jitter_delta_amplitude"):
        v[k] = voice[k]
    v["jitter_seed"] = jitter_seed
    v["synth_seed"] = synth_seed
    return v


# Phoneme ids (enum order of src/lib.rs:632-649 + make_phonemes!(A, E) :686-689)
SILENCE, STOP, GLIDE, A, E = range(5)


def select(phonemes, voice: dict, length: float = 0.5, blend_length: float = 0.5) -> np.ndarray:
    """Intonator (constants, src/lib.rs:1068-1073) + Selector (src/lib.rs:990-1005)."""
    out = np.zeros(len(phonemes), SEQ_ELEM_DT)
    for i, p in enumerate(phonemes):
        out[i]["length"] = length
        out[i]["blend_length"] = blend_length
        if p == A or p == E:
            out[i]["has_elem"] = 1
            out[i]["elem"] = copy_with_frequency(voice["phonemes"]["a" if p == A else "e"],
                                                 voice["center_frequency"])
    return out


# ---------------------------------------------------------------------------------------------
# the chain
# ---------------------------------------------------------------------------------------------
TRACE_FIELDS = {"time": np.float32, "alpha": np.float32, "jitter_phase": np.float32, "frequency": np.float32,
                "carrier_phase": np.float32, "phoneme_index": np.uint32,
                "state_a": np.float32, "state_b": np.float32, "state_c": np.float32}


def count_samples(elems: np.ndarray, sample_rate: float) -> int:
    e = np.ascontiguousarray(elems, SEQ_ELEM_DT)
    return int(lib().grail_oracle_count_samples(_p(e), len(e), C.c_float(sample_rate)))


def synthesize(elems: np.ndarray, vparams: np.ndarray, trace: bool = False, o3: bool = False):
    """returns (samples f32[n], trace dict | None, final_states u32[30])"""
    e = np.ascontiguousarray(elems, SEQ_ELEM_DT)
    v = np.ascontiguousarray(vparams, VOICE_DT).reshape(())
    n = count_samples(e, float(v["sample_rate"]))
    out = np.empty(n, np.float32)
    tr = None
    tstruct = None
    if trace:
        tr = {}
        tstruct = _Trace()
        for k, dt in TRACE_FIELDS.items():
            shape = (n, NF) if k.startswith("state_") else (n,)
            tr[k] = np.zeros(shape, dt)
            setattr(tstruct, k, tr[k].ctypes.data)
    fin = np.zeros(30, np.uint32)
    got = lib(o3).grail_oracle_synthesize(_p(e), len(e), _p(v), _p(out), n,
                                          C.byref(tstruct) if tstruct is not None else None, _p(fin))
    assert got == n, (got, n)
    return out, tr, fin


def synthesize_batch(elems: np.ndarray, utt_offsets: np.ndarray, voices: np.ndarray,
                     out_offsets: np.ndarray | None = None, n_threads: int = 1, o3: bool = True,
                     out: np.ndarray | None = None):
    """one utterance per task on n_threads host threads; returns (out, out_offsets, counts)"""
    e = np.ascontiguousarray(elems, SEQ_ELEM_DT)
    offs = np.ascontiguousarray(utt_offsets, np.uint32)
    v = np.ascontiguousarray(voices, VOICE_DT)
    n = len(offs) - 1
    assert len(v) == n
    if out_offsets is None:
        counts = np.array([count_samples(e[offs[u]:offs[u + 1]], float(v[u]["sample_rate"])) for u in range(n)],
                          np.uint64)
        out_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    oo = np.ascontiguousarray(out_offsets, np.uint64)
    if out is None:
        out = np.empty(int(oo[-1]), np.float32)
    counts = np.zeros(n, np.uint64)
    rc = lib(o3).grail_oracle_synthesize_batch(_p(e), _p(offs), _p(v), n, _p(out), _p(oo), _p(counts), n_threads)
    assert rc == 0
    return out, oo, counts


def fnv(x: np.ndarray) -> int:
    x = np.ascontiguousarray(x, np.float32)
    return int(lib().grail_oracle_fnv(_p(x), x.size))


def jitter_init_states(seed: int):
    st = np.zeros(3, np.uint32)
    cn = np.zeros(2, np.float32)
    lib().grail_oracle_jitter_init_states(seed, _p(st), _p(cn))
    return st, cn
