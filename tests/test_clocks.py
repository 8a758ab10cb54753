"""Exact closed-form clocks (grail-rs_b200/csrc/grail_common.cuh) against literal f32 loops, and the exact
sample counts of the C ABI against the oracle's Sequencer.  Host only: no device needed."""
import ctypes as C

import numpy as np
import pytest

import grail_rs_b200 as g

f32 = np.float32
L = g._ffi.lib()


def desc(x, d, max_steps):
    xo, st, sk = C.c_float(), C.c_uint64(), C.c_int()
    L.grail_cuda_debug_clock_desc(C.c_float(x), C.c_float(d), max_steps, C.byref(xo), C.byref(st), C.byref(sk))
    return f32(xo.value), st.value, sk.value


def asc(x, d, max_steps):
    xo, st, sk = C.c_float(), C.c_uint64(), C.c_int()
    L.grail_cuda_debug_clock_asc(C.c_float(x), C.c_float(d), max_steps, C.byref(xo), C.byref(st), C.byref(sk))
    return f32(xo.value), st.value, sk.value


def literal_desc(x, d, max_steps):
    x, d = f32(x), f32(d)
    n = 0
    while n < max_steps:
        y = f32(x - d)
        if y == x:
            return x, n, 1
        x = y
        n += 1
        if x < 0:
            break
    return x, n, 0


def literal_asc(x, d, max_steps):
    x, d = f32(x), f32(d)
    n = 0
    while n < max_steps:
        y = f32(x + d)
        if y == x:
            return x, n, 1
        x = y
        n += 1
        if x > 1:
            break
    return x, n, 0


RATES = [8000.0, 16000.0, 16384.0, 22050.0, 32768.0, 44100.0, 48000.0, 96000.0]


@pytest.mark.parametrize("rate", RATES)
def test_desc_full_run_matches_literal(rate):
    rng = np.random.default_rng(int(rate))
    dt = f32(1.0) / f32(rate)
    for length in [0.5, 0.02, 1.0, 2.0, 0.123456, 1e-5, 0.0] + list(rng.uniform(0.02, 2.0, 6)):
        x0 = f32(f32(-dt * f32(rng.uniform(0, 1))) + f32(length))
        got = desc(x0, dt, 1 << 40)
        want = literal_desc(x0, dt, 1 << 40)
        assert got[1] == want[1] and got[0].view(np.uint32) == want[0].view(np.uint32), (rate, length)


@pytest.mark.parametrize("rate", [16000.0, 44100.0, 48000.0])
def test_desc_partial_runs_match_literal(rate):
    rng = np.random.default_rng(7)
    dt = f32(1.0) / f32(rate)
    x0 = f32(0.5) - dt
    # walk the literal loop once, compare the closed form at random step counts
    xs = [x0]
    x = x0
    while x >= 0:
        x = f32(x - dt)
        xs.append(x)
    for k in [0, 1, 2, 3, 100, len(xs) - 2, len(xs) - 1] + list(rng.integers(0, len(xs) - 1, 200)):
        got = desc(x0, dt, int(k))
        assert got[1] == k and got[0].view(np.uint32) == xs[int(k)].view(np.uint32), k


def test_asc_matches_literal_with_wraps():
    rng = np.random.default_rng(3)
    for rate, hz in [(44100.0, 16.0), (16000.0, 16.0), (48000.0, 31.7), (44100.0, 8.0), (22050.0, 16.0)]:
        inc = f32(hz) / f32(rate)
        ph = f32(0.0)
        for _ in range(12):   # 12 consecutive periods, phase carried through the wrap like src/lib.rs:245-246
            got = asc(ph, inc, 1 << 40)
            want = literal_asc(ph, inc, 1 << 40)
            assert got[1] == want[1] and got[0].view(np.uint32) == want[0].view(np.uint32)
            assert got[0] > 1
            # partial runs inside the period
            for k in rng.integers(0, want[1], 20):
                a = asc(ph, inc, int(k))
                b = literal_asc(ph, inc, int(k))
                assert a[1] == b[1] == k and a[0].view(np.uint32) == b[0].view(np.uint32)
            ph = f32(got[0] - f32(1.0))


def test_clock_edge_cases():
    # power-of-two increment: exact ties in one binade (round-half-even) and a stuck clock above it
    d = f32(2.0 ** -14)
    for x0 in [f32(3.0), f32(0.75), f32(1000.0)]:
        got, want = desc(x0, d, 50000), literal_desc(x0, d, 50000)
        assert got[1] == want[1] and got[0].view(np.uint32) == want[0].view(np.uint32)
    x, steps, stuck = desc(f32(4096.0), d, 10)    # d < ulp/2: the reference would loop forever
    assert stuck == 1 and steps == 0
    # zero increment never moves
    assert asc(f32(0.0), f32(0.0), 100)[2] == 1
    # increments larger than the value, denormal start
    for x0, dd in [(f32(1e-7), f32(1e-3)), (f32(1e-42), f32(1e-5)), (f32(0.0), f32(0.3))]:
        assert desc(x0, dd, 1 << 30)[:2] == literal_desc(x0, dd, 1 << 30)[:2]
        a, b = asc(x0, dd, 1 << 30), literal_asc(x0, dd, 1 << 30)
        assert a[1] == b[1] and a[0].view(np.uint32) == b[0].view(np.uint32)
    # odd-mantissa ties: d = 3 * 2^-25 against x in [0.5, 1) (grid 2^-24)
    d = f32(3 * 2.0 ** -25)
    for x0 in [f32(0.5 + 2.0 ** -24), f32(0.5 + 2.0 ** -23), f32(0.999)]:
        a, b = desc(x0, d, 1 << 30), literal_desc(x0, d, 1 << 30)
        assert a[1] == b[1] and a[0].view(np.uint32) == b[0].view(np.uint32)
        a, b = asc(f32(x0 / 4), d, 1 << 30), literal_asc(f32(x0 / 4), d, 1 << 30)
        assert a[1] == b[1] and a[0].view(np.uint32) == b[0].view(np.uint32)


def test_lcg_jump_and_draw_indices(oracle):
    states = oracle.lcg_states(12345, 3000)
    for n in [0, 1, 2, 17, 18, 34, 35, 999, 2999]:
        want = 12345 if n == 0 else int(states[n - 1])
        assert L.grail_cuda_debug_lcg_jump(12345, n) == want
    assert "%08x" % L.grail_cuda_debug_lcg_jump(0, 26457161) == "fa9c2461"
    # SURVEY Appendix C table
    J = L.grail_cuda_debug_jitter_index
    assert [J(-1, 0, 0, w) for w in range(3)] == [1, 2, 3] and [J(-1, 1, 0, w) for w in range(3)] == [2, 3, 4]
    for i in range(8):
        assert J(0, 0, i, 0) == 3 + 2 * i and J(0, 1, i, 0) == 4 + 2 * i
        assert J(0, 0, i, 1) == 4 + 2 * i and J(0, 1, i, 1) == 18 + i + 1
        assert J(0, 0, i, 2) == 18 + i + 1 and J(0, 1, i, 2) == 18 + 8 + i + 1
        assert J(1, 0, i, 0) == 19 + 2 * i and J(1, 1, i, 0) == 20 + 2 * i
        assert J(1, 0, i, 1) == 20 + 2 * i and J(1, 1, i, 1) == 34 + i + 1
        assert J(1, 0, i, 5) == 34 + 8 * 3 + i + 1 and J(1, 1, i, 5) == 34 + 8 * 4 + i + 1


def test_count_samples_matches_oracle(oracle):
    rng = np.random.default_rng(11)
    elems, offs, voices, want = [], [0], [], []
    for rate in [8000.0, 16000.0, 22050.0, 44100.0, 48000.0, 96000.0]:
        v = oracle.generic_voice(rate)
        for trial in range(12):
            n = int(rng.integers(0, 7))
            e = oracle.select([int(x) for x in rng.integers(0, 5, n)], v)
            if trial % 3:
                e["length"] = rng.uniform(0.02, 2.0, n).astype(f32)
            if trial % 4 == 3 and n:
                e["length"][0] = f32(1e-6)      # shorter than one sample
            elems.append(e)
            offs.append(offs[-1] + n)
            voices.append(oracle.voice_params(v, trial))
            want.append(oracle.count_samples(e, rate))
    got = g.count_samples(np.concatenate(elems), np.array(offs, np.uint32), np.array(voices))
    assert got.tolist() == want


def test_count_samples_rejects_bad_input(oracle, voice):
    e = oracle.select([3], voice)
    bad = oracle.voice_params(voice)
    bad["sample_rate"] = 0.0
    with pytest.raises(g.GrailError) as ei:
        g.count_samples(e, np.array([0, 1], np.uint32), bad.reshape(1))
    assert ei.value.status == g._ffi.ERR_INVALID_ARG
    bad = oracle.voice_params(voice)
    bad["jitter_frequency"] = 0.6
    with pytest.raises(g.GrailError) as ei:
        g.count_samples(e, np.array([0, 1], np.uint32), bad.reshape(1))
    assert ei.value.status == g._ffi.ERR_UNSUPPORTED
