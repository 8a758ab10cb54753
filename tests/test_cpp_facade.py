"""The header-only C++ facade over the C ABI: compiles and links against libgrail_cuda.so with g++ (CPU), and on a
GPU reproduces the ctypes path exactly (same library, same call)."""
import os
import subprocess

import numpy as np
import pytest

import grail_rs_b200 as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "grail-rs_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "facade_test")


def build_exe():
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"),
                           "-L" + LIBDIR, "-l:libgrail_cuda.so", "-Wl,-rpath," + LIBDIR])


def test_facade_compiles_and_links():
    g._ffi.lib()
    build_exe()
    if g._ffi.lib().grail_cuda_device_count() == 0:
        r = subprocess.run([EXE], capture_output=True, text=True)
        assert r.returncode == 3 and "no-device" in r.stdout      # no CPU fallback


@pytest.mark.gpu
def test_facade_matches_ctypes_path():
    build_exe()
    r = subprocess.run([EXE], capture_output=True, text=True, check=True)
    n, checksum = r.stdout.split()
    e = np.zeros(3, g.SEQ_ELEM_DT)
    e["length"] = [0.05, 0.08, 0.08]
    e["blend_length"] = e["length"]
    for k in (1, 2):
        e[k]["has_elem"] = 1
        el = e[k]["elem"]
        el["frequency"] = np.float32(120.0) / np.float32(44100.0)
        el["formant_freq"] = np.array([910, 1271, 2851, 3213, 1200, 2000, 3000, 4000], np.float32) / np.float32(44100.0)
        el["formant_bw"] = np.array([60, 160, 180, 200, 100, 100, 100, 100], np.float32) / np.float32(44100.0)
        el["formant_smooth"] = np.float32(1600.0) / np.float32(44100.0)
        el["formant_breath"] = 0.2
        el["formant_turb"] = 0.1
        el["formant_amp"] = [0.4, 0.35, 0.25, 0, 0, 0, 0, 0]
    vp = np.zeros(1, g.VOICE_DT)
    vp[0] = (44100.0, np.float32(16.0) / np.float32(44100.0), np.float32(6.0) / np.float32(44100.0),
             np.float32(6.0) / np.float32(44100.0), 0.2, 7, 0)
    with g.Context(0) as ctx:
        out, _ = ctx.synthesize_batch(e, np.array([0, 3], np.uint32), vp)
    assert int(n) == len(out)
    assert abs(float(checksum) - float(np.abs(out.astype(np.float64)).sum())) < 1e-6 * max(1.0, float(checksum))


STREAM_EXE = os.path.join(ROOT, "tests", "cpp", "stream_test")


def build_stream_exe():
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", STREAM_EXE, os.path.join(ROOT, "tests", "cpp", "stream_test.cpp"),
                           "-L" + LIBDIR, "-l:libgrail_cuda.so", "-Wl,-rpath," + LIBDIR])


def test_stream_adaptor_compiles_and_links():
    g._ffi.lib()
    build_stream_exe()
    if g._ffi.lib().grail_cuda_device_count() == 0:
        r = subprocess.run([STREAM_EXE], capture_output=True, text=True)
        assert r.returncode == 3 and "no-device" in r.stdout


def _stream_source(k: int):
    """element k of stream_test.cpp's infinite source, as a packed record"""
    e = np.zeros(1, g.SEQ_ELEM_DT)
    e["length"] = 0.05
    e["blend_length"] = 0.05
    if k % 5 != 0:
        e["has_elem"] = 1
        el = e[0]["elem"]
        el["frequency"] = np.float32(120.0 if k % 2 else 150.0) / np.float32(44100.0)
        el["formant_freq"] = np.array([910, 1271, 2851, 3213, 1200, 2000, 3000, 4000], np.float32) / np.float32(44100.0)
        el["formant_bw"] = np.array([60, 160, 180, 200, 100, 100, 100, 100], np.float32) / np.float32(44100.0)
        el["formant_smooth"] = np.float32(1600.0) / np.float32(44100.0)
        el["formant_breath"] = 0.2
        el["formant_turb"] = 0.1
        el["formant_amp"] = [0.4, 0.35, 0.25, 0, 0, 0, 0, 0]
    return e


@pytest.mark.gpu
def test_stream_adaptor_infinite_upstream_matches_oracle(oracle, tmp_path):
    """examples/interactive.rs shape: infinite repeat-with source, lazily pulled windows of one audio callback (441
    frames), two channels.  Channel 0 must be the oracle's rendering of the same source prefix, and the adaptor must have
    pulled only as much upstream as the look-ahead needs."""
    build_stream_exe()
    raw = str(tmp_path / "ch0.f32")
    n_cb, frames = 120, 441
    r = subprocess.run([STREAM_EXE, str(n_cb), str(frames), "2", raw], capture_output=True, text=True, check=True)
    words = r.stdout.split()
    pulled, calls, equal = int(words[1]), int(words[3]), int(words[7])
    assert pulled == n_cb * frames and equal == 1
    got = np.fromfile(raw, np.float32)
    # 52 920 samples = 24 elements of 0.05 s (2 205 samples): the adaptor needs element 24 in progress + one look-ahead
    assert 25 <= calls <= 27, calls
    elems = np.concatenate([_stream_source(k) for k in range(calls)])
    vp = np.zeros(1, g.VOICE_DT)
    vp[0] = (44100.0, np.float32(16.0) / np.float32(44100.0), np.float32(6.0) / np.float32(44100.0),
             np.float32(6.0) / np.float32(44100.0), 0.2, 0, 0)
    want, _, _ = oracle.synthesize(elems, vp[0])
    from grail_rs_b200 import workloads as W
    st = W.parity_stats(got, want[: len(got)])
    assert st["max_abs"] <= 1e-4 and st["snr_db"] >= 90.0, st
