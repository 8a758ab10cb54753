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
