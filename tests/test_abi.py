"""The C-ABI library loads and exports every symbol include/*.h declares; host-side behaviour without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import grail_rs_b200 as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("grail_cuda.h", "grail_cuda_debug.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(grail_cuda_\w+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported():
    L = g._ffi.lib()
    decl = declared_symbols()
    assert len(decl) >= 30
    for name in sorted(decl):
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert decl == set(g._ffi.EXPORTS)


def test_struct_layout_matches_header():
    assert g.ELEM_DT.itemsize == 196 and g.SEQ_ELEM_DT.itemsize == 208 and g.VOICE_DT.itemsize == 28
    assert g.SEQ_ELEM_DT.fields["length"][1] == 200 and g.SEQ_ELEM_DT.fields["blend_length"][1] == 204
    assert g.SEQ_ELEM_DT.fields["elem"][1] == 4
    assert g._ffi.lib().grail_cuda_abi_version() == 4


def test_status_strings():
    L = g._ffi.lib()
    assert L.grail_cuda_status_string(0) == b"ok"
    assert b"no CPU path" in L.grail_cuda_status_string(g._ffi.ERR_NO_DEVICE)


@pytest.mark.skipif(g._ffi.lib().grail_cuda_device_count() > 0, reason="checks the no-device behaviour")
def test_no_device_is_an_error_not_a_fallback():
    with pytest.raises(g.GrailError) as ei:
        g.Context(0)
    assert ei.value.status == g._ffi.ERR_NO_DEVICE
    v = g.voices.generic()
    chain = g.sequence([g.SequenceElem.new(None, 0.5, 0.5)], v).jitter(0, v).synthesize()
    with pytest.raises(g.GrailError):
        next(chain)


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under the product package or include/ may reference it"""
    bad = []
    for base in ("grail-rs_b200", "grail_rs_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".rs", ".toml")) or fn == "Makefile":
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|grail_oracle", txt):
                        if "Nothing here comes from or calls oracle/" in txt and len(re.findall(r"oracle", txt)) == 1:
                            continue
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_public_headers_are_plain_c():
    """include/*.h is what a cgo / bindgen / ctypes consumer reads: it must compile as C99, with no C++ or CUDA types"""
    import subprocess, tempfile
    inc = os.path.join(ROOT, "include")
    src = ('#include "grail_cuda.h"\n#include "grail_cuda_debug.h"\n'
           "int main(void) { grail_seq_elem e; grail_phoneme_elem p; grail_transcription_rule r; grail_voice_params v;\n"
           "  (void)e; (void)p; (void)r; (void)v; return sizeof(grail_seq_elem) == 208 && sizeof(grail_phoneme_elem) == 16 ? 0 : 1; }\n")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "hdr.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "hdr")
        subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + inc, c, "-o", exe])
        assert subprocess.call([exe]) == 0
