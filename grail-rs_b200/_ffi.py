"""ctypes binding of libgrail_cuda.so (include/grail_cuda.h).  No CPU fallback: if the library is
missing the import fails, and if no device is present every compute call raises GrailError."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GRAIL_CUDA_LIB") or os.path.join(HERE, "libgrail_cuda.so")   # (the override loads experimental builds of the same library)
NF = 8

ELEM_DT = np.dtype([
    ("frequency", "<f4"),
    ("formant_freq", "<f4", (NF,)), ("formant_bw", "<f4", (NF,)), ("formant_smooth", "<f4", (NF,)),
    ("formant_breath", "<f4", (NF,)), ("formant_turb", "<f4", (NF,)), ("formant_amp", "<f4", (NF,)),
])
SEQ_ELEM_DT = np.dtype([("has_elem", "<u4"), ("elem", ELEM_DT), ("length", "<f4"), ("blend_length", "<f4")])
PHONEME_ELEM_DT = np.dtype([("phoneme", "<u4"), ("length", "<f4"), ("blend_length", "<f4"), ("frequency", "<f4")])   # 16 B
VOICE_DT = np.dtype([
    ("sample_rate", "<f4"), ("jitter_frequency", "<f4"), ("jitter_delta_frequency", "<f4"),
    ("jitter_delta_formant_frequency", "<f4"), ("jitter_delta_amplitude", "<f4"),
    ("jitter_seed", "<u4"), ("synth_seed", "<u4"),
])
assert ELEM_DT.itemsize == 196 and SEQ_ELEM_DT.itemsize == 208 and VOICE_DT.itemsize == 28

OK, ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_OOM, ERR_COUNT_MISMATCH, ERR_UNSUPPORTED = range(7)
F32, I16 = 0, 1


class Timings(C.Structure):
    _fields_ = [("schedule_ms", C.c_float), ("frequency_ms", C.c_float), ("phase_ms", C.c_float),
                ("formant_ms", C.c_float), ("total_ms", C.c_float), ("n_launches", C.c_uint32)]


class GrailError(RuntimeError):
    def __init__(self, status: int, message: str = ""):
        self.status = status
        super().__init__(f"grail_cuda status {status} ({_status_string(status)}){': ' + message if message else ''}")


_lib = None

# every symbol include/grail_cuda.h and include/grail_cuda_debug.h declare
EXPORTS = [
    "grail_cuda_abi_version", "grail_cuda_device_count", "grail_cuda_status_string", "grail_cuda_create",
    "grail_cuda_destroy", "grail_cuda_last_error", "grail_cuda_stream_handle", "grail_cuda_synchronize",
    "grail_cuda_set_option", "grail_cuda_host_alloc", "grail_cuda_host_free", "grail_cuda_count_samples",
    "grail_cuda_transcribe_batch", "grail_cuda_synthesize_batch", "grail_cuda_synthesize_batch_i16", "grail_cuda_plan_create", "grail_cuda_plan_create_phoneme_elems",
    "grail_cuda_plan_create_phonemes", "grail_cuda_plan_destroy",
    "grail_cuda_plan_join", "grail_cuda_plan_total_samples", "grail_cuda_plan_out_offsets", "grail_cuda_plan_launch", "grail_cuda_plan_launch_interleaved",
    "grail_cuda_plan_device_output", "grail_cuda_plan_read_output", "grail_cuda_plan_timings",
    "grail_cuda_plan_read_intermediates", "grail_cuda_plan_phase_scan_stats", "grail_cuda_plan_phase_stats", "grail_cuda_stream_new", "grail_cuda_stream_push",
    "grail_cuda_stream_finish", "grail_cuda_stream_pull", "grail_cuda_stream_free", "grail_cuda_probe_fp32_peak",
    "grail_cuda_copy_segments", "grail_cuda_streams_pull",
    "grail_cuda_debug_clock_desc", "grail_cuda_debug_clock_asc", "grail_cuda_debug_lcg_jump",
    "grail_cuda_debug_jitter_index", "grail_cuda_debug_div_check",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "grail_cuda_abi_version": (i32, []),
        "grail_cuda_device_count": (i32, []),
        "grail_cuda_status_string": (C.c_char_p, [i32]),
        "grail_cuda_create": (i32, [i32, C.POINTER(vp)]),
        "grail_cuda_destroy": (None, [vp]),
        "grail_cuda_last_error": (C.c_char_p, [vp]),
        "grail_cuda_stream_handle": (vp, [vp]),
        "grail_cuda_synchronize": (i32, [vp]),
        "grail_cuda_set_option": (i32, [vp, C.c_char_p, C.c_double]),
        "grail_cuda_host_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "grail_cuda_host_free": (None, [vp, vp]),
        "grail_cuda_count_samples": (i32, [vp, vp, vp, u32, vp]),
        "grail_cuda_transcribe_batch": (i32, [vp, vp, u32, vp, u32, i32, i32, vp, u64, vp, i32]),
        "grail_cuda_synthesize_batch": (i32, [vp, vp, vp, vp, u32, vp, vp, i32]),
        "grail_cuda_synthesize_batch_i16": (i32, [vp, vp, vp, vp, u32, vp, vp, i32]),
        "grail_cuda_plan_create": (i32, [vp, vp, vp, vp, u32, C.POINTER(vp)]),
        "grail_cuda_plan_create_phoneme_elems": (i32, [vp, vp, vp, vp, u32, u32, vp, vp, u32, C.POINTER(vp)]),
        "grail_cuda_plan_create_phonemes": (i32, [vp, vp, vp, vp, vp, u32, u32, vp, vp, u32, C.POINTER(vp)]),
        "grail_cuda_plan_destroy": (None, [vp]),
        "grail_cuda_plan_join": (i32, [vp]),
        "grail_cuda_plan_total_samples": (u64, [vp]),
        "grail_cuda_plan_out_offsets": (i32, [vp, vp]),
        "grail_cuda_plan_launch": (i32, [vp, vp, i32]),
        "grail_cuda_plan_launch_interleaved": (i32, [vp, vp, i32, u32]),
        "grail_cuda_plan_device_output": (i32, [vp, i32, C.POINTER(vp)]),
        "grail_cuda_plan_read_output": (i32, [vp, i32, vp]),
        "grail_cuda_plan_timings": (i32, [vp, C.POINTER(Timings)]),
        "grail_cuda_plan_read_intermediates": (i32, [vp, vp, vp, vp]),
        "grail_cuda_plan_phase_scan_stats": (i32, [vp, vp]),
        "grail_cuda_plan_phase_stats": (i32, [vp, vp]),
        "grail_cuda_stream_new": (i32, [vp, vp, C.POINTER(vp)]),
        "grail_cuda_stream_push": (i32, [vp, vp, u32]),
        "grail_cuda_stream_finish": (i32, [vp]),
        "grail_cuda_stream_pull": (i32, [vp, vp, u64, C.POINTER(u64)]),
        "grail_cuda_streams_pull": (i32, [vp, u32, vp, vp, vp]),
        "grail_cuda_stream_free": (None, [vp]),
        "grail_cuda_copy_segments": (i32, [vp, vp, vp, vp, vp, vp, u64, u32]),
        "grail_cuda_probe_fp32_peak": (i32, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "grail_cuda_debug_clock_desc": (None, [C.c_float, C.c_float, u64, C.POINTER(C.c_float), C.POINTER(u64), C.POINTER(i32)]),
        "grail_cuda_debug_clock_asc": (None, [C.c_float, C.c_float, u64, C.POINTER(C.c_float), C.POINTER(u64), C.POINTER(i32)]),
        "grail_cuda_debug_lcg_jump": (u32, [u32, u64]),
        "grail_cuda_debug_jitter_index": (u64, [i32, i32, i32, u64]),
        "grail_cuda_debug_div_check": (i32, [vp, u32, u64, C.POINTER(u64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class TranscriptionRuleC(C.Structure):
    _fields_ = [("string", C.c_char_p), ("phonemes", C.POINTER(C.c_uint8)), ("n_phonemes", C.c_uint32)]


def _status_string(status: int) -> str:
    try:
        return lib().grail_cuda_status_string(status).decode()
    except Exception:
        return "?"


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
