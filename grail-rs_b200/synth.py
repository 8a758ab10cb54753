"""Host-side mirror of the reference's operator interface for the waveform path.

Same names and argument meaning as reference src/lib.rs: SynthesisElem (:316-460), SequenceElem (:814-835),
Voice (:696-717), and the iterator verbs `.sequence(voice)` (:936-953), `.jitter(seed, voice)` (:781-801),
`.synthesize()` (:582-600).  `Sequencer` and `Jitter` are lazy descriptors; `Synthesize` drains the upstream
SequenceElems on the host, runs the per-sample work on the B200 through the C ABI, and yields f32 samples.
There is no CPU synthesis path in this package.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, replace
from typing import Iterable, Iterator, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import ELEM_DT, SEQ_ELEM_DT, VOICE_DT, GrailError, Timings, ptr

f32 = np.float32
NUM_FORMANTS = 8                 # src/lib.rs:24
DEFAULT_SAMPLE_RATE = f32(44100.0)  # src/lib.rs:21


def _arr(x) -> np.ndarray:
    a = np.asarray(x, dtype=np.float32)
    if a.shape != (NUM_FORMANTS,):
        raise ValueError(f"expected {NUM_FORMANTS} formant values, got shape {a.shape}")
    return a.copy()


@dataclass
class SynthesisElem:
    """reference src/lib.rs:316-337; all frequencies normalised to the sample rate"""
    frequency: np.float32
    formant_freq: np.ndarray
    formant_bw: np.ndarray
    formant_smooth: np.ndarray
    formant_breath: np.ndarray
    formant_turb: np.ndarray
    formant_amp: np.ndarray

    @staticmethod
    def new(sample_rate, frequency, formant_freq, formant_smooth, formant_bw, formant_breath, formant_turb,
            formant_amp) -> "SynthesisElem":
        """src/lib.rs:343-364 (note the reference's argument order: smooth before bw)"""
        return SynthesisElem(f32(frequency), _arr(formant_freq), _arr(formant_bw), _arr(formant_smooth),
                             _arr(formant_breath), _arr(formant_turb), _arr(formant_amp)).resample(1.0, sample_rate)

    @staticmethod
    def silent() -> "SynthesisElem":
        """src/lib.rs:367-377"""
        q, z = np.full(8, 0.25, f32), np.zeros(8, f32)
        return SynthesisElem(f32(0.25), q.copy(), q.copy(), q.copy(), z.copy(), z.copy(), z.copy())

    @staticmethod
    def new_phoneme(formant_freq, formant_bw, formant_smooth, formant_turb, formant_breath, formant_amp) -> "SynthesisElem":
        """src/lib.rs:381-401: amplitudes normalised by their sequential f32 sum, then resample(1, 44100)"""
        amp = _arr(formant_amp)
        s = f32(0.0)
        for v in amp:               # Array::sum is a left fold (:123)
            s = f32(s + v)
        return SynthesisElem(f32(0.0), _arr(formant_freq), _arr(formant_bw), _arr(formant_smooth), _arr(formant_breath),
                             _arr(formant_turb), (amp / s).astype(f32)).resample(1.0, DEFAULT_SAMPLE_RATE)

    def blend(self, other: "SynthesisElem", alpha) -> "SynthesisElem":
        """src/lib.rs:404-414"""
        a = f32(alpha)
        om = f32(f32(1.0) - a)
        b = lambda x, y: (x * om + y * a).astype(f32)  # noqa: E731
        return SynthesisElem(f32(f32(self.frequency * om) + f32(other.frequency * a)),
                             b(self.formant_freq, other.formant_freq), b(self.formant_bw, other.formant_bw),
                             b(self.formant_smooth, other.formant_smooth), b(self.formant_breath, other.formant_breath),
                             b(self.formant_turb, other.formant_turb), b(self.formant_amp, other.formant_amp))

    def resample(self, old_sample_rate, new_sample_rate) -> "SynthesisElem":
        """src/lib.rs:418-440"""
        scale = f32(f32(old_sample_rate) / f32(new_sample_rate))
        ff = (self.formant_freq * scale).astype(f32)
        return SynthesisElem(
            f32(min(f32(self.frequency * scale), f32(0.5))),
            np.minimum(ff, f32(0.5)).astype(f32),
            (self.formant_bw * scale).astype(f32),
            (self.formant_smooth * scale).astype(f32),
            self.formant_breath.copy(), self.formant_turb.copy(),
            np.where(ff > f32(0.5), f32(0.0), self.formant_amp).astype(f32))

    def copy_with_frequency(self, frequency) -> "SynthesisElem":
        """src/lib.rs:445-450"""
        return replace(self, frequency=f32(min(f32(frequency), f32(0.5))))

    def copy_silent(self) -> "SynthesisElem":
        """src/lib.rs:454-459"""
        return replace(self, formant_amp=np.zeros(8, f32))

    def to_record(self) -> np.ndarray:
        r = np.zeros((), ELEM_DT)
        for k in ELEM_DT.names:
            r[k] = getattr(self, k)
        return r


@dataclass
class SequenceElem:
    """reference src/lib.rs:814-835"""
    elem: Optional[SynthesisElem]
    length: np.float32
    blend_length: np.float32

    @staticmethod
    def new(elem, length, blend_length) -> "SequenceElem":
        return SequenceElem(elem, f32(length), f32(blend_length))


@dataclass
class Voice:
    """reference src/lib.rs:696-717"""
    sample_rate: np.float32
    phonemes: "object"
    center_frequency: np.float32
    jitter_frequency: np.float32
    jitter_delta_frequency: np.float32
    jitter_delta_formant_frequency: np.float32
    jitter_delta_amplitude: np.float32

    def storage(self) -> np.ndarray:
        """VoiceStorage as grail_elem records indexed by (phoneme id - 3), in make_phonemes! order (src/lib.rs:684-687)"""
        out = np.zeros(2, _ffi.ELEM_DT)
        out[0] = self.phonemes.a.to_record()
        out[1] = self.phonemes.e.to_record()
        return out

    def params(self, jitter_seed: int = 0, synth_seed: int = 0) -> np.ndarray:
        """the grail_voice_params record crossing the C ABI"""
        v = np.zeros((), VOICE_DT)
        for k in ("sample_rate", "jitter_frequency", "jitter_delta_frequency", "jitter_delta_formant_frequency",
                  "jitter_delta_amplitude"):
            v[k] = getattr(self, k)
        v["jitter_seed"] = jitter_seed & 0xFFFFFFFF
        v["synth_seed"] = synth_seed & 0xFFFFFFFF
        return v


def pack_sequence(elems: Iterable[SequenceElem]) -> np.ndarray:
    """SequenceElems -> grail_seq_elem records"""
    lst = list(elems)
    out = np.zeros(len(lst), SEQ_ELEM_DT)
    for i, e in enumerate(lst):
        out[i]["length"] = e.length
        out[i]["blend_length"] = e.blend_length
        if e.elem is not None:
            out[i]["has_elem"] = 1
            out[i]["elem"] = e.elem.to_record()
    return out


# ------------------------------------------------------------------------------------------------
# device context / plans
# ------------------------------------------------------------------------------------------------
class Context:
    """one CUDA device + stream (grail_ctx).  Raises GrailError(ERR_NO_DEVICE) without a GPU."""

    def __init__(self, device: int = 0):
        self._L = _ffi.lib()
        h = C.c_void_p()
        rc = self._L.grail_cuda_create(device, C.byref(h))
        if rc:
            raise GrailError(rc)
        self._h = h
        self.device = device
        self._pinned = []

    def close(self):
        if getattr(self, "_h", None):
            for p in self._pinned:
                self._L.grail_cuda_host_free(self._h, p)
            self._pinned = []
            self._L.grail_cuda_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc:
            raise GrailError(rc, self._L.grail_cuda_last_error(self._h).decode())

    def set_option(self, key: str, value: float):
        self._check(self._L.grail_cuda_set_option(self._h, key.encode(), float(value)))

    def synchronize(self):
        self._check(self._L.grail_cuda_synchronize(self._h))

    @property
    def stream_handle(self) -> int:
        return int(self._L.grail_cuda_stream_handle(self._h) or 0)

    def pinned_empty(self, n: int, dtype=np.float32) -> np.ndarray:
        """a numpy array backed by page-locked host memory (freed with the context's process)"""
        nbytes = int(n) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._check(self._L.grail_cuda_host_alloc(self._h, max(nbytes, 1), C.byref(p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        self._pinned.append(p)      # owned by the context; released in close()
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    def probe_fp32_peak(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._check(self._L.grail_cuda_probe_fp32_peak(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"ffma_flops": a.value, "mufu_ops": b.value, "sm_mhz_effective": c.value}

    def copy_segments(self, d_dst: int, d_src: int, dst_off: np.ndarray, src_off: np.ndarray, lengths: np.ndarray,
                      elem_bytes: int = 4):
        """device segment copy (grail_cuda_copy_segments): the reorder step of a multi-GPU output gather"""
        d, s_, l = (np.ascontiguousarray(x, np.uint64) for x in (dst_off, src_off, lengths))
        self._check(self._L.grail_cuda_copy_segments(self._h, C.c_void_p(d_dst), C.c_void_p(d_src), ptr(d), ptr(s_), ptr(l),
                                                     len(l), int(elem_bytes)))

    # -- one-shot (host buffers in, host buffer out): the drop-in for draining the iterator chain
    def synthesize_batch(self, elems: np.ndarray, utt_offsets: np.ndarray, voices: np.ndarray,
                         out: Optional[np.ndarray] = None, out_offsets: Optional[np.ndarray] = None, fmt: int = _ffi.F32):
        """drain a batch of chains into `out` (f32, or i16 as the reference's WAV writer converts: fmt=_ffi.I16)"""
        e = np.ascontiguousarray(elems, SEQ_ELEM_DT)
        offs = np.ascontiguousarray(utt_offsets, np.uint32)
        v = np.ascontiguousarray(voices, VOICE_DT).reshape(-1)
        n = len(offs) - 1
        if len(v) != n:
            raise ValueError("one grail_voice_params per utterance")
        if out_offsets is None:
            out_offsets = np.concatenate([[0], np.cumsum(count_samples(e, offs, v))]).astype(np.uint64)
        oo = np.ascontiguousarray(out_offsets, np.uint64)
        dt = np.float32 if fmt == _ffi.F32 else np.int16
        if out is None:
            out = np.empty(int(oo[-1]), dt)
        if out.dtype != dt:
            raise ValueError("output buffer dtype does not match the sample format")
        fn = self._L.grail_cuda_synthesize_batch if fmt == _ffi.F32 else self._L.grail_cuda_synthesize_batch_i16
        self._check(fn(self._h, ptr(e), ptr(offs), ptr(v), n, ptr(out), ptr(oo), 0))
        return out, oo

    def plan(self, elems: np.ndarray, utt_offsets: np.ndarray, voices: np.ndarray) -> "Plan":
        return Plan(self, elems, utt_offsets, voices)

    def plan_phonemes(self, phonemes: np.ndarray, utt_offsets: np.ndarray, storages: np.ndarray, voices: np.ndarray,
                      center_frequency: Optional[np.ndarray] = None, utt_storage: Optional[np.ndarray] = None) -> "Plan":
        """phoneme-level input, Selector (and the Intonator stub) on the device.  `phonemes` is either
        PHONEME_ELEM_DT records (PhonemeElem, src/lib.rs:961-973) or uint8 phoneme ids with one `center_frequency`
        per utterance; `storages` is [n_storages, n_sounds] grail_elem records (VoiceStorage, :651-659)."""
        return Plan(self, None, utt_offsets, voices, phonemes=phonemes, storages=storages,
                    center_frequency=center_frequency, utt_storage=utt_storage)

    def stream(self, voice_params: np.ndarray) -> "Stream":
        return Stream(self, voice_params)


class Plan:
    """a batch resident in HBM (grail_plan): upload once, launch many times"""

    def __init__(self, ctx: Context, elems, utt_offsets, voices, phonemes=None, storages=None, center_frequency=None,
                 utt_storage=None):
        self.ctx = ctx
        self._L = ctx._L
        offs = np.ascontiguousarray(utt_offsets, np.uint32)
        v = np.ascontiguousarray(voices, VOICE_DT).reshape(-1)
        self.n_utts = len(offs) - 1
        h = C.c_void_p()
        if phonemes is None:
            e = np.ascontiguousarray(elems, SEQ_ELEM_DT)
            ctx._check(self._L.grail_cuda_plan_create(ctx._h, ptr(e), ptr(offs), ptr(v), self.n_utts, C.byref(h)))
        else:
            st = np.ascontiguousarray(storages, _ffi.ELEM_DT)
            st = st.reshape(1, -1) if st.ndim == 1 else st
            us = None if utt_storage is None else np.ascontiguousarray(utt_storage, np.uint32)
            ph = np.asarray(phonemes)
            if ph.dtype == _ffi.PHONEME_ELEM_DT:
                ph = np.ascontiguousarray(ph)
                ctx._check(self._L.grail_cuda_plan_create_phoneme_elems(ctx._h, ptr(ph), ptr(offs), ptr(st), st.shape[1],
                                                                        st.shape[0], ptr(us), ptr(v), self.n_utts, C.byref(h)))
            else:
                ids = np.ascontiguousarray(ph, np.uint8)
                cf = np.ascontiguousarray(center_frequency, np.float32).reshape(-1)
                if len(cf) != self.n_utts:
                    raise ValueError("one center_frequency per utterance")
                ctx._check(self._L.grail_cuda_plan_create_phonemes(ctx._h, ptr(ids), ptr(offs), ptr(cf), ptr(st), st.shape[1],
                                                                   st.shape[0], ptr(us), ptr(v), self.n_utts, C.byref(h)))
        self._h = h
        self.total_samples = int(self._L.grail_cuda_plan_total_samples(h))
        oo = np.zeros(self.n_utts + 1, np.uint64)
        ctx._check(self._L.grail_cuda_plan_out_offsets(h, ptr(oo)))
        self.out_offsets = oo

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self._L.grail_cuda_plan_destroy(self._h)
        self._h = None

    __del__ = close

    def device_output(self, fmt: int = _ffi.F32) -> int:
        p = C.c_void_p()
        self.ctx._check(self._L.grail_cuda_plan_device_output(self._h, fmt, C.byref(p)))
        return int(p.value)

    def launch(self, d_out: Optional[int] = None, fmt: int = _ffi.F32):
        """enqueue the whole path on the ctx stream (asynchronous)"""
        if d_out is None:
            d_out = self.device_output(fmt)
        self.ctx._check(self._L.grail_cuda_plan_launch(self._h, C.c_void_p(d_out), fmt))

    def launch_interleaved(self, d_out: int, channels: int, fmt: int = _ffi.F32):
        """every sample written `channels` times, interleaved (the examples' channel duplication) into a caller buffer"""
        self.ctx._check(self._L.grail_cuda_plan_launch_interleaved(self._h, C.c_void_p(d_out), fmt, channels))

    def join(self):
        """make the ctx's main stream wait (on the device) for this plan's in-flight launches"""
        self.ctx._check(self._L.grail_cuda_plan_join(self._h))

    def read_output(self, out: Optional[np.ndarray] = None, fmt: int = _ffi.F32) -> np.ndarray:
        if out is None:
            out = np.empty(self.total_samples, np.float32 if fmt == _ffi.F32 else np.int16)
        self.ctx._check(self._L.grail_cuda_plan_read_output(self._h, fmt, ptr(out)))
        return out

    def timings(self) -> dict:
        t = Timings()
        self.ctx._check(self._L.grail_cuda_plan_timings(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timings._fields_}

    def phase_scan_stats(self) -> dict:
        """exact parallel phase scan of the last launch: scans, converged, max refinement rounds, refused"""
        st = np.zeros(4, np.uint32)
        self.ctx._check(self._L.grail_cuda_plan_phase_scan_stats(self._h, ptr(st)))
        return {"scans": int(st[0]), "converged": int(st[1]), "max_rounds": int(st[2]), "refused": int(st[3])}

    def phase_stats(self) -> dict:
        """chunk-parallel exact carrier phase of the last launch (grail_cuda_plan_phase_stats)"""
        st = np.zeros(8, np.uint32)
        self.ctx._check(self._L.grail_cuda_plan_phase_stats(self._h, ptr(st)))
        return {"chunks": int(st[0]), "walks": int(st[1]), "unproven_utterances": int(st[2]), "repair_rounds": int(st[3]),
                "failed_boundaries": int(st[4]), "phase_chunk": int(st[5]), "warmup_from_zero": int(st[6])}

    def read_intermediates(self):
        """bit-exact taps: (F_t, carrier phase before each sample, polyBLEP saw), packed like the output"""
        n = self.total_samples
        f, p, s = (np.zeros(n, np.float32) for _ in range(3))
        self.ctx._check(self._L.grail_cuda_plan_read_intermediates(self._h, ptr(f), ptr(p), ptr(s)))
        return f, p, s


class Stream:
    """one unbounded utterance with carried iterator state (grail_stream): the drop-in for pulling the reference's
    iterator chain from an audio callback (examples/interactive.rs:31-69)"""

    def __init__(self, ctx: Context, voice_params: np.ndarray):
        self.ctx = ctx
        self._L = ctx._L
        v = np.ascontiguousarray(voice_params, VOICE_DT).reshape(-1)[:1].copy()
        h = C.c_void_p()
        ctx._check(self._L.grail_cuda_stream_new(ctx._h, ptr(v), C.byref(h)))
        self._h = h

    def push(self, elems: np.ndarray):
        e = np.ascontiguousarray(elems, SEQ_ELEM_DT).reshape(-1)
        self.ctx._check(self._L.grail_cuda_stream_push(self._h, ptr(e), len(e)))

    def finish(self):
        self.ctx._check(self._L.grail_cuda_stream_finish(self._h))

    def pull(self, max_samples: int) -> np.ndarray:
        """up to max_samples more samples; fewer (possibly none) when the upstream has run dry"""
        out = np.empty(int(max_samples), np.float32)
        n = C.c_uint64(0)
        self.ctx._check(self._L.grail_cuda_stream_pull(self._h, ptr(out), int(max_samples), C.byref(n)))
        return out[: n.value]

    def close(self):
        if getattr(self, "_h", None):
            self._L.grail_cuda_stream_free(self._h)
            self._h = None


def pull_streams(streams, max_samples) -> list:
    """the next windows of several streams of one Context as one launch per kernel (grail_cuda_streams_pull); returns
    one array per stream"""
    if not streams:
        return []
    n = len(streams)
    ms = np.ascontiguousarray(np.broadcast_to(np.asarray(max_samples, np.uint64), (n,)))
    outs = [np.empty(int(m), np.float32) for m in ms]
    hs = (C.c_void_p * n)(*[s._h for s in streams])
    ps = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    wr = np.zeros(n, np.uint64)
    ctx = streams[0].ctx
    ctx._check(ctx._L.grail_cuda_streams_pull(hs, n, ps, ptr(ms), ptr(wr)))
    return [o[: int(w)] for o, w in zip(outs, wr)]

    __del__ = close


def save_wav(path: str, data: np.ndarray, sample_rate: int, channels: int = 1) -> None:
    """16-bit PCM RIFF writer of examples/cli.rs:28-67; `data` is f32 (converted like the reference,
    `(x * i16::MAX as f32) as i16`) or already int16 (GRAIL_I16 device output)"""
    import struct
    pcm = data if data.dtype == np.int16 else np.clip(np.trunc(np.nan_to_num(data.astype(np.float32)) * np.float32(32767.0)), -32768, 32767).astype(np.int16)
    raw = pcm.astype("<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " +
                struct.pack("<IHHIIHH", 16, 1, channels, sample_rate, sample_rate * 2 * channels, 2 * channels, 16) +
                b"data" + struct.pack("<I", len(raw)) + raw)


def count_samples(elems: np.ndarray, utt_offsets: np.ndarray, voices: np.ndarray) -> np.ndarray:
    """exact f32-clock sample counts (host only; grail_cuda_count_samples)"""
    L = _ffi.lib()
    e = np.ascontiguousarray(elems, SEQ_ELEM_DT)
    offs = np.ascontiguousarray(utt_offsets, np.uint32)
    v = np.ascontiguousarray(voices, VOICE_DT).reshape(-1)
    counts = np.zeros(len(offs) - 1, np.uint64)
    rc = L.grail_cuda_count_samples(ptr(e), ptr(offs), ptr(v), len(offs) - 1, ptr(counts))
    if rc:
        raise GrailError(rc)
    return counts


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ------------------------------------------------------------------------------------------------
# the iterator verbs
# ------------------------------------------------------------------------------------------------
class Sequencer:
    """`.sequence(voice)` (src/lib.rs:941-949): a lazy descriptor; nothing runs until `.synthesize()`"""

    def __init__(self, upstream: Iterable[SequenceElem], voice: Voice):
        self.upstream = upstream
        self.voice = voice

    def jitter(self, seed: int, voice: Voice) -> "Jitter":
        return Jitter(self, seed, voice)

    def synthesize(self, ctx: Optional[Context] = None) -> "Synthesize":
        # no Jitter stage == a Jitter whose three deltas are zero (x + n*0 and a * (1 - d*0) are identities)
        quiet = replace(self.voice, jitter_delta_frequency=f32(0), jitter_delta_formant_frequency=f32(0),
                        jitter_delta_amplitude=f32(0))
        return Synthesize(self, 0, quiet, ctx)


class Jitter:
    """`.jitter(seed, voice)` (src/lib.rs:786-797): lazy descriptor"""

    def __init__(self, sequencer: Sequencer, seed: int, voice: Voice):
        if not isinstance(sequencer, Sequencer):
            raise TypeError("the GPU path takes `.sequence(voice).jitter(seed, voice).synthesize()`; an arbitrary "
                            "Iterator<Item = SynthesisElem> cannot be lowered and there is no CPU path")
        self.sequencer = sequencer
        self.seed = seed
        self.voice = voice

    def synthesize(self, ctx: Optional[Context] = None) -> "Synthesize":
        return Synthesize(self.sequencer, self.seed, self.voice, ctx)


class Synthesize:
    """`.synthesize()` (src/lib.rs:587-596): Iterator<Item = f32>.  The first `next()` drains the upstream
    SequenceElems, synthesizes the whole utterance on the device and then yields from the buffer."""

    def __init__(self, sequencer: Sequencer, seed: int, jitter_voice: Voice, ctx: Optional[Context]):
        self.sequencer = sequencer
        self.seed = seed
        self.jitter_voice = jitter_voice
        self.ctx = ctx
        self._buf: Optional[np.ndarray] = None
        self._pos = 0

    def _run(self):
        seq = pack_sequence(self.sequencer.upstream)
        v = self.jitter_voice.params(self.seed, 0)
        v["sample_rate"] = self.sequencer.voice.sample_rate  # the Sequencer's voice sets delta_time (:944)
        ctx = self.ctx or default_context()
        out, _ = ctx.synthesize_batch(seq, np.array([0, len(seq)], np.uint32), v.reshape(1))
        self._buf = out

    def collect(self) -> np.ndarray:
        """the `Vec::extend(iterator)` of examples/cli.rs:175-184"""
        if self._buf is None:
            self._run()
        out = self._buf[self._pos:]
        self._pos = len(self._buf)
        return out

    def __iter__(self) -> Iterator[np.float32]:
        return self

    def __next__(self) -> np.float32:
        if self._buf is None:
            self._run()
        if self._pos >= len(self._buf):
            raise StopIteration
        x = self._buf[self._pos]
        self._pos += 1
        return x


def sequence(elems: Iterable[SequenceElem], voice: Voice) -> Sequencer:
    """IntoSequencer::sequence (src/lib.rs:936-953)"""
    return Sequencer(elems, voice)
