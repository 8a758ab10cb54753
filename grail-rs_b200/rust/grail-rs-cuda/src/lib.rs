//! Same verbs as the reference (`IntoSequencer::sequence` src/lib.rs:936-953, `IntoJitter::jitter` :781-801,
//! `IntoSynthesize::synthesize` :582-600), but `Sequencer` and `Jitter` are lazy descriptors and `Synthesize`
//! pulls the upstream `SequenceElem`s lazily, a window of samples at a time, through the C ABI's streaming entry
//! points (`grail_cuda_stream_*`): finite and infinite upstreams alike (examples/cli.rs and examples/interactive.rs).  Use it by swapping the three trait imports:
//!
//! ```ignore
//! use grail_rs::{IntoIntonator, IntoSelector, IntoTranscriber};
//! use grail_rs_cuda::{IntoSequencer, IntoJitter, IntoSynthesize};   // instead of grail_rs::{...}
//! ```
#![forbid(unsafe_code)]
use grail_cuda_sys as sys;
use grail_rs::{SequenceElem, SynthesisElem, Voice};

use std::sync::{Mutex, OnceLock};

pub struct Sequencer<T: Iterator<Item = SequenceElem>> { iter: T, voice: Voice }
pub struct Jitter<T: Iterator<Item = SequenceElem>> { seq: Sequencer<T>, seed: u32, voice: Voice }

/// `Iterator<Item = f32>` that pulls its upstream LAZILY, a window at a time, through `grail_cuda_stream_{push,pull}`:
/// the upstream may be infinite (`examples/interactive.rs:31-38` feeds it from `repeat_with`), and only as many
/// `SequenceElem`s are drawn as the stream's one-element look-ahead needs.  The same logic, line for line, is the C++
/// facade's `StreamSynthesize` (grail-rs_b200/cpp/grail.hpp), which is what the repository's tests run -- this crate
/// cannot be compiled in the build image (no rustc).  Errors never panic and never unwind: the reference's hot path
/// has no error channel, so an error ends the iterator (`None`) and stays readable through `status()`.
pub struct Synthesize<T: Iterator<Item = SequenceElem>> {
    up: T,
    voice: sys::grail_voice_params,
    stream: Option<sys::Stream>,
    buf: Vec<f32>,
    pos: usize,
    len: usize,
    up_done: bool,
    ended: bool,
    status: i32,
}

pub trait IntoSequencer: IntoIterator<Item = SequenceElem> + Sized {
    fn sequence(self, voice: Voice) -> Sequencer<Self::IntoIter> { Sequencer { iter: self.into_iter(), voice } }
}
impl<T: IntoIterator<Item = SequenceElem> + Sized> IntoSequencer for T {}

pub trait IntoJitter<T: Iterator<Item = SequenceElem>> { fn jitter(self, seed: u32, voice: Voice) -> Jitter<T>; }
impl<T: Iterator<Item = SequenceElem>> IntoJitter<T> for Sequencer<T> {
    fn jitter(self, seed: u32, voice: Voice) -> Jitter<T> { Jitter { seq: self, seed, voice } }
}

fn pack_elem(e: &SynthesisElem) -> sys::grail_elem {
    // grail-rs keeps `Array`'s field private; a one-line `pub fn to_array(self) -> [f32; 8]` upstream (or the
    // `Debug` round trip used by the tests) is the only change the reference needs for this adapter.
    sys::grail_elem {
        frequency: e.frequency,
        formant_freq: e.formant_freq.to_array(), formant_bw: e.formant_bw.to_array(),
        formant_smooth: e.formant_smooth.to_array(), formant_breath: e.formant_breath.to_array(),
        formant_turb: e.formant_turb.to_array(), formant_amp: e.formant_amp.to_array(),
    }
}
fn pack_seq(s: &SequenceElem) -> sys::grail_seq_elem {
    const ZERO: sys::grail_elem = sys::grail_elem { frequency: 0.0, formant_freq: [0.0; 8], formant_bw: [0.0; 8], formant_smooth: [0.0; 8],
                                                    formant_breath: [0.0; 8], formant_turb: [0.0; 8], formant_amp: [0.0; 8] };
    sys::grail_seq_elem { has_elem: s.elem.is_some() as u32, elem: s.elem.as_ref().map(pack_elem).unwrap_or(ZERO),
                          length: s.length, blend_length: s.blend_length }
}

/// ONE context per process, created on first use (a `Ctx` owns a CUDA stream and a device-memory pool: creating one
/// per `.synthesize()` call, as the first version of this crate did, costs milliseconds each time).
fn shared_ctx() -> Option<&'static Mutex<sys::Ctx>> {
    static CTX: OnceLock<Option<Mutex<sys::Ctx>>> = OnceLock::new();
    CTX.get_or_init(|| sys::Ctx::new(0).ok().map(Mutex::new)).as_ref()      // no device: no CPU path, the iterator ends at once
}

/// samples per device round trip; an audio callback's buffer size is a good value
pub const WINDOW: usize = 2048;

/// The GPU path engages for the chain the reference's examples build, `.sequence(v).jitter(seed, v).synthesize()`.
/// An arbitrary `Iterator<Item = SynthesisElem>` (the reference's blanket impl, src/lib.rs:582-600) is REFUSED at
/// compile time: this trait is implemented for `Jitter<T>` only, because a per-sample stream of 196-byte parameter
/// frames has no phoneme structure left to plan from; such chains keep using `grail_rs::IntoSynthesize` on the CPU.
pub trait IntoSynthesize<T: Iterator<Item = SequenceElem>> { fn synthesize(self) -> Synthesize<T>; }
impl<T: Iterator<Item = SequenceElem>> IntoSynthesize<T> for Jitter<T> {
    fn synthesize(self) -> Synthesize<T> {
        let voice = sys::grail_voice_params {
            sample_rate: self.seq.voice.sample_rate,              // the Sequencer's voice sets delta_time (:944)
            jitter_frequency: self.voice.jitter_frequency,
            jitter_delta_frequency: self.voice.jitter_delta_frequency,
            jitter_delta_formant_frequency: self.voice.jitter_delta_formant_frequency,
            jitter_delta_amplitude: self.voice.jitter_delta_amplitude,
            jitter_seed: self.seed, synth_seed: 0,                // Synthesize noise seed is 0 (:594)
        };
        // lazy like the reference: nothing is pulled, and no device is touched, before the first next()
        Synthesize { up: self.seq.iter, voice, stream: None, buf: Vec::new(), pos: 0, len: 0, up_done: false, ended: false, status: 0 }
    }
}

impl<T: Iterator<Item = SequenceElem>> Synthesize<T> {
    /// 0 unless an error ended the iterator (a `grail_status`)
    pub fn status(&self) -> i32 { self.status }
    fn fail(&mut self, rc: i32) -> bool { self.status = rc; self.ended = true; false }
    fn refill(&mut self) -> bool {
        if self.ended { return false; }
        if self.stream.is_none() {
            let Some(ctx) = shared_ctx() else { return self.fail(2 /* GRAIL_ERR_NO_DEVICE */) };
            let Ok(guard) = ctx.lock() else { return self.fail(3) };
            match sys::Stream::new(&guard, &self.voice) {
                Ok(s) => { self.stream = Some(s); self.buf.resize(WINDOW, 0.0); }
                Err(rc) => return self.fail(rc),
            }
        }
        loop {
            // the ctx is used from one thread at a time: hold the lock across the device round trip
            let Some(ctx) = shared_ctx() else { return self.fail(2) };
            let Ok(_guard) = ctx.lock() else { return self.fail(3) };
            let stream = self.stream.as_mut().unwrap();
            match stream.pull(&mut self.buf) {
                Err(rc) => return self.fail(rc),
                Ok(n) if n > 0 => { self.pos = 0; self.len = n; return true; }
                Ok(_) => {}
            }
            if self.up_done { self.ended = true; return false; }      // finished and drained: None, like the reference
            // the stream ran dry: one more upstream element (its look-ahead), or the end of the upstream
            let r = match self.up.next() {
                Some(e) => stream.push(&[pack_seq(&e)]),
                None => { self.up_done = true; stream.finish() }
            };
            if let Err(rc) = r { return self.fail(rc); }
        }
    }
}

impl<T: Iterator<Item = SequenceElem>> Iterator for Synthesize<T> {
    type Item = f32;
    fn next(&mut self) -> Option<f32> {
        if self.pos >= self.len && !self.refill() { return None; }
        let x = self.buf[self.pos];
        self.pos += 1;
        Some(x)
    }
}
