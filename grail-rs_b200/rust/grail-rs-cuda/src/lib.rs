//! Same verbs as the reference (`IntoSequencer::sequence` src/lib.rs:936-953, `IntoJitter::jitter` :781-801,
//! `IntoSynthesize::synthesize` :582-600), but `Sequencer` and `Jitter` are lazy descriptors and `Synthesize`
//! drains the upstream `SequenceElem`s, runs the per-sample work on the GPU through `grail-cuda-sys`, and then
//! yields `f32`s from the returned buffer.  Use it by swapping the three trait imports:
//!
//! ```ignore
//! use grail_rs::{IntoIntonator, IntoSelector, IntoTranscriber};
//! use grail_rs_cuda::{IntoSequencer, IntoJitter, IntoSynthesize};   // instead of grail_rs::{...}
//! ```
#![forbid(unsafe_code)]
use grail_cuda_sys as sys;
use grail_rs::{SequenceElem, SynthesisElem, Voice};

pub struct Sequencer<T: Iterator<Item = SequenceElem>> { iter: T, voice: Voice }
pub struct Jitter<T: Iterator<Item = SequenceElem>> { seq: Sequencer<T>, seed: u32, voice: Voice }
pub struct Synthesize { buf: Vec<f32>, pos: usize }

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

pub trait IntoSynthesize { fn synthesize(self) -> Synthesize; }
impl<T: Iterator<Item = SequenceElem>> IntoSynthesize for Jitter<T> {
    fn synthesize(self) -> Synthesize {
        let elems: Vec<sys::grail_seq_elem> = self.seq.iter.map(|s| sys::grail_seq_elem {
            has_elem: s.elem.is_some() as u32,
            elem: s.elem.as_ref().map(pack_elem).unwrap_or(unsafe_free_zeroed_elem()),
            length: s.length, blend_length: s.blend_length,
        }).collect();
        let v = sys::grail_voice_params {
            sample_rate: self.seq.voice.sample_rate,              // the Sequencer's voice sets delta_time (:944)
            jitter_frequency: self.voice.jitter_frequency,
            jitter_delta_frequency: self.voice.jitter_delta_frequency,
            jitter_delta_formant_frequency: self.voice.jitter_delta_formant_frequency,
            jitter_delta_amplitude: self.voice.jitter_delta_amplitude,
            jitter_seed: self.seed, synth_seed: 0,                // Synthesize noise seed is 0 (:594)
        };
        let offs = [0u32, elems.len() as u32];
        let n = sys::Ctx::count_samples(&elems, &offs, &[v]).expect("grail_cuda_count_samples")[0];
        let mut buf = vec![0f32; n as usize];
        sys::Ctx::new(0).expect("no CUDA device: this crate has no CPU path")
            .synthesize_batch(&elems, &offs, &[v], &mut buf, &[0, n]).expect("grail_cuda_synthesize_batch");
        Synthesize { buf, pos: 0 }
    }
}
fn unsafe_free_zeroed_elem() -> sys::grail_elem {
    sys::grail_elem { frequency: 0.0, formant_freq: [0.0; 8], formant_bw: [0.0; 8], formant_smooth: [0.0; 8],
                      formant_breath: [0.0; 8], formant_turb: [0.0; 8], formant_amp: [0.0; 8] }
}

impl Iterator for Synthesize {
    type Item = f32;
    fn next(&mut self) -> Option<f32> {
        let x = self.buf.get(self.pos).copied();
        self.pos += 1;
        x
    }
}
