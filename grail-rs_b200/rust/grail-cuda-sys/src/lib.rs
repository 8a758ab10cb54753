//! Raw bindings to `include/grail_cuda.h`.  All `unsafe` lives in this crate so that the facade crate
//! (`grail-rs-cuda`) can keep the reference's `#![forbid(unsafe_code)]` (reference src/lib.rs:2).
#![allow(non_camel_case_types, non_upper_case_globals, non_snake_case)]
include!(concat!(env!("OUT_DIR"), "/bindings.rs"));

/// Safe-ish owner of a `grail_ctx`.
pub struct Ctx(*mut grail_ctx);

impl Ctx {
    pub fn new(device: i32) -> Result<Self, i32> {
        let mut p = std::ptr::null_mut();
        match unsafe { grail_cuda_create(device, &mut p) } {
            0 => Ok(Ctx(p)),
            e => Err(e),
        }
    }

    /// exact per-utterance sample counts (host only)
    pub fn count_samples(elems: &[grail_seq_elem], utt_offsets: &[u32], voices: &[grail_voice_params]) -> Result<Vec<u64>, i32> {
        assert_eq!(utt_offsets.len(), voices.len() + 1, "utt_offsets has one entry more than there are utterances");
        assert!(utt_offsets.windows(2).all(|w| w[0] <= w[1]) && *utt_offsets.last().unwrap() as usize <= elems.len());
        let n = voices.len() as u32;
        let mut counts = vec![0u64; voices.len()];
        match unsafe { grail_cuda_count_samples(elems.as_ptr(), utt_offsets.as_ptr(), voices.as_ptr(), n, counts.as_mut_ptr()) } {
            0 => Ok(counts),
            e => Err(e),
        }
    }

    /// drains the whole batch into `out` (packed, utterance `u` at `out_offsets[u]`)
    pub fn synthesize_batch(&mut self, elems: &[grail_seq_elem], utt_offsets: &[u32], voices: &[grail_voice_params],
                            out: &mut [f32], out_offsets: &[u64]) -> Result<(), i32> {
        check_batch_slices(elems.len(), utt_offsets, voices.len(), out.len(), out_offsets);
        let rc = unsafe {
            grail_cuda_synthesize_batch(self.0, elems.as_ptr(), utt_offsets.as_ptr(), voices.as_ptr(), voices.len() as u32,
                                        out.as_mut_ptr(), out_offsets.as_ptr(), 0)
        };
        if rc == 0 { Ok(()) } else { Err(rc) }
    }
}

impl Ctx {
    /// `synthesize_batch` with the samples converted on the device as the reference's WAV writer does
    /// (`(x * i16::MAX as f32) as i16`, examples/cli.rs:49-51): half the device-to-host bytes.
    pub fn synthesize_batch_i16(&mut self, elems: &[grail_seq_elem], utt_offsets: &[u32], voices: &[grail_voice_params],
                                out: &mut [i16], out_offsets: &[u64]) -> Result<(), i32> {
        check_batch_slices(elems.len(), utt_offsets, voices.len(), out.len(), out_offsets);
        let rc = unsafe {
            grail_cuda_synthesize_batch_i16(self.0, elems.as_ptr(), utt_offsets.as_ptr(), voices.as_ptr(), voices.len() as u32,
                                            out.as_mut_ptr(), out_offsets.as_ptr(), 0)
        };
        if rc == 0 { Ok(()) } else { Err(rc) }
    }
}

/// the slice invariants the C side relies on (safe Rust must not be able to make C read or write out of bounds)
fn check_batch_slices(n_elems: usize, utt_offsets: &[u32], n_utts: usize, out_len: usize, out_offsets: &[u64]) {
    assert_eq!(utt_offsets.len(), n_utts + 1);
    assert_eq!(out_offsets.len(), n_utts + 1);
    assert!(utt_offsets.windows(2).all(|w| w[0] <= w[1]) && *utt_offsets.last().unwrap() as usize <= n_elems);
    assert!(out_offsets.windows(2).all(|w| w[0] <= w[1]) && *out_offsets.last().unwrap() as usize <= out_len);
}

/// One unbounded utterance with carried iterator state (`grail_stream`): what the facade's lazy `Synthesize` pulls from.
pub struct Stream(*mut grail_stream);
// the handle is used from one thread at a time and may move to the audio thread (examples/interactive.rs moves the iterator)
unsafe impl Send for Stream {}

impl Stream {
    pub fn new(ctx: &Ctx, voice: &grail_voice_params) -> Result<Self, i32> {
        let mut p = std::ptr::null_mut();
        match unsafe { grail_cuda_stream_new(ctx.0, voice, &mut p) } { 0 => Ok(Stream(p)), e => Err(e) }
    }
    pub fn push(&mut self, elems: &[grail_seq_elem]) -> Result<(), i32> {
        match unsafe { grail_cuda_stream_push(self.0, elems.as_ptr(), elems.len() as u32) } { 0 => Ok(()), e => Err(e) }
    }
    pub fn finish(&mut self) -> Result<(), i32> {
        match unsafe { grail_cuda_stream_finish(self.0) } { 0 => Ok(()), e => Err(e) }
    }
    /// up to `out.len()` more samples; returns how many were written (0: the upstream has run dry)
    pub fn pull(&mut self, out: &mut [f32]) -> Result<usize, i32> {
        let mut n = 0u64;
        match unsafe { grail_cuda_stream_pull(self.0, out.as_mut_ptr(), out.len() as u64, &mut n) } { 0 => Ok(n as usize), e => Err(e) }
    }
}
impl Drop for Stream {
    fn drop(&mut self) { unsafe { grail_cuda_stream_free(self.0) } }
}
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}   // the C side serialises nothing: the facade keeps its shared Ctx behind a Mutex

/// The reference's `Transcriber` (src/lib.rs:1098-1207) over a batch of texts on all host cores (host only).
/// `rules` must be sorted by string; returns (phoneme ids, utterance offsets) in the layout `Plan::from_phonemes` takes.
pub fn transcribe_batch(texts: &[&str], rules: &[(&std::ffi::CStr, &[u8])], case_sensitive: bool, leading_silence: bool)
                        -> Result<(Vec<u8>, Vec<u32>), i32> {
    let ptrs: Vec<*const std::os::raw::c_char> = texts.iter().map(|t| t.as_ptr() as *const _).collect();
    let lens: Vec<usize> = texts.iter().map(|t| t.len()).collect();
    let crules: Vec<grail_transcription_rule> = rules.iter()
        .map(|(s, p)| grail_transcription_rule { string: s.as_ptr(), phonemes: p.as_ptr(), n_phonemes: p.len() as u32 })
        .collect();
    let mut offs = vec![0u32; texts.len() + 1];
    let count = |ids: *mut u8, cap: u64, offs: &mut Vec<u32>| unsafe {
        grail_cuda_transcribe_batch(ptrs.as_ptr(), lens.as_ptr(), texts.len() as u32, crules.as_ptr(), crules.len() as u32,
                                    case_sensitive as i32, leading_silence as i32, ids, cap, offs.as_mut_ptr(), 0)
    };
    match count(std::ptr::null_mut(), 0, &mut offs) { 0 => {}, e => return Err(e) }
    let mut ids = vec![0u8; *offs.last().unwrap() as usize];
    match count(ids.as_mut_ptr(), ids.len() as u64, &mut offs) { 0 => Ok((ids, offs)), e => Err(e) }
}

/// A batch resident in HBM (`grail_plan`): built once, launched many times.
pub struct Plan(*mut grail_plan);

impl Plan {
    /// Sequencer records in (the path's own boundary).
    pub fn new(ctx: &mut Ctx, elems: &[grail_seq_elem], utt_offsets: &[u32], voices: &[grail_voice_params]) -> Result<Self, i32> {
        assert_eq!(utt_offsets.len(), voices.len() + 1);
        assert!(utt_offsets.windows(2).all(|w| w[0] <= w[1]) && *utt_offsets.last().unwrap() as usize <= elems.len());
        let mut p = std::ptr::null_mut();
        match unsafe { grail_cuda_plan_create(ctx.0, elems.as_ptr(), utt_offsets.as_ptr(), voices.as_ptr(), voices.len() as u32, &mut p) } {
            0 => Ok(Plan(p)),
            e => Err(e),
        }
    }

    /// Phoneme ids in: `Intonator` (reference src/lib.rs:1057-1075) and `Selector` (:979-1005) run on the device.
    /// `storages` holds `n_sounds` records per voice; `utt_storage[u]` picks the voice of utterance `u`.
    pub fn from_phonemes(ctx: &mut Ctx, phoneme_ids: &[u8], utt_offsets: &[u32], center_frequency: &[f32],
                         storages: &[grail_elem], n_sounds: u32, utt_storage: Option<&[u32]>,
                         voices: &[grail_voice_params]) -> Result<Self, i32> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe {
            grail_cuda_plan_create_phonemes(ctx.0, phoneme_ids.as_ptr(), utt_offsets.as_ptr(), center_frequency.as_ptr(),
                                            storages.as_ptr(), n_sounds, storages.len() as u32 / n_sounds.max(1),
                                            utt_storage.map_or(std::ptr::null(), |s| s.as_ptr()), voices.as_ptr(),
                                            voices.len() as u32, &mut p)
        };
        if rc == 0 { Ok(Plan(p)) } else { Err(rc) }
    }

    /// `PhonemeElem` records in (own lengths and pitches): `Selector` runs on the device.
    pub fn from_phoneme_elems(ctx: &mut Ctx, phonemes: &[grail_phoneme_elem], utt_offsets: &[u32], storages: &[grail_elem],
                              n_sounds: u32, utt_storage: Option<&[u32]>, voices: &[grail_voice_params]) -> Result<Self, i32> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe {
            grail_cuda_plan_create_phoneme_elems(ctx.0, phonemes.as_ptr(), utt_offsets.as_ptr(), storages.as_ptr(), n_sounds,
                                                 storages.len() as u32 / n_sounds.max(1),
                                                 utt_storage.map_or(std::ptr::null(), |s| s.as_ptr()), voices.as_ptr(),
                                                 voices.len() as u32, &mut p)
        };
        if rc == 0 { Ok(Plan(p)) } else { Err(rc) }
    }

    pub fn total_samples(&self) -> u64 {
        unsafe { grail_cuda_plan_total_samples(self.0) }
    }

    /// run the path and copy the packed mono f32 samples to `out` (`total_samples()` entries)
    pub fn synthesize(&mut self, out: &mut [f32]) -> Result<(), i32> {
        assert!(out.len() as u64 >= self.total_samples());
        let mut d = std::ptr::null_mut();
        let mut rc = unsafe { grail_cuda_plan_device_output(self.0, GRAIL_F32 as i32, &mut d) };
        if rc == 0 { rc = unsafe { grail_cuda_plan_launch(self.0, d, GRAIL_F32 as i32) }; }
        if rc == 0 { rc = unsafe { grail_cuda_plan_read_output(self.0, GRAIL_F32 as i32, out.as_mut_ptr() as *mut _) }; }
        if rc == 0 { Ok(()) } else { Err(rc) }
    }
}

impl Drop for Plan {
    fn drop(&mut self) {
        unsafe { grail_cuda_plan_destroy(self.0) }
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { grail_cuda_destroy(self.0) }
    }
}
