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
        let rc = unsafe {
            grail_cuda_synthesize_batch(self.0, elems.as_ptr(), utt_offsets.as_ptr(), voices.as_ptr(), voices.len() as u32,
                                        out.as_mut_ptr(), out_offsets.as_ptr(), 0)
        };
        if rc == 0 { Ok(()) } else { Err(rc) }
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { grail_cuda_destroy(self.0) }
    }
}
