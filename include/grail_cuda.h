/*
 * grail_cuda.h -- C ABI of the B200-native waveform-generation path of grail-rs.
 *
 * The library replaces exactly one piece of the reference: the per-sample iterator chain
 *
 *     <IntoIterator<Item = SequenceElem>>.sequence(voice).jitter(seed, voice).synthesize()
 *
 * (reference src/lib.rs:936-953 IntoSequencer::sequence, :781-801 IntoJitter::jitter,
 *  :582-600 IntoSynthesize::synthesize, drained by examples/cli.rs:175-184).
 * Everything above that cut (Transcriber, Intonator, Selector; per phoneme) stays on the host in
 * the caller's language.  The reference has no FFI of its own; these entry points are what a
 * `grail-cuda-sys` crate would bind (see INTEGRATION.md for the Rust declarations).
 *
 * Conventions: plain pointers and sizes, caller owns every buffer, every function returns a
 * grail_status (0 = ok) and never unwinds.  There is NO CPU fallback: without a CUDA device
 * grail_cuda_create fails with GRAIL_ERR_NO_DEVICE.  A ctx is used from one thread at a time.
 */
#ifndef GRAIL_CUDA_H
#define GRAIL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRAIL_NUM_FORMANTS 8 /* reference NUM_FORMANTS, src/lib.rs:24 */
#define GRAIL_ABI_VERSION 4   /* 2: phoneme-level plans, grail_cuda_transcribe_batch; 3: grail_cuda_plan_phase_stats; 4: grail_cuda_copy_segments, grail_cuda_streams_pull */

typedef enum grail_status {
    GRAIL_OK = 0,
    GRAIL_ERR_INVALID_ARG = 1,   /* null pointer, non-monotone offsets, non-finite scalar ... */
    GRAIL_ERR_NO_DEVICE = 2,     /* no usable CUDA device (there is no CPU path) */
    GRAIL_ERR_CUDA = 3,          /* a CUDA runtime call failed; see grail_cuda_last_error */
    GRAIL_ERR_OOM = 4,           /* device or pinned-host allocation failed */
    GRAIL_ERR_COUNT_MISMATCH = 5,/* out_offsets disagree with the exact f32-clock sample counts */
    GRAIL_ERR_UNSUPPORTED = 6    /* input outside the supported domain (documented per call) */
} grail_status;

/* SynthesisElem, reference src/lib.rs:316-337 (same field order; Rust structs are not repr(C),
 * the -sys crate converts).  All frequencies are normalised to the sample rate. */
typedef struct grail_elem {
    float frequency;
    float formant_freq[GRAIL_NUM_FORMANTS];
    float formant_bw[GRAIL_NUM_FORMANTS];
    float formant_smooth[GRAIL_NUM_FORMANTS];
    float formant_breath[GRAIL_NUM_FORMANTS];
    float formant_turb[GRAIL_NUM_FORMANTS];
    float formant_amp[GRAIL_NUM_FORMANTS];
} grail_elem; /* 196 bytes */

/* SequenceElem, reference src/lib.rs:814-824.  has_elem = 0 encodes elem: None (Silence / Stop /
 * Glide, src/lib.rs:666); `elem` is then ignored. */
typedef struct grail_seq_elem {
    uint32_t   has_elem;
    grail_elem elem;
    float      length;       /* seconds */
    float      blend_length; /* seconds */
} grail_seq_elem; /* 208 bytes */

/* PhonemeElem (src/lib.rs:961-973): Selector's input.  frequency is already divided by the sample rate. */
typedef struct grail_phoneme_elem {
    uint32_t phoneme;       /* GRAIL_PHONEME_* or GRAIL_PHONEME_FIRST_SOUND + index into the voice storage */
    float    length, blend_length, frequency;
} grail_phoneme_elem;       /* 16 bytes */
enum { GRAIL_PHONEME_SILENCE = 0, GRAIL_PHONEME_STOP = 1, GRAIL_PHONEME_GLIDE = 2, GRAIL_PHONEME_FIRST_SOUND = 3 };

/* The Voice scalars the hot path reads (reference src/lib.rs:696-717) plus the two seeds:
 * jitter_seed is the `seed` argument of .jitter(seed, voice) (src/lib.rs:786); synth_seed is the
 * Synthesize noise seed, which the reference hard-codes to 0 (src/lib.rs:594). */
typedef struct grail_voice_params {
    float    sample_rate;
    float    jitter_frequency;
    float    jitter_delta_frequency;
    float    jitter_delta_formant_frequency;
    float    jitter_delta_amplitude;
    uint32_t jitter_seed;
    uint32_t synth_seed;
} grail_voice_params; /* 28 bytes */

typedef struct grail_ctx grail_ctx;       /* one CUDA device + stream + scratch arena */
typedef struct grail_plan grail_plan;     /* a batch resident in HBM, ready to launch */
typedef struct grail_stream grail_stream; /* one unbounded utterance with carried state */

/* Output sample encodings.  F32 is the reference's Iterator<Item = f32>; I16 is the WAV writer's
 * `(x * i16::MAX as f32) as i16` saturating, truncating cast (examples/cli.rs:49-51). */
typedef enum grail_sample_format { GRAIL_F32 = 0, GRAIL_I16 = 1 } grail_sample_format;

/* Kernel-level timing of the most recent launch (CUDA events on the ctx stream), milliseconds. */
typedef struct grail_timings {
    float schedule_ms;  /* exact clock schedule (Sequencer time / jitter phase)        */
    float frequency_ms; /* bit-exact per-sample fundamental F_t                         */
    float phase_ms;     /* bit-exact carrier phase + polyBLEP saw                       */
    float formant_ms;   /* noise, 8x low-pass, 8x SVF band-pass, sum: the dominant kernel */
    float total_ms;     /* first launch to last launch, same stream                     */
    uint32_t n_launches;/* kernels launched by that call                                */
} grail_timings;

/* ---- library ------------------------------------------------------------------------------ */
int grail_cuda_abi_version(void);
/* number of CUDA devices visible to this process (0 if none / no driver) */
int grail_cuda_device_count(void);
const char* grail_cuda_status_string(int status);

/* ---- context ------------------------------------------------------------------------------ */
int  grail_cuda_create(int device, grail_ctx** out_ctx);
void grail_cuda_destroy(grail_ctx* ctx);
/* message of the last failing call on this ctx ("" if none); valid until the next call */
const char* grail_cuda_last_error(const grail_ctx* ctx);
/* the ctx's cudaStream_t as an opaque pointer (for event timing / interop) */
void* grail_cuda_stream_handle(grail_ctx* ctx);
int  grail_cuda_synchronize(grail_ctx* ctx);
/* tuning knobs: "warmup_nepers" (float, filter warm-up depth, default 11.5 ~ 1e-5),
 * "target_lanes" (int, time-chunks the planner aims for), "max_chunk" / "min_chunk" (samples),
 * "formants_per_lane" (1|2), "pipeline" (0|1: overlap consecutive launches of a plan), "e2e_groups", "phase_mode",
 * "phase_chunk", "phase_rounds" (see grail_cuda_plan_phase_stats), "pscan_min_samples",
 * "pscan_cost_model" (0|1), "zero_copy_out" (0|1), "interleave" (0|1: interleave equally long utterances chunk by
 * chunk in the formant kernel's CTAs), "phase_lean" (-1 auto | 0 | 1: which build of the phase kernel) */
int  grail_cuda_set_option(grail_ctx* ctx, const char* key, double value);
/* Environment (diagnostics only): GRAIL_E2E_TRACE=1 prints the timeline of every one-shot batch call on stderr (per
 * utterance group: host time at enqueue, device times of kernels-done and copy-done); GRAIL_PLAN_TRACE=1 prints the host
 * planner's phases per plan. */

/* pinned host memory for full-rate H2D/D2H (optional; pageable buffers also work) */
int  grail_cuda_host_alloc(grail_ctx* ctx, size_t bytes, void** out_ptr);
void grail_cuda_host_free(grail_ctx* ctx, void* ptr);

/* ---- exact sample counts (host only, no device needed) --------------------------------------
 * counts[u] = number of samples Sequencer yields for utterance u, i.e. the number of times the
 * f32 clock `time -= 1/sample_rate` (src/lib.rs:861) stays non-negative, with the carried
 * remainder of src/lib.rs:873,882.  Bit-exact; evaluated in closed form per binade.
 * utt_offsets has n_utts+1 entries indexing `elems`. */
int grail_cuda_count_samples(const grail_seq_elem* elems, const uint32_t* utt_offsets,
                             const grail_voice_params* voices, uint32_t n_utts, uint64_t* counts);

/* ---- host text front-end, batched ------------------------------------------------
 * Transcriber (src/lib.rs:1098-1207): longest-match find-and-replace over a SORTED rule list, one UTF-8 text per
 * utterance, spread over n_threads host threads (0 = all cores).  Host only, no device needed.
 * Writes utt_offsets[0..n_texts] (phoneme index of each utterance, the layout grail_cuda_plan_create_phonemes takes)
 * and, unless ids is NULL (a counting call), the phoneme ids.  leading_silence != 0 starts every text with one
 * Silence as IntoTranscriber::transcribe does (:1201); 0 is the raw struct the reference's tests build (:1213-1358).
 * text_bytes may be NULL (NUL-terminated texts).  Errors: GRAIL_ERR_INVALID_ARG (null pointers, unsorted rules, a
 * rule with an empty string or no phonemes: the reference never terminates on those), GRAIL_ERR_COUNT_MISMATCH
 * (ids_capacity too small; utt_offsets is valid), GRAIL_ERR_UNSUPPORTED (more than 2^32-1 phonemes). */
typedef struct grail_transcription_rule {   /* TranscriptionRule, src/lib.rs:1029-1036 */
    const char*    string;                  /* UTF-8, NUL-terminated */
    const uint8_t* phonemes;                /* GRAIL_PHONEME_* ids */
    uint32_t       n_phonemes;
} grail_transcription_rule;
int grail_cuda_transcribe_batch(const char* const* texts, const size_t* text_bytes, uint32_t n_texts,
                                const grail_transcription_rule* rules, uint32_t n_rules, int case_sensitive,
                                int leading_silence, uint8_t* ids, uint64_t ids_capacity, uint32_t* utt_offsets,
                                int n_threads);

/* ---- one-shot batch synthesis (the drop-in for draining the iterator chain) ------------------
 * elems / utt_offsets / voices are HOST pointers.  out receives utterance u at
 * out[out_offsets[u] .. out_offsets[u+1]) (mono f32); out_offsets[u+1]-out_offsets[u] must equal
 * the exact count, else GRAIL_ERR_COUNT_MISMATCH.  out is a host pointer (pinned or pageable)
 * unless out_is_device != 0.  Blocks until the samples are in `out`.  With a host `out` the batch is processed as a
 * few utterance groups of growing size whose device-to-host copies overlap the next group's kernels (ctx option
 * "e2e_groups": -1 auto, 1 = one plan and one copy); the samples are the same to rounding level either way. */
int grail_cuda_synthesize_batch(grail_ctx* ctx, const grail_seq_elem* elems, const uint32_t* utt_offsets,
                                const grail_voice_params* voices, uint32_t n_utts, float* out,
                                const uint64_t* out_offsets, int out_is_device);
/* the same with the samples converted on the device as the reference's WAV writer does, `(x * i16::MAX as f32) as i16`
 * (examples/cli.rs:49-51: truncating, saturating, NaN -> 0): half the device-to-host bytes */
int grail_cuda_synthesize_batch_i16(grail_ctx* ctx, const grail_seq_elem* elems, const uint32_t* utt_offsets,
                                    const grail_voice_params* voices, uint32_t n_utts, int16_t* out,
                                    const uint64_t* out_offsets, int out_is_device);

/* ---- resident plans (throughput path: inputs stay in HBM between launches) ------------------ */
int  grail_cuda_plan_create(grail_ctx* ctx, const grail_seq_elem* elems, const uint32_t* utt_offsets,
                            const grail_voice_params* voices, uint32_t n_utts, grail_plan** out_plan);
/* ---- phoneme-level input: Selector (src/lib.rs:979-1005) and the Intonator stub (:1057-1075) on the device ----
 * The step immediately before the path: instead of a 208-byte Sequencer record per phoneme, the caller
 * sends PhonemeElem records (16 B) or bare phoneme ids (1 B) plus the voice storages they index; the records the
 * Sequencer reads are written in HBM by a kernel.  Output is bit-identical to expanding on the host and calling
 * grail_cuda_plan_create.  Phoneme ids follow the reference's enum order (:632-649): 0 Silence, 1 Stop, 2 Glide
 * (no sound: elem = None), then the sounds of make_phonemes! in declaration order (3 = A, 4 = E, ...);
 * storages[(s * n_sounds) + (id - 3)] is VoiceStorage field (id - 3) of voice s (:651-659), at the voice's sample
 * rate; utt_storage[u] picks the voice of utterance u (NULL: all use storage 0). */
int  grail_cuda_plan_create_phoneme_elems(grail_ctx* ctx, const grail_phoneme_elem* phonemes, const uint32_t* utt_offsets,
                                          const grail_elem* storages, uint32_t n_sounds, uint32_t n_storages,
                                          const uint32_t* utt_storage, const grail_voice_params* voices, uint32_t n_utts,
                                          grail_plan** out_plan);
/* bare ids: every phoneme gets length 0.5 s, blend 0.5 s and center_frequency[u] (Voice::center_frequency, already
 * divided by the sample rate) exactly as Intonator::next does today */
int  grail_cuda_plan_create_phonemes(grail_ctx* ctx, const uint8_t* phoneme_ids, const uint32_t* utt_offsets,
                                     const float* center_frequency, const grail_elem* storages, uint32_t n_sounds,
                                     uint32_t n_storages, const uint32_t* utt_storage, const grail_voice_params* voices,
                                     uint32_t n_utts, grail_plan** out_plan);
void grail_cuda_plan_destroy(grail_plan* plan);
uint64_t grail_cuda_plan_total_samples(const grail_plan* plan);
/* exact per-utterance offsets (n_utts+1 entries) of the plan's packed output */
int  grail_cuda_plan_out_offsets(const grail_plan* plan, uint64_t* out_offsets);
/* enqueue every kernel of the path on the ctx stream; d_out is a DEVICE pointer to
 * total_samples elements of `format`.  Asynchronous: pair with grail_cuda_synchronize. */
int  grail_cuda_plan_launch(grail_plan* plan, void* d_out, int format);
/* The same with every sample written `channels` times, interleaved: the examples' channel duplication
 * `flat_map(|x| repeat(x).take(num_channels))` (examples/cli.rs:229, interactive.rs:38) done on the device.
 * d_out holds total_samples * channels elements; utterance u starts at out_offsets[u] * channels. */
int  grail_cuda_plan_launch_interleaved(grail_plan* plan, void* d_out, int format, uint32_t channels);
/* With ctx option "pipeline" = 1 at plan creation (off by default, so that launches complete in stream order without
 * a join) consecutive launches of one plan are pipelined across the library's internal streams: launch k+1's
 * frequency / phase kernels run on a second scratch set under launch k's filter kernel (+4 % at config 2, same bits).  grail_cuda_plan_join makes the ctx's main stream (grail_cuda_stream_handle) wait, on the device,
 * for everything the plan has in flight -- call it before recording an event or enqueueing a consumer on that stream
 * (a no-op for unpipelined plans).  grail_cuda_synchronize, plan_read_output and plan_timings join implicitly. */
int  grail_cuda_plan_join(grail_plan* plan);
/* the plan's own device output buffer (allocated on first use), for callers with no allocator */
int  grail_cuda_plan_device_output(grail_plan* plan, int format, void** out_dptr);
/* D2H of the packed output into a host buffer (one cudaMemcpyAsync on the ctx stream, then a synchronize; pinned host
 * memory -- grail_cuda_host_alloc -- gets the full PCIe rate) */
int  grail_cuda_plan_read_output(grail_plan* plan, int format, void* host_out);
int  grail_cuda_plan_timings(const grail_plan* plan, grail_timings* out);
/* Long utterances (>= option "pscan_min_samples", "pscan_cost_model" (0|1), default 2^18) may get their carrier phase from the exact parallel
 * phase scan instead of the serial chain: the k <= 16 longest that minimise (scans + longest remaining chain) under
 * a measured cost model; option "pscan_cost_model" = 0 scans every utterance over the threshold (tests).  stats[4] = {scans in this plan, scans that converged (the rest fell back
 * to the serial chain), largest number of refinement rounds used, scans refused (an F_t outside [2^-16, 0.5])}
 * for the most recent launch. */
int  grail_cuda_plan_phase_scan_stats(grail_plan* plan, uint32_t* stats);
/* The carrier phase (src/lib.rs:520-525) is computed exactly AND in parallel over time chunks of option "phase_chunk"
 * samples (0 = chosen by the planner, 1024-4096; option "phase_mode" = 0 falls back to one serial chain per utterance
 * (with the phase scan above for long ones), 1 (default; 2 is an alias) = chunk-parallel always -- utterances of 2048
 * chunks or more have their chunk records scanned by a CTA of 32 warps instead of one warp --, 3 = chunk-parallel except
 * for plans of at most 16 utterances with one of >= "pscan_min_samples", which take the phase scan above (the default
 * until the CTA-wide scans: 3.3 ms against 0.92 ms for one 10-minute utterance, same bits)): every chunk is walked from a start phase derived from its
 * neighbours, and the result is accepted only when each chunk's end equals the next chunk's start bit for bit (a proof
 * by induction from the exact phase at sample 0); mismatching chunks are shifted and walked again for up to option
 * "phase_rounds" rounds, and an utterance that is still unproven then gets the serial chain.  stats[8] of the most
 * recent launch = {phase chunks in the plan, chunks walked in all rounds together, utterances that fell back to the
 * serial chain, last repair round that was needed (0 = none), chunk boundaries that failed a proof, phase_chunk, W, 0},
 * where W counts the (time chunk, formant) pairs of the filter kernel whose decay-bounded warm-up reached all the way
 * back to sample 0: a formant that rings longer than the utterance has lasted is recomputed from the start by every
 * later chunk (exact, but its time parallelism is gone -- a performance cliff that would otherwise be silent; 0 for
 * ordinary voices). */
int  grail_cuda_plan_phase_stats(grail_plan* plan, uint32_t* stats);
/* debug / parity taps, host buffers of total_samples entries; any may be NULL:
 * the bit-exact fundamental F_t, the carrier phase BEFORE each sample, and the polyBLEP saw */
int  grail_cuda_plan_read_intermediates(grail_plan* plan, float* frequency, float* carrier_phase, float* saw);

/* ---- streaming (unbounded input, examples/interactive.rs:31-38) -----------------------------
 * A grail_stream is the Copy-able state of the three iterators (Sequencer{cur,next,time},
 * Jitter{3 generators}, Synthesize{phase, a, b, c, seed}).  push appends
 * upstream SequenceElems; pull synthesizes up to max_samples more samples and returns how many
 * were produced (fewer only when the upstream ran dry: the last pushed element is held back as
 * `next` until another element or grail_cuda_stream_finish arrives).  `out` is a host buffer.  A stream pulled in
 * windows of any size reproduces the one-shot result: clocks, F_t and carrier phase bit for bit, the filters continue
 * from their exact states. */
int  grail_cuda_stream_new(grail_ctx* ctx, const grail_voice_params* voice, grail_stream** out_stream);
int  grail_cuda_stream_push(grail_stream* s, const grail_seq_elem* elems, uint32_t n_elems);
int  grail_cuda_stream_finish(grail_stream* s);
int  grail_cuda_stream_pull(grail_stream* s, float* out, uint64_t max_samples, uint64_t* n_written);
/* the next windows of n_streams streams of one ctx as ONE plan and one launch per kernel (a server's tick): the same
 * result per stream as n_streams separate pulls, for one launch / synchronisation latency instead of n_streams.
 * outs[k] receives up to max_samples[k] samples of streams[k]; n_written[k] how many it got. */
int  grail_cuda_streams_pull(grail_stream* const* streams, uint32_t n_streams, float* const* outs, const uint64_t* max_samples,
                             uint64_t* n_written);
void grail_cuda_stream_free(grail_stream* s);

/* ---- multi-GPU output gather helper ("optional gather of outputs to one rank") ----------------------------
 * Utterances are sharded over devices (one ctx each) and every shard's output is packed in ITS utterance order; after
 * the caller's all-gather (NCCL) this puts the pieces back in the order of the whole batch: n_segments copies
 * dst[dst_off[i] .. dst_off[i] + len[i]) = src[src_off[i] .. src_off[i] + len[i]), offsets and lengths in ELEMENTS of
 * elem_bytes (4: f32, 2: i16).  dst and src are DEVICE pointers on ctx's device; the three tables are HOST arrays.
 * One kernel launch on the ctx stream (asynchronous; segments must not overlap in dst). */
int  grail_cuda_copy_segments(grail_ctx* ctx, void* dst, const void* src, const uint64_t* dst_off, const uint64_t* src_off,
                              const uint64_t* len, uint64_t n_segments, uint32_t elem_bytes);

/* ---- roofline probes (used by bench.py; device microbenchmarks, not part of the path) -------- */
/* dense FFMA issue rate of this device, in FP32 flop/s (2 per FFMA), and MUFU.RCP rate in op/s */
int grail_cuda_probe_fp32_peak(grail_ctx* ctx, double* ffma_flops, double* mufu_ops, double* sm_mhz_effective);

#ifdef __cplusplus
}
#endif
#endif /* GRAIL_CUDA_H */
