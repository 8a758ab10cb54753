// grail_kernels.cuh -- the sm_100a kernels of the waveform-generation path.
//
//   k_jitter_schedule  exact wrap schedule of the value-noise phase clock      (src/lib.rs:242-251)
//   k_frequency        bit-exact per-sample fundamental F_t                     (src/lib.rs:861-931, 753-763)
//   k_phase_pair       bit-exact carrier phase + polyBLEP saw, two warps/utt    (src/lib.rs:503-525)
//   k_formant<NW,FPT>  noise, low-pass, turbulence, SVF band-pass, formant sum  (src/lib.rs:528-577, 764-773)
//
// Exactness classes: the clocks, the LCG streams, F_t, the carrier phase and the saw are
// computed with strict, never-contracted f32 ops in the reference's order and are bit-identical to it.
// Everything 8-wide and downstream of the saw is in the tolerance class (FMA, rcp.approx, re-association).
#pragma once
#include <cuda_runtime.h>

#include "grail_common.cuh"

namespace grail {

// ------------------------------------------------------------------------------------------------
// Device-resident plan
// ------------------------------------------------------------------------------------------------
constexpr int SEQ_WORDS = 52;      // sizeof(grail_seq_elem) / 4
constexpr int SE_HAS = 0;          // word index of has_elem
constexpr int SE_FREQ = 1;         // elem.frequency
constexpr int SE_ARR = 2;          // first array word: formant_freq[0]; array k at SE_ARR + 8k
constexpr int SE_LEN = 50;
constexpr int SE_BLEND = 51;
enum { P_FF = 0, P_BW = 1, P_SM = 2, P_BR = 3, P_TB = 4, P_AMP = 5 };

struct UttDev {
    uint32_t elem_first, n_elems;  // phonemes of this utterance in `elems` / `segs`
    uint32_t n_samples;
    uint32_t jit_sched;            // index into JitSchedDev
    uint32_t item_first, n_items;  // work items (time chunks) of this utterance: item_first + c * item_stride
    uint32_t n_active;             // number of formants whose amplitude is not identically zero
    uint8_t  active[8];            // their indices, ascending
    uint32_t item_stride;          // 1, or 32 when 32 equally long utterances are interleaved chunk by chunk
    uint64_t out_off;              // first output sample
    uint64_t f_off;                // first entry in the linear F_t scratch (multiple of 8)
    grail_voice_params voice;
    float    init_phase;           // carrier phase at sample 0 (0 for a fresh utterance)
    int32_t  pscan;                // >= 0: index of this utterance's exact parallel phase scan (long utterances)
    uint32_t jw0;                  // value-noise wraps that happened before sample 0 (continued streams)
    uint32_t has_init;             // filter states at sample 0 come from PlanDev::utt_init (continued streams)
    uint64_t sample0;              // absolute index of sample 0 in its stream (aspiration-noise draw index)
    uint32_t pc_first, pc_count;   // this utterance's phase chunks (k_phase_a / k_phase_b), pc_count = ceil(n_samples / phase_chunk)
};

// phoneme-level input of a plan: Selector and the Intonator stub run on the device
struct SelectDev {
    const grail_phoneme_elem* ph;  // PhonemeElem records, or
    const uint8_t* ids;            // bare phoneme ids with
    const float* center;           // one centre frequency per utterance
    const float* storages;         // [n_storages][n_sounds][49]
    const uint32_t* utt_storage;   // per utterance, null = 0
    uint32_t n_sounds, n_elems, n_utts;
};
constexpr uint32_t SELECT_WORDS = sizeof(grail_seq_elem) / 4;   // 52

struct JitSchedDev {
    float    inc;                  // voice.jitter_frequency
    float    phase0;               // value-noise phase before sample 0
    uint32_t n_max;                // samples the schedule must cover
    uint32_t rec_first, rec_cap;   // slice of the JitRec array
    uint32_t n_recs;               // written by k_jitter_schedule
    uint32_t overflow;
};

struct ItemDev {
    uint32_t utt;
    uint32_t n0;                   // first sample (multiple of chunk_len)
    uint32_t len;                  // samples in this chunk
    uint32_t pad;
};

struct DirtyRec;
struct PlanDev {
    const float*        elems;     // n_elems_total x 52 words
    const SegRec*       segs;      // one per phoneme
    const UttDev*       utts;
    const ItemDev*      items;
    JitSchedDev*        jscheds;
    JitRec*             jrecs;
    float*              F;         // F_t: linear, per utterance at f_off -- or, when f_tiled, in the saw's tiled layout
    float*              saw;       // tiled: [group][j/8][lane][8]
    float*              phase_dbg; // optional linear carrier phase tap (same indexing as F), may be null
    uint32_t*           fflags;    // one word per 128 F_t entries: nonzero if any is negative or NaN
    const float*        utt_init;  // per utterance 32 floats: a[8], b[8], c[8] at sample 0 (used when has_init)
    float*              utt_final; // per utterance 32 floats: a[8], b[8], c[8] after the last sample, [24] = carrier phase
    const uint32_t*     pscan_status; // 16 words per parallel phase scan: {mismatches, done, rounds, unsupported, history[12]}
    float*              pchunks;   // chunk-parallel exact phase: per-chunk records, one array per field (grail_phase.cuh; null: serial chains only)
    uint32_t            pc_stride; // elements per field array
    uint32_t*           pdirty;    // 2 x pc_stride chunk ids: the dirty chunks of the repair round in flight (by round parity)
    struct DirtyRec*    pdrec;     // pc_stride records: geometry of those chunks (k_phase_chain -> k_phase_saw)
    float*              ppark;     // [blocks per chunk][pc_stride]: phase at the start of every 8-sample block of a dirty chunk
    double*             bsum;      // sum of F_t over every 256-sample run (k_frequency), the guesses' raw material
    uint32_t*           utt_status;// per utterance: bit 0 = carrier phase proven exact by the chunk-parallel path
    uint32_t*           pstats;    // 64 words: see PSTAT_*
    uint32_t            phase_chunk; // PC, multiple of 256 (0: chunk-parallel path off)
    uint32_t            n_pchunks;
    uint32_t            pc_per_item; // K = ceil(chunk_len / phase_chunk): phase chunks per work item
    uint32_t            f_tiled;   // F is tiled like the saw (plans with the chunk-parallel phase)
    uint32_t*           err;       // device error word
    uint32_t n_utts, n_items, n_groups, n_jscheds;
    uint32_t chunk_len;            // CL, multiple of 256
    uint32_t out_channels;         // every sample is written this many times, interleaved (examples/cli.rs:229)
    float    warmup_nepers;
};

enum : uint32_t { DEV_ERR_JIT_OVERFLOW = 1u, DEV_ERR_SEG = 2u };

__device__ __forceinline__ size_t saw_index(uint32_t item, uint32_t j, uint32_t chunk_len)
{
    return ((size_t)(item >> 5) * (chunk_len >> 3) + (j >> 3)) * 256u + (item & 31u) * 8u + (j & 7u);
}

__device__ __forceinline__ float frcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Packed fp32 (sm_100: FFMA2 / FADD2 / FMUL2 on 64-bit register pairs): one issue slot for two IEEE operations, each
// half rounded exactly like its scalar counterpart.  k_formant is issue-bound and a lane's two formants run the very
// same instruction sequence on different data, so its inner loop carries them as (formant 0, formant 1) pairs.
// A scalar that both halves share is written pk(s, s): ptxas folds it into the instruction's broadcast operand form.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk(float lo, float hi)
{
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk(f2_t x, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c)
{
    f2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b)
{
    f2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2_t sub2(f2_t a, f2_t b)
{
    f2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b)
{
    f2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// largest index i in [0, n) with key(i) <= v; the caller guarantees key(0) <= v
template <typename F>
__device__ __forceinline__ uint32_t last_le(uint32_t n, F key, int64_t v)
{
    uint32_t lo = 0, hi = n; // invariant: key(lo) <= v, key(hi) > v (hi == n is +inf)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((int64_t)key(mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// Selector::next (src/lib.rs:987-1005): phoneme -> Option<SynthesisElem> from the voice storage, with
// copy_with_frequency (:445-450: frequency.min(0.5)); with bare ids also Intonator::next (:1057-1075: length 0.5,
// blend 0.5, the voice's centre frequency).  One thread per 32-bit word of the 208-byte Sequencer record.
__global__ void k_select(uint32_t* __restrict__ elems, const UttDev* __restrict__ utts, SelectDev S)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t p = (uint32_t)(t / SELECT_WORDS), w = (uint32_t)(t % SELECT_WORDS);
    if (p >= S.n_elems) return;
    uint32_t lo = 0, hi = S.n_utts;                          // last utterance whose first element is <= p
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (utts[mid].elem_first <= p) lo = mid; else hi = mid;
    }
    uint32_t id;
    float len, bl, fr;
    if (S.ph) {
        const grail_phoneme_elem e = S.ph[p];
        id = e.phoneme; len = e.length; bl = e.blend_length; fr = e.frequency;
    } else {
        id = S.ids[p]; len = 0.5f; bl = 0.5f; fr = S.center[lo];
    }
    const bool sound = id >= GRAIL_PHONEME_FIRST_SOUND;
    uint32_t v = 0;
    if (w == 0) v = sound ? 1u : 0u;
    else if (w == SELECT_WORDS - 2) v = __float_as_uint(len);
    else if (w == SELECT_WORDS - 1) v = __float_as_uint(bl);
    else if (sound) {
        if (w == 1) v = __float_as_uint(fminf(fr, 0.5f));
        else {
            const uint32_t st = S.utt_storage ? S.utt_storage[lo] : 0u;
            v = __float_as_uint(S.storages[((size_t)st * S.n_sounds + (id - GRAIL_PHONEME_FIRST_SOUND)) * 49u + (w - 1)]);
        }
    }
    elems[t] = v;
}


// ------------------------------------------------------------------------------------------------
// K0: jitter schedule.  One lane per distinct jitter_frequency; walks the value-noise phase clock
// binade by binade (closed form) and records every wrap.
// ------------------------------------------------------------------------------------------------
__global__ void k_jitter_schedule(PlanDev P)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_jscheds) return;
    JitSchedDev& S = P.jscheds[s];
    // The schedule is a function of (increment, length, start phase) alone -- part of the plan like the host planner's
    // per-phoneme records: a resident plan that is launched again finds it done (and an overflow re-raised).
    if (S.n_recs != 0) {
        if (S.overflow) atomicOr(P.err, DEV_ERR_JIT_OVERFLOW);
        return;
    }
    const uint32_t n = jitter_schedule_walk(S.inc, S.n_max, P.jrecs + S.rec_first, S.rec_cap, S.phase0);
    if (n == 0) {
        S.overflow = 1;
        S.n_recs = 1;
        atomicOr(P.err, DEV_ERR_JIT_OVERFLOW);
        return;
    }
    S.n_recs = n;
}

// ------------------------------------------------------------------------------------------------
// Segment view used by K1 (frequency lane only)
// ------------------------------------------------------------------------------------------------
struct FreqSeg {
    float xf, yf;   // blend endpoints: out = xf*(1-alpha) + yf*alpha
    float blend_len;
    float rcp_bl;   // RN(1 / blend_len) when the fast exact division applies, else 0
    int   silent;   // both cur and next have no element: SynthesisElem::silent(), frequency 0.25
};

// RN(a / b) for many a and one b.  With y = RN(1/b) (correctly rounded), q = RN(a y), r = a - b q (exact in an FMA),
// RN(q + r y) is the correctly rounded quotient (Markstein 1990) as long as nothing under- or overflows on the way;
// the guard keeps a and b far inside the normal range and everything else goes through the IEEE division.
__device__ __forceinline__ float div_by_const(float a, float b, float y)
{
    if (y != 0.0f && a > 1e-25f && a < 1e25f) {
        const float q = __fmul_rn(a, y);
        const float r = __fmaf_rn(-b, q, a);
        return __fmaf_rn(r, y, q);
    }
    return sdiv(a, b);
}
__device__ __forceinline__ float div_const_rcp(float b)
{
    return (b > 1e-12f && b < 1e12f) ? __frcp_rn(b) : 0.0f;
}

__device__ __forceinline__ FreqSeg load_freq_seg(const float* ue, uint32_t p, uint32_t n_elems)
{
    const float* cur = ue + (size_t)p * SEQ_WORDS;
    const bool b_on = __float_as_uint(cur[SE_HAS]) != 0u;
    bool c_on = false;
    const float* nxt = cur + SEQ_WORDS;
    if (p + 1 < n_elems) c_on = __float_as_uint(nxt[SE_HAS]) != 0u;
    FreqSeg s;
    s.blend_len = cur[SE_BLEND];
    s.rcp_bl = div_const_rcp(s.blend_len);
    s.silent = 0;
    if (b_on && c_on) { s.xf = nxt[SE_FREQ]; s.yf = cur[SE_FREQ]; }      // c.blend(b, alpha)          :902
    else if (b_on)    { s.xf = cur[SE_FREQ]; s.yf = cur[SE_FREQ]; }      // b.copy_silent().blend(b)   :911
    else if (c_on)    { s.xf = nxt[SE_FREQ]; s.yf = nxt[SE_FREQ]; }      // c.blend(c.copy_silent())   :920
    else              { s.xf = 0.25f; s.yf = 0.25f; s.silent = 1; }      // SynthesisElem::silent()    :926
    return s;
}

// ------------------------------------------------------------------------------------------------
// K1: bit-exact F_t.  One lane per run of FREQ_RUN consecutive samples: the two clocks are evaluated in
// closed form at the run start, then replayed literally; the scalar frequency path uses strict ops in
// the reference's order (blend :406, value-noise lerp :254, jitter add :763).
// ------------------------------------------------------------------------------------------------
// Samples per lane: `run_len`, a multiple of 256 chosen per plan.  The lane's start-up -- two binary searches, two
// closed-form clock fast-forwards, an LCG jump -- is paid once per run; with the straight-line quiet blocks it was more
// than half of the kernel's samples at 256 (ncu).  Measured: 0.344 / 0.295 / 0.312 ms at 256 / 512 / 1024 for config 2
// (one voice, interleaved chunks), but 0.478 / 0.648 / 0.605 ms on a config-4 slice (a voice per utterance, whole
// utterances as items): the planner takes 512 when the batch shares one jitter schedule and 256 otherwise.
constexpr int FREQ_RUN_MAX = 512;
#ifndef KFREQ_LANE_IS_ITEM
#define KFREQ_LANE_IS_ITEM 1
#endif

__global__ void __launch_bounds__(128) k_frequency(PlanDev P, uint32_t runs_per_item, uint32_t run_len)
{
    // Lane mapping.  Until F_t moved into the saw's tiled layout the lanes of a warp were consecutive runs of ONE item; ncu
    // then showed the kernel waiting on its own stores (the loop's back edge parked on the store scoreboard, 17 % of all
    // samples; lg_throttle): 32 lanes x 16 bytes into 32 different 1 KB tiles per instruction.  Measured at config 2:
    // 0.293 ms -> 0.247 ms with one 256-bit store per lane -> 0.212 ms with lane = item (config-4 slice 0.48 -> 0.37 ms).
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
#if KFREQ_LANE_IS_ITEM
    // a warp = the same run of the 32 items of one k_formant group: with F_t in the tiled layout its 32 lanes store
    // 1 KB of consecutive addresses per 8-sample block (one 256-bit store per lane), and with a shared voice the
    // value-noise wraps and hand-overs fall on the same samples in every lane
    const uint64_t wid = t >> 5;
    const uint32_t item = (uint32_t)(wid / runs_per_item) * 32u + (uint32_t)(t & 31u);
    const uint32_t run = (uint32_t)(wid % runs_per_item);
#else
    const uint32_t item = (uint32_t)(t / runs_per_item);
    const uint32_t run = (uint32_t)(t % runs_per_item);
#endif
    if (item >= P.n_items) return;
    const ItemDev it = P.items[item];
    const uint32_t off = run * run_len;
    if (off >= it.len) return;
    const uint32_t count = min(run_len, it.len - off);
    const uint32_t ns = it.n0 + off;
    const UttDev& U = P.utts[it.utt];
    const float* ue = P.elems + (size_t)U.elem_first * SEQ_WORDS;
    const SegRec* segs = P.segs + U.elem_first;
    const uint32_t n_elems = U.n_elems;
    const float dt = sdiv(1.0f, U.voice.sample_rate); // src/lib.rs:944
    const float jinc = U.voice.jitter_frequency;
    const float dfreq = U.voice.jitter_delta_frequency;

    // Sequencer state at sample ns
    uint32_t p = last_le(n_elems, [&](uint32_t i) { return segs[i].start; }, (int64_t)ns);
    float time = clock_desc_run(segs[p].time0, dt, ns - segs[p].start).x;
    FreqSeg seg = load_freq_seg(ue, p, n_elems);

    // value-noise state at sample ns
    const JitSchedDev& JS = P.jscheds[U.jit_sched];
    const JitRec* recs = P.jrecs + JS.rec_first;
    const uint32_t w = last_le(JS.n_recs, [&](uint32_t i) { return recs[i].n; }, (int64_t)ns);
    float jph = clock_asc_run(recs[w].phase, jinc, (uint64_t)((int64_t)ns - recs[w].n)).x;
    uint32_t s_next = lcg_jump(U.voice.jitter_seed, jit_freq_cur_idx((uint64_t)w + U.jw0));
    float cur = lcg_float(s_next);
    s_next = lcg_step(s_next);
    float nxt = lcg_float(s_next);

    // linear: one row per utterance; tiled: the saw's layout, 8-sample block b of this run 256 floats further each
    const bool tiled = P.f_tiled != 0u;
    float* dst = tiled ? P.F + saw_index(item, off, P.chunk_len) : P.F + U.f_off + ns;
    const uint32_t bstep = tiled ? 256u : 8u;   // distance between consecutive 8-sample blocks
    double fsum = 0.0;  // exact sum of this run's F_t (every term is a multiple of 2^-40 or so): k_phase_guess's raw material
    bool odd = false;   // any increment that is negative or NaN: k_phase then takes its fully general path
    // one sample of the scalar frequency path, strict ops in the reference's order
    auto freq_sample = [&]() -> float {
        float fb;
        if (seg.silent) {
            fb = 0.25f;
        } else {
            const float alpha = fminf(div_by_const(time, seg.blend_len, seg.rcp_bl), 1.0f);   // :899
            fb = sadd(smul(seg.xf, ssub(1.0f, alpha)), smul(seg.yf, alpha));               // :406
        }
        const float n0 = sadd(smul(cur, ssub(1.0f, jph)), smul(nxt, jph));                 // :254
        const float fr = sadd(fb, smul(n0, dfreq));                                        // :763
        odd |= !(fr >= 0.0f);
        fsum += (double)fr;
        return fr;
    };
    for (uint32_t k0 = 0; k0 < count; k0 += 8) {
        if (k0 != 0 && (k0 & 127u) == 0) {   // one flag word per 128 samples (f_off, n0 and the run start are 128-aligned)
            P.fflags[(U.f_off + ns + k0 - 128) >> 7] = odd ? 1u : 0u;
            odd = false;
            if ((k0 & 255u) == 0 && P.bsum) {   // one sum per 256 samples
                P.bsum[(U.f_off + ns + k0 - 256) >> 8] = fsum;
                fsum = 0.0;
            }
        }
        const bool quiet = (k0 + 8 <= count) && (time > 9.0f * dt) && (jph + 9.0f * jinc < 1.0f);
        if (quiet) { // no hand-over and no wrap inside these 8 samples: literal clocks, no event tests
            float buf[8];
            float bsum8 = 0.0f;   // the block's sum in f32 first (8 terms of ~2^-9: the error is far below 2^-23), one f64 add per block
            // the segment is fixed for the block and `time` only falls: the guards of div_by_const and the silent case
            // are decided once per block, the 8 samples are then straight-line code
            if (!seg.silent && seg.rcp_bl != 0.0f && time < 1e25f) {          // (time - 8 dt > dt > 1e-25 here)
                const float y = seg.rcp_bl, nb = -seg.blend_len;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float q = __fmul_rn(time, y);
                    const float alpha = fminf(__fmaf_rn(__fmaf_rn(nb, q, time), y, q), 1.0f);          // :899 (exact quotient)
                    const float fb = sadd(smul(seg.xf, ssub(1.0f, alpha)), smul(seg.yf, alpha));    // :406
                    const float n0 = sadd(smul(cur, ssub(1.0f, jph)), smul(nxt, jph));              // :254
                    const float fr = sadd(fb, smul(n0, dfreq));                                     // :763
                    odd |= !(fr >= 0.0f);
                    bsum8 = sadd(bsum8, fr);
                    buf[k] = fr;
                    time = ssub(time, dt);                                                          // :861
                    jph = sadd(jph, jinc);                                                          // :242
                }
                fsum += (double)bsum8;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    buf[k] = freq_sample();
                    time = ssub(time, dt);                                                     // :861
                    jph = sadd(jph, jinc);                                                     // :242
                }
            }
            float* d8 = dst + (k0 >> 3) * bstep;
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d8), "f"(buf[0]), "f"(buf[1]), "f"(buf[2]), "f"(buf[3]),
                         "f"(buf[4]), "f"(buf[5]), "f"(buf[6]), "f"(buf[7])
                         : "memory");
        } else {
            const uint32_t kend = min(k0 + 8, count);
#pragma unroll 1
            for (uint32_t k = k0; k < kend; ++k) {
                dst[(k >> 3) * bstep + (k & 7u)] = freq_sample();
                time = ssub(time, dt);                                                     // :861
                if (time < 0.0f) {                                                         // :864
                    ++p;
                    if (p < n_elems) {
                        time = sadd(time, ue[(size_t)p * SEQ_WORDS + SE_LEN]);             // :873
                        seg = load_freq_seg(ue, p, n_elems);
                    }
                }
                jph = sadd(jph, jinc);                                                     // :242
                if (jph > 1.0f) {                                                          // :245
                    jph = ssub(jph, 1.0f);
                    cur = nxt;
                    s_next = lcg_step(s_next);
                    nxt = lcg_float(s_next);
                }
            }
        }
    }
    P.fflags[(U.f_off + ns + ((count - 1) & ~127u)) >> 7] = odd ? 1u : 0u;   // the last (possibly partial) 128-block
    if (P.bsum) P.bsum[(U.f_off + ns + ((count - 1) & ~255u)) >> 8] = fsum;   // the last (possibly partial) 256-run
    if (it.n0 + it.len == U.n_samples && off + count == it.len) {
        // the lane that wrote the utterance's last sample rounds the row up: k_phase_pair copies F_t in groups of 8
        // and the flag words in pairs, so nothing it touches is left unwritten (the values themselves are unused)
        const uint32_t n = U.n_samples;
        for (uint32_t i = n; i < ((n + 7u) & ~7u); ++i) {
            const uint32_t k = i - ns;             // same 8-sample block as the last sample
            dst[(k >> 3) * bstep + (k & 7u)] = 0.0f;
        }
        if ((((n - 1u) >> 7) & 1u) == 0u) P.fflags[((U.f_off + n - 1u) >> 7) + 1u] = 0u;
    }
}

// test hook: div_by_const against the IEEE division on pseudo-random pairs (a spread over ~60 binades around the
// Sequencer's operating range, b over the blend lengths the guard admits); counts the pairs that differ in any bit
__global__ void k_debug_div_check(uint32_t seed, uint32_t per_thread, unsigned long long* mismatches)
{
    uint32_t s = seed ^ ((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u);
    unsigned long long bad = 0;
    for (uint32_t i = 0; i < per_thread; ++i) {
        s = lcg_step(s); const uint32_t ma = s;
        s = lcg_step(s); const uint32_t mb = s;
        s = lcg_step(s); const uint32_t ex = s;
        // mantissas from the draws, exponents: a in 2^[-40, 20), b in 2^[-30, 30)
        const float a = __uint_as_float((ma >> 9) | ((127u - 40u + (ex >> 8) % 60u) << 23));
        const float b = __uint_as_float((mb >> 9) | ((127u - 30u + (ex >> 20) % 60u) << 23));
        const float y = div_const_rcp(b);
        bad += __float_as_uint(div_by_const(a, b, y)) != __float_as_uint(sdiv(a, b));
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ------------------------------------------------------------------------------------------------
// K2 (serial-chain form): bit-exact carrier phase and polyBLEP saw.  The f32 chain
// phase <- RN(phase + F_t) is the only truly serial dependency of the path; blocks of 8 are run
// speculatively as bare adds (4 cycles each) and redone carefully only when a wrap falls inside.
// Output goes straight into the tiled layout k_formant reads with perfectly coalesced 128-bit loads.
// ------------------------------------------------------------------------------------------------
// CTA = 4 utterances x 2 warps.  Warps 0-3 each own one utterance's chain (one per SM sub-partition):
// F_t tiles of 256 samples stream through an 8-deep cp.async ring in shared memory (the prefetch distance
// hides HBM latency), every lane walks the same chain from broadcast LDS reads and lane 0 parks the phase at
// the start of every 8-sample block in a double-buffered shared tile -- nothing else is on the chain warp's
// instruction stream, because a lone warp issues only about one instruction every four cycles.  Warps 4-7
// take each finished tile: lane l replays block l's 8 steps from its start phase (exactly, with the
// reference's wrap test), forms the polyBLEP saw and writes one 32-byte sector, while the chain warp is
// already on the next tile.  Producer and consumer meet on mbarriers (full / empty per buffer).
constexpr int PH_TILE = 256, PH_STAGES = 8, PH_AHEAD = 6, PH_UTTS = 4;

__device__ __forceinline__ void cp_async16(unsigned sa, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned a)
{
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}

// explicit shared-window accesses (32-bit addresses): keeps generic-address arithmetic out of the chain loop
__device__ __forceinline__ float4 lds128(unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts32(unsigned a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
__device__ __forceinline__ unsigned lds32u(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void cp_async8(unsigned sa, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}

// the rare edge samples of the saw (one before and one after each carrier wrap), kept out of line
__device__ __noinline__ float saw_edge(float phase, float f)
{
    float polyblep = 0.0f;
    if (phase < f) {                                             // :503-506
        const float t = sdiv(phase, f);
        polyblep = ssub(ssub(smul(2.0f, t), smul(t, t)), 1.0f);
    } else if (phase > ssub(1.0f, f)) {                          // :507-510
        const float t = sdiv(ssub(phase, 1.0f), f);
        polyblep = sadd(sadd(smul(t, t), smul(2.0f, t)), 1.0f);
    }
    return ssub(ssub(smul(2.0f, phase), 1.0f), polyblep);        // :517
}

// (x >= 1) ? 1.0f : 0.0f as a float-valued compare: keeps the wrap off the predicate path, whose 13-cycle latency
// would otherwise sit on every step of the literal chain (a negative or NaN x gives 0)
__device__ __forceinline__ float ge_one(float x)
{
    float r;
    asm("set.ge.f32.f32 %0, %1, 0f3F800000;" : "=f"(r) : "f"(x));
    return r;
}

// 8 steps exactly as the reference takes them (any increment, any wrap)  :520-525
__device__ __forceinline__ float phase_steps8(float phase, const float4& fa, const float4& fb, uint32_t valid)
{
    const float fv[8] = { fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w };
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if ((uint32_t)k < valid) {
            phase = sadd(phase, fv[k]);
            phase = ssub(phase, ge_one(phase));     // if phase >= 1 { phase -= 1 }: x - 0.0 is x, bit for bit (x >= +0 or NaN)
        }
    }
    return phase;
}

// One 32-sample quad whose speculative chain reached 1.0: a carrier wrap lies inside (rare: kept out of line so
// the chain warp's straight-line code stays small).  The speculative values are exact up to the wrap, so the
// 8-sample blocks before the first block whose END value reached 1.0 keep their start phases; that block is
// stepped exactly as the reference does, and the blocks after it are chained again from the corrected phase.
// ps[b] = speculative phase at the start of block b, ps[4] = at the end of the quad.
__device__ __noinline__ float phase_redo_quad(float ps0, float ps1, float ps2, float ps3, float ps4, unsigned f_a,
                                              unsigned p_a, bool lane0)
{
    // all 32 increments up front: eight loads in flight while the wrap block is located
    float4 f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = lds128(f_a + i * 16);
    // blocks 0..b-1 end below 1.0 (the chain is monotone): their start phases stand; block b holds the wrap
    const int b = (ps1 < 1.0f ? 1 : 0) + (ps2 < 1.0f ? 1 : 0) + (ps3 < 1.0f ? 1 : 0);
    float phase, s1 = ps1, s2 = ps2, s3 = ps3;
    // one plain block after the wrap: speculate again, literal steps only if a second wrap follows at once
    auto next_block = [&](float ph, const float4& a, const float4& c) -> float {
        const float p4 = sadd(sadd(sadd(sadd(ph, a.x), a.y), a.z), a.w);
        const float p8 = sadd(sadd(sadd(sadd(p4, c.x), c.y), c.z), c.w);
        return (p8 < 1.0f) ? p8 : phase_steps8(ph, a, c, 8);
    };
    if (b == 0) {
        s1 = phase_steps8(ps0, f[0], f[1], 8);
        s2 = next_block(s1, f[2], f[3]);
        s3 = next_block(s2, f[4], f[5]);
        phase = next_block(s3, f[6], f[7]);
    } else if (b == 1) {
        s2 = phase_steps8(ps1, f[2], f[3], 8);
        s3 = next_block(s2, f[4], f[5]);
        phase = next_block(s3, f[6], f[7]);
    } else if (b == 2) {
        s3 = phase_steps8(ps2, f[4], f[5], 8);
        phase = next_block(s3, f[6], f[7]);
    } else {
        phase = phase_steps8(ps3, f[6], f[7], 8);
    }
    if (lane0) sts128(p_a, ps0, s1, s2, s3);
    (void)ps4;
    return phase;
}

// LEAN: the few-register build for batches with more utterances than two CTAs per SM can hold (see the redo call)
template <bool LEAN>
__global__ void __launch_bounds__(PH_UTTS * 64) k_phase_pair(PlanDev P)
{
    __shared__ __align__(16) float sF[PH_UTTS][PH_STAGES][PH_TILE];
    __shared__ __align__(16) float sP[PH_UTTS][2][32];          // phase at the start of each 8-sample block
    __shared__ __align__(8) uint64_t s_full[PH_UTTS][2], s_empty[PH_UTTS][2];
    __shared__ __align__(8) uint32_t sFlag[PH_UTTS][PH_STAGES][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp & (PH_UTTS - 1);
    const bool is_chain = warp < PH_UTTS;
    if (threadIdx.x < PH_UTTS * 2) {
        mbar_init(&s_full[threadIdx.x >> 1][threadIdx.x & 1], 1);
        mbar_init(&s_empty[threadIdx.x >> 1][threadIdx.x & 1], 1);
    }
    __syncthreads();
    const uint32_t u = blockIdx.x * PH_UTTS + slot;
    if (u >= P.n_utts) return;
    const UttDev& U = P.utts[u];
    const uint32_t n = U.n_samples;
    if (n == 0) return;
    if (U.pscan >= 0 && P.pscan_status[16 * U.pscan + 1] != 0u) return;   // the exact parallel scan already did it
    if (P.pchunks && (P.utt_status[u] & 1u)) return;                       // the chunk-parallel path proved it (grail_phase.cuh)
    const uint32_t ntiles = (n + PH_TILE - 1) / PH_TILE;
    const unsigned full_a = (unsigned)__cvta_generic_to_shared(&s_full[slot][0]);     // + 8 * buf
    const unsigned empty_a = (unsigned)__cvta_generic_to_shared(&s_empty[slot][0]);

    if (is_chain) {
        // ---------------- chain warp ----------------
        const float* src = P.F + U.f_off;
        const bool tiled = P.f_tiled != 0u;                  // (plans of the chunk-parallel phase: this kernel is their fallback)
        const uint32_t CLt = P.chunk_len;
        const uint32_t* flags = P.fflags + (U.f_off >> 7);   // f_off is a multiple of 256: 8-byte aligned pairs
        const uint32_t npad = (n + 7u) & ~7u;
        const unsigned sF_a = (unsigned)__cvta_generic_to_shared(&sF[slot][0][0]);
        const unsigned sP_a = (unsigned)__cvta_generic_to_shared(&sP[slot][0][0]);
        const unsigned sG_a = (unsigned)__cvta_generic_to_shared(&sFlag[slot][0][0]);
        auto issue = [&](uint32_t tile) {
            if (tile < ntiles) {
                const uint32_t off = tile * PH_TILE + lane * 8;
                const unsigned dst = sF_a + ((tile % PH_STAGES) * PH_TILE + lane * 8) * 4;
                if (off < npad) {
                    const float* g8 = src + off;
                    if (tiled) {
                        const uint32_t ci = off / CLt;
                        g8 = P.F + saw_index(U.item_first + ci * U.item_stride, off - ci * CLt, CLt);
                    }
                    cp_async16(dst, g8);
                    cp_async16(dst + 16, g8 + 4);
                }
                // the tile's two sign-flag words ride the same ring (the flag array is padded past the end)
                if (lane == 0) cp_async8(sG_a + (tile % PH_STAGES) * 8, flags + 2 * tile);
            }
            cp_async_commit();
        };
        for (uint32_t t = 0; t < PH_AHEAD; ++t) issue(t);
        float phase = U.init_phase;
        const bool lane0 = lane == 0;
        for (uint32_t tile = 0; tile < ntiles; ++tile) {
            const int buf = tile & 1;
            // the saw warp has consumed tile-2: its phase buffer AND its F stage (the one refilled next) are free
            mbar_wait(empty_a + buf * 8, ((tile >> 1) & 1) ^ 1);   // first use of each buffer passes at once
            issue(tile + PH_AHEAD);             // into stage (tile-2) % 8
            cp_async_wait<PH_AHEAD>();
            __syncwarp();
            const unsigned fa_ = sF_a + (tile % PH_STAGES) * (PH_TILE * 4);   // this tile's F_t
            const unsigned pa_ = sP_a + buf * (32 * 4);                        // block-start phases out
            const uint32_t odd = lds32u(sG_a + (tile % PH_STAGES) * 8) | lds32u(sG_a + (tile % PH_STAGES) * 8 + 4);
            const uint32_t left = n - tile * PH_TILE;
            const uint32_t nblk = min(32u, (left + 7u) >> 3);
            if (odd || left < PH_TILE) {
                // a tile with a negative / NaN increment somewhere, or the ragged last tile: fully general path
#pragma unroll 1
                for (uint32_t blk = 0; blk < nblk; ++blk) {
                    const float4 fa = lds128(fa_ + blk * 32), fb = lds128(fa_ + blk * 32 + 16);
                    if (lane0) sts32(pa_ + blk * 4, phase);
                    phase = phase_steps8(phase, fa, fb, min(8u, left - blk * 8));
                }
            } else {
                // main path: the tile's 256 samples as straight-line code, 32 at a time: a chain of 32 dependent
                // adds (the critical path of the whole path) plus 8 broadcast loads and one store.  A lone warp
                // pays ~20 cycles for every taken branch, so there is no loop here and the only branch (the rare
                // redo) is forward.  Increments are non-negative (k_frequency's tile flag), so the chain is
                // monotone and its last value bounds the rest: p32 < 1 proves no wrap happened.
                float4 fv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) fv[i] = lds128(fa_ + i * 16);
#pragma unroll
                for (uint32_t q = 0; q < 8; ++q) {
                    float4 fn[8];                       // next quad's increments, loaded in the shadow of this chain
                    if (q < 7) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) fn[i] = lds128(fa_ + (q + 1) * 128 + i * 16);
                    }
                    float p = phase;
                    float ps[4];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if ((i & 1) == 0) ps[i >> 1] = p;
                        p = sadd(sadd(sadd(sadd(p, fv[i].x), fv[i].y), fv[i].z), fv[i].w);
                    }
                    if (lane0) sts128(pa_ + q * 16, ps[0], ps[1], ps[2], ps[3]);
                    if (__builtin_expect(p < 1.0f, 1)) {
                        phase = p;
                    } else {
                        phase = phase_redo_quad(ps[0], ps[1], ps[2], ps[3], p, fa_ + q * 128, pa_ + q * 16, lane0);
                        // LEAN: the prefetched increments are loaded again after the call, so that nothing but a
                        // handful of scalars is live across it and the callee's registers do not add to the kernel's:
                        // 59 registers and 4 CTAs per SM instead of 99 and 2.  Worth 16 % when the kernel is
                        // throughput-bound (4 096 short utterances: 1.95 -> 1.63 ms), but the tighter schedule costs
                        // the latency-bound case 10 % (1 024 long utterances: 1.17 -> 1.28 ms), hence two builds.
                        if (LEAN && q < 7) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) fn[i] = lds128(fa_ + (q + 1) * 128 + i * 16);
                        }
                    }
                    if (q < 7) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) fv[i] = fn[i];
                    }
                }
            }
            __syncwarp();
            if (lane0) mbar_arrive(full_a + buf * 8);   // release: the tile's phases (and its F stage) are ready
        }
        if (lane0) P.utt_final[(size_t)u * 32 + 24] = phase;   // Synthesize.phase after the last sample (stream state)
    } else {
        // ---------------- saw warp ----------------
        float* dbg = P.phase_dbg ? P.phase_dbg + U.f_off : nullptr;
        const uint32_t CL = P.chunk_len;
        // destination of this lane's 32-byte sector: tiles advance 256 samples and CL is a multiple of 256, so the
        // (chunk, offset) pair is tracked incrementally instead of divided out every tile
        uint32_t dst_item = U.item_first, dst_j = lane * 8;
        for (uint32_t tile = 0; tile < ntiles; ++tile) {
            const int buf = tile & 1;
            mbar_wait(full_a + buf * 8, (tile >> 1) & 1);
            const float* f = sF[slot][tile % PH_STAGES];
            const uint32_t b0 = tile * PH_TILE + lane * 8;
            float4 fa, fb;
            float p = 0.0f;
            if (b0 < n) {
                fa = *reinterpret_cast<const float4*>(f + lane * 8);
                fb = *reinterpret_cast<const float4*>(f + lane * 8 + 4);
                p = sP[slot][buf][lane];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_a + buf * 8);   // both shared tiles are in registers now
            if (b0 < n) {
                const float fv[8] = { fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w };
                const uint32_t valid = min(8u, n - b0);
                float pv[8], s[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {   // replay this block's 8 steps from its start phase  :520-525
                    pv[k] = p;
                    p = sadd(p, fv[k]);
                    if (p >= 1.0f) p = ssub(p, 1.0f);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    s[k] = fmaf(2.0f, pv[k], -1.0f);                                    // :517 with polyblep = 0 (2p is exact)
                    const bool edge = !((pv[k] >= fv[k]) && (pv[k] <= ssub(1.0f, fv[k])));
                    if (edge) s[k] = saw_edge(pv[k], fv[k]);
                    if ((uint32_t)k >= valid) s[k] = 0.0f;
                }
                float4* dst = reinterpret_cast<float4*>(P.saw + saw_index(dst_item, dst_j, CL));
                dst[0] = make_float4(s[0], s[1], s[2], s[3]);
                dst[1] = make_float4(s[4], s[5], s[6], s[7]);
                if (dbg) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if ((uint32_t)k < valid) dbg[b0 + k] = pv[k];
                }
            }
            dst_j += PH_TILE;
            if (dst_j >= CL) { dst_j -= CL; dst_item += U.item_stride; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2 (parallel form, for long utterances): the exact carrier phase without the serial chain.
//
// f32 addition is "add exactly, then round to the grid of the result's binade", so in fixed point (units of
// 2^-40 cycle; every phase and every F_t >= 2^-16 is an integer there) one step is
//     P' = RHE(P + Fi, 2^g) mod 2^40,    g = max(0, msb(P + Fi) - 23)        (RHE = round half to even)
// The only thing that keeps it from being a prefix sum is g (and the parity bit of an exact tie), which depend
// on P itself -- but only weakly: an ESTIMATE of P that is off by a few ulps still classifies almost every
// step correctly.  Fixed-point iteration:
//   round 0   estimate P~ = exact prefix sum of Fi (no rounding at all)
//   each round  classify every step from P~ (grid g, "wraps", tie parity); GIVEN those, the rounded increment of a
//             step depends only on the low 17 bits L of P, and L restarts from 0 after every wrap -- so each
//             wrap-to-wrap segment is replayed independently (one lane per 256-sample block replays the segments
//             that start in it), giving the rounded increments; an exact integer prefix sum of them is the new P~
//   stop      when a fully parallel check with REAL f32 ops, step(P_t, F_t) == P_{t+1} for every t, passes:
//             by induction from P_0 the trajectory is then the reference's, bit for bit.
// Converges in 1-5 rounds (measured); if it has not after PS_MAX_ROUNDS, or some F_t is outside [2^-16, 0.5],
// the done flag stays 0 and k_phase_pair runs the serial chain for that utterance instead.
// ------------------------------------------------------------------------------------------------
constexpr int PS_MAX_ROUNDS = 12;
constexpr int PS_BLOCK = 256;          // samples per replay lane
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long PS_ONE = 1ull << 40, PS_MASK = PS_ONE - 1ull;

struct PScanDev {
    const float* F;                 // the utterance's F_t
    unsigned long long* P;          // n + 1 fixed-point phases (estimate, then result)
    unsigned long long* inc;        // n rounded increments
    unsigned long long* bsum;       // scan spine
    unsigned char* cls;             // per step, from the current estimate: grid exponent g (bits 0-4), wraps (bit 5)
    unsigned char* sflag;           // per step: bit0 parity of the carry into the high part, bit1 tie on a wrap step, bit2 parity of c at the tie, bit3 the tie is currently rounded up
    unsigned char* bpar;            // per 256-step block: parity transducer (bit1 has_tie, bit0 xor)
    unsigned char* bpin;            // per block: parity entering it (k_ps_parity_spine)
    uint32_t* stamp;                // per block: the last round whose replay must redo it (its step classes changed)
    uint32_t* lbk;                  // per block: first block of its look-back range when it was last replayed
    uint32_t* status;               // {mismatches, done, rounds, unsupported}
    uint32_t n;
    unsigned long long p0;          // phase at sample 0
};

__device__ __forceinline__ unsigned long long ps_fix(float f) { return (unsigned long long)(f * 1099511627776.0f); } // * 2^40, exact
__device__ __forceinline__ int ps_grid(unsigned long long x) { const int m = 63 - __clzll((long long)(x | 1ull)); return m > 23 ? m - 23 : 0; }
// round x to a multiple of 2^g, half to even; `odd_hint` overrides the parity bit when it is not inside x
__device__ __forceinline__ unsigned long long ps_rhe(unsigned long long x, int g, int parity_override /* -1: use x */)
{
    if (g == 0) return x;
    const unsigned long long unit = 1ull << g, rem = x & (unit - 1ull), base = x - rem, half = unit >> 1;
    const unsigned long long odd = parity_override < 0 ? ((base >> g) & 1ull) : (unsigned long long)parity_override;
    return (rem > half || (rem == half && odd)) ? base + unit : base;
}

__global__ void k_ps_init(PScanDev S)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.n) return;                        // (status words are zeroed by the host before this launch)
    const float f = S.F[t];
    if (!(f >= 1.52587890625e-05f && f <= 0.5f)) atomicOr(S.status + 3, 1u);   // outside [2^-16, 0.5]: not this path
    S.inc[t] = ps_fix(f);
}

// exclusive prefix sum of inc (mod 2^64, which 2^40 divides) -> P[t] = (p0 + sum_{s<t} inc_s) mod 2^40, t = 0..n
__global__ void __launch_bounds__(SCAN_THREADS) k_ps_scan_reduce(PScanDev S)
{
    if (S.status[1] | S.status[3]) return;
    __shared__ unsigned long long sh[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < S.n) s += S.inc[base + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        for (int i = 0; i < SCAN_THREADS / 32; ++i) tot += sh[i];
        S.bsum[blockIdx.x] = tot;
    }
}
__global__ void __launch_bounds__(1024) k_ps_scan_spine(PScanDev S, uint32_t nb)
{
    if (S.status[1] | S.status[3]) return;
    __shared__ unsigned long long sh[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const unsigned long long v = i < nb ? S.bsum[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long w = sh[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            sh[threadIdx.x] = w;
        }
        __syncthreads();
        const unsigned long long wprev = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0ull;
        const unsigned long long incl = x + wprev + carry;
        if (i < nb) S.bsum[i] = incl - v;            // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
}
// Third scan pass, fused with what every step needs from the new estimate: (a) the proof -- the step redone with
// real f32 operations must land on the next phase (counted into status[0] when `verify`), and (b) the step's class
// for the next round's replay: rounding grid and "reaches 1.0".
__global__ void __launch_bounds__(SCAN_THREADS) k_ps_scan_apply(PScanDev S, int verify, uint32_t next_round)
{
    if (S.status[1] | S.status[3]) return;
    __shared__ unsigned long long sh[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned long long v[SCAN_ITEMS], s = 0;
    float f[SCAN_ITEMS];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < S.n) ? S.inc[base + i] : 0ull;
        f[i] = (base + i < S.n) ? __ldg(S.F + base + i) : 0.25f;
        s += v[i];
    }
    unsigned long long x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    unsigned long long woff = 0;
    for (int i = 0; i < (int)(threadIdx.x >> 5); ++i) woff += sh[i];
    unsigned long long run = S.p0 + S.bsum[blockIdx.x] + woff + (x - s);   // exclusive prefix of this lane's first item
    unsigned bad = 0;
    unsigned long long cpack = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const unsigned long long p = run & PS_MASK;
        if (base + i <= S.n) S.P[base + i] = p;                            // P[n] (the final phase) included
        run += v[i];
        if (base + i < S.n) {
            const unsigned long long q = run & PS_MASK;
            const float a = __ull2float_rn(p) * 9.094947017729282e-13f;    // * 2^-40
            const float c = __ull2float_rn(q) * 9.094947017729282e-13f;
            bool b = ((unsigned long long)(a * 1099511627776.0f) != p) || ((unsigned long long)(c * 1099511627776.0f) != q);
            float nx = sadd(a, f[i]);                                      // :520
            if (nx >= 1.0f) nx = ssub(nx, 1.0f);                           // :523-525
            b |= __float_as_uint(nx) != __float_as_uint(c);
            bad += b ? 1u : 0u;
            const unsigned long long xx = p + ps_fix(f[i]);
            const int g = ps_grid(xx);
            cpack |= (unsigned long long)(unsigned)(g | ((ps_rhe(xx, g, -1) >= PS_ONE) ? 32 : 0)) << (8 * i);
        }
    }
    static_assert(SCAN_ITEMS == 8, "class bytes are packed into one 8-byte store");
    // a block whose classes changed (all of them after round 0) is stamped for the next round's replay; the others
    // keep their increments, which depend on nothing else (k_ps_replay checks the stamps of its look-back range too)
    if (base + SCAN_ITEMS <= S.n) {
        unsigned long long* cp = reinterpret_cast<unsigned long long*>(S.cls + base);
        if (!verify || *cp != cpack) S.stamp[base / PS_BLOCK] = next_round;
        *cp = cpack;
    } else {
        for (int i = 0; i < SCAN_ITEMS; ++i)
            if (base + i < S.n) {
                const unsigned char c = (unsigned char)(cpack >> (8 * i));
                if (!verify || S.cls[base + i] != c) S.stamp[(base + i) / PS_BLOCK] = next_round;
                S.cls[base + i] = c;
            }
    }
    if (verify) {
        const unsigned m = __reduce_add_sync(0xffffffffu, bad);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(S.status, m);
    }
}

// One lane per 256-sample block.  The low bits L restart from 0 after every wrap, so the lane first looks BACK for
// the last step before its block that wraps under the current estimate (a whole carrier period at most), replays
// forward from there without writing, and then replays its own block writing the rounded increments and the
// parity flags.  Grids and wraps come from k_ps_classify.
constexpr uint32_t PS_MAX_LOOKBACK = 1u << 16;

__global__ void __launch_bounds__(128) k_ps_replay(PScanDev S, uint32_t round)
{
    if (S.status[1] | S.status[3]) return;                       // uniform over the grid
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t t0_ = (uint64_t)b * PS_BLOCK;
    const bool live = t0_ < S.n;                                 // dead lanes stay with the warp: no divergent exits
    const uint32_t t0 = live ? (uint32_t)t0_ : 0u;
    const uint32_t t1 = live ? (uint32_t)min((uint64_t)S.n, t0_ + PS_BLOCK) : 0u;
    // Nothing this lane reads has changed since it last ran (the classes of its block and of its look-back range, whose
    // first block it noted then): its increments, flags and block parity map stand.  After the second round this is
    // true of almost every block.  (A range that changed in an earlier round was redone, and re-noted, in that round.)
    bool todo = false;
    if (live) {
        if (round == 1u) todo = true;
        else for (uint32_t bb = S.lbk[b]; bb <= b; ++bb) todo |= S.stamp[bb] == round;
    }
    if (!__any_sync(0xffffffffu, todo)) return;
    // Look back for the segment start, 4 class bytes at a time (t0 is a multiple of 256, so s0 stays 4-aligned).
    uint32_t s0 = t0, back = 0;
    uint32_t w4 = 0;
    while (todo && s0 >= 4 && back <= PS_MAX_LOOKBACK) {
        w4 = __ldg(reinterpret_cast<const uint32_t*>(S.cls + s0 - 4)) & 0x20202020u;
        if (w4) break;
        s0 -= 4;
        back += 4;
    }
    __syncwarp();
    if (back > PS_MAX_LOOKBACK) atomicOr(S.status + 3, 1u);      // a carrier slower than 0.7 Hz: serial chain instead
    if (w4) s0 = s0 - 4 + ((31 - __clz(w4)) >> 3) + 1;           // first step after the last wrap
    if (todo) S.lbk[b] = (s0 ? s0 - 1u : 0u) / PS_BLOCK;          // the wrap step that ends the look-back is part of what was read
    unsigned long long L = (s0 == 0) ? (S.p0 & 0x1FFFFull) : 0ull;
    // one step: the low bits decide the rounding.  Only an exact tie on a wrap step (g == 17, low 17 bits == 2^16)
    // needs a bit from above them: the parity of the high part.  Those steps are rounded DOWN here and flagged;
    // k_ps_parity_* resolves them with an exact parity scan (common when the pitch is 0.25).
    auto step = [&](float f, unsigned c, unsigned long long& inc, unsigned& flag) {
        const unsigned long long fi = ps_fix(f);
        const int g = (int)(c & 31u);
        const unsigned long long xl = L + fi;
        const bool tie = (g == 17) && ((xl & 0x1FFFFull) == 0x10000ull);
        const unsigned long long r = tie ? (xl & ~0x1FFFFull) : ps_rhe(xl, g, -1);
        inc = r - L;
        flag = tie ? (2u | (unsigned)(((xl >> 17) & 1ull) << 2)) : (unsigned)((r >> 17) & 1ull);
        L = (c & 32u) ? 0ull : (r & 0x1FFFFull);   // after a wrap the phase is a multiple of 2^-23: no low bits
    };
    unsigned long long inc;
    unsigned flag;
    for (uint32_t s = s0; todo && s < t0; ++s) step(__ldg(S.F + s), __ldg(S.cls + s), inc, flag);
    __syncwarp();
    // Parity of the high part Q = floor(P / 2^17): a non-tie step adds a known carry (bit 0 of its flag); a tie step
    // rounds to even, so Q is even after it whatever came before.  Per block the parity map is either p -> p ^ x or
    // the constant x, which composes associatively (k_ps_parity_spine scans it over the blocks).
    unsigned px = 0, has_tie = 0;
    auto parity = [&](unsigned fl) { if (fl & 2u) { has_tie = 1; px = 0; } else px ^= fl & 1u; };
    // own block, four steps at a time: 16 bytes of F and 4 class bytes in, one full 32-byte sector of increments
    // and 4 flag bytes out
    uint32_t s = todo ? t0 : t1;
    for (; s + 4 <= t1; s += 4) {
        const float4 f4 = __ldg(reinterpret_cast<const float4*>(S.F + s));
        const uint32_t c4 = __ldg(reinterpret_cast<const uint32_t*>(S.cls + s));
        unsigned long long i0, i1, i2, i3;
        unsigned g0, g1, g2, g3;
        step(f4.x, c4 & 0xFFu, i0, g0);
        step(f4.y, (c4 >> 8) & 0xFFu, i1, g1);
        step(f4.z, (c4 >> 16) & 0xFFu, i2, g2);
        step(f4.w, c4 >> 24, i3, g3);
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(S.inc + s);
        dst[0] = make_ulonglong2(i0, i1);
        dst[1] = make_ulonglong2(i2, i3);
        *reinterpret_cast<uint32_t*>(S.sflag + s) = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
        parity(g0); parity(g1); parity(g2); parity(g3);
    }
    for (; s < t1; ++s) {
        step(__ldg(S.F + s), __ldg(S.cls + s), inc, flag);
        S.inc[s] = inc;
        S.sflag[s] = (unsigned char)flag;
        parity(flag);
    }
    if (todo) S.bpar[b] = (unsigned char)((has_tie << 1) | px);
    if (has_tie) atomicMax(S.status + 15, b + 1u);   // 1 + the last block that holds a tie (0: none so far; sticky)
}

// exclusive scan of the block transducers -> incoming parity of every block (one CTA walks the spine)
__global__ void __launch_bounds__(1024) k_ps_parity_spine(PScanDev S, uint32_t nblk)
{
    if (S.status[1] | S.status[3] | (S.status[15] == 0u)) return;    // no tie anywhere so far: nothing to fix
    nblk = min(nblk, S.status[15]);                                  // parities past the last tie block are never read
    __shared__ unsigned sh[32];
    __shared__ unsigned carry;                       // parity entering the current stripe
    if (threadIdx.x == 0) carry = (unsigned)((S.p0 >> 17) & 1ull);
    __syncthreads();
    auto compose = [](unsigned a, unsigned b) -> unsigned { return (b & 2u) ? b : ((a & 2u) | ((a ^ b) & 1u)); };   // a then b
    for (uint32_t b0 = 0; b0 < nblk; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const unsigned v = i < nblk ? (unsigned)S.bpar[i] : 0u;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x = compose(y, x);
        }
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = sh[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w = compose(y, w);
            }
            sh[threadIdx.x] = w;
        }
        __syncthreads();
        // inclusive map from the stripe start to the end of this block, then apply it to the stripe's incoming parity
        unsigned incl = x;
        if (threadIdx.x >> 5) incl = compose(sh[(threadIdx.x >> 5) - 1], x);
        // exclusive: the map up to (not including) this block = shifted inclusive
        const unsigned prev = __shfl_up_sync(0xffffffffu, incl, 1);
        unsigned excl;
        if ((threadIdx.x & 31) == 0) excl = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0u;   // identity = (no tie, xor 0)
        else excl = prev;
        const unsigned pin = (excl & 2u) ? (excl & 1u) : ((carry ^ excl) & 1u);
        const unsigned pout = (incl & 2u) ? (incl & 1u) : ((carry ^ incl) & 1u);
        __syncthreads();
        if (i < nblk) S.bpin[i] = (unsigned char)pin;
        if (threadIdx.x == 1023) carry = pout;
        __syncthreads();
    }
}
// second walk with the incoming parity known: round every flagged tie to even
__global__ void __launch_bounds__(128) k_ps_parity_fix(PScanDev S)
{
    if (S.status[1] | S.status[3]) return;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t t0 = (uint64_t)b * PS_BLOCK;
    if (t0 >= S.n || b >= S.status[15] || !(S.bpar[b] & 2u)) return;   // only blocks that hold a tie have anything to fix
    const uint32_t t1 = (uint32_t)min((uint64_t)S.n, t0 + PS_BLOCK);
    unsigned p = S.bpin[b] & 1u;
    for (uint32_t t = (uint32_t)t0; t < t1; ++t) {
        const unsigned f = S.sflag[t];
        if (f & 2u) {
            // (Q + c) odd: round half UP to the even multiple.  Bit 3 remembers that this tie is currently rounded up:
            // blocks the replay skipped this round still carry last round's decision, which the new incoming parity
            // may confirm or reverse.
            const unsigned want = (p ^ (f >> 2)) & 1u, have = (f >> 3) & 1u;
            if (want != have) {
                S.inc[t] += want ? 0x20000ull : ~0x20000ull + 1ull;
                S.sflag[t] = (unsigned char)(f ^ 8u);
            }
            p = 0;
        } else {
            p ^= f & 1u;
        }
    }
}

// end of a round: status[0] holds the steps whose f32 redo missed the next phase (counted by k_ps_scan_apply)
__global__ void k_ps_check(PScanDev S)
{
    if (S.status[1] | S.status[3]) return;
    if (S.status[2] < 11u) S.status[4 + S.status[2]] = S.status[0];   // mismatch history (diagnostics)
    S.status[2] += 1;
    if (S.status[0] == 0) S.status[1] = 1;      // converged: exact
    S.status[0] = 0;
}

// phases -> polyBLEP saw in the tiled layout k_formant reads (same arithmetic as k_phase_pair's saw warp)
__global__ void k_ps_saw(PScanDev S, PlanDev P, uint32_t utt)
{
    if (!S.status[1]) return;                   // not converged: k_phase_pair does this utterance
    const uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;   // 8 samples per lane
    const uint64_t b0 = (uint64_t)blk * 8;
    if (blk == 0) P.utt_final[(size_t)utt * 32 + 24] = __ull2float_rn(S.P[S.n]) * 9.094947017729282e-13f;
    if (b0 >= S.n) return;
    const UttDev& U = P.utts[utt];
    const uint32_t valid = (uint32_t)min((uint64_t)8, S.n - b0);
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        s[k] = 0.0f;
        if ((uint32_t)k < valid) {
            const float ph = __ull2float_rn(S.P[b0 + k]) * 9.094947017729282e-13f;
            const float f = S.F[b0 + k];
            s[k] = ssub(smul(2.0f, ph), 1.0f);
            if (!((ph >= f) && (ph <= ssub(1.0f, f)))) s[k] = saw_edge(ph, f);
            if (P.phase_dbg) P.phase_dbg[U.f_off + b0 + k] = ph;
        }
    }
    const uint32_t CL = P.chunk_len;
    const uint32_t item = U.item_first + (uint32_t)(b0 / CL) * U.item_stride, j = (uint32_t)(b0 % CL);
    float4* dst = reinterpret_cast<float4*>(P.saw + saw_index(item, j, CL));
    dst[0] = make_float4(s[0], s[1], s[2], s[3]);
    dst[1] = make_float4(s[4], s[5], s[6], s[7]);
}

// ------------------------------------------------------------------------------------------------
// K3: the dominant kernel.  CTA = one group of 32 work items (time chunks) x NW warps; warp w owns FPT
// consecutive active formants of each item's utterance (they share the clocks, the noise and the saw), lane l
// owns chunk l.  Each lane replays the exact
// clocks literally, generates its formant's parameters, noise, low-pass and SVF in registers, and
// leaves v1 in a shared tile; every 32 samples the CTA sums the tile over formants (reference order)
// and writes 128-byte rows.  Filter state at a chunk start comes from a warm-up over the preceding
// samples, long enough that the zero-state error has decayed by exp(-warmup_nepers).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void seg_endpoints(const float* ue, uint32_t p, uint32_t n_elems, int fi, float* x, float* y,
                                              float* blend_len)
{
    const float* cur = ue + (size_t)p * SEQ_WORDS;
    const float* nxt = cur + SEQ_WORDS;
    const bool b_on = __float_as_uint(cur[SE_HAS]) != 0u;
    const bool c_on = (p + 1 < n_elems) && (__float_as_uint(nxt[SE_HAS]) != 0u);
    *blend_len = cur[SE_BLEND];
    if (!b_on && !c_on) { // SynthesisElem::silent() :367-377
        x[P_FF] = y[P_FF] = 0.25f; x[P_BW] = y[P_BW] = 0.25f; x[P_SM] = y[P_SM] = 0.25f;
        x[P_BR] = y[P_BR] = 0.0f;  x[P_TB] = y[P_TB] = 0.0f;  x[P_AMP] = y[P_AMP] = 0.0f;
        return;
    }
    const float* xs = c_on ? nxt : cur; // blend "self"
    const float* ys = b_on ? cur : nxt; // blend "other"
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        x[k] = xs[SE_ARR + 8 * k + fi];
        y[k] = ys[SE_ARR + 8 * k + fi];
    }
    if (!c_on) x[P_AMP] = 0.0f; // b.copy_silent().blend(b, alpha)   :911
    if (!b_on) y[P_AMP] = 0.0f; // c.blend(c.copy_silent(), alpha)   :920
}

// tan(pi x) approximation of src/lib.rs:63-70, as numerator / denominator
__device__ __forceinline__ void tan_nd(float x, float* num, float* den)
{
    const float p = fmaf(-x, x, x);      // (1 - x) x
    const float q = fmaf(-x, x, 0.25f);  // (x + .5)(.5 - x)
    *num = p * fmaf(-4.0f, q, 5.0f);
    *den = q * fmaf(-4.0f, p, 5.0f);
}

// slowest natural decay (nepers per sample) of one formant's low-pass + SVF pair
__device__ float decay_rate(float ff, float bw, float sm)
{
    float num, den;
    tan_nd(ff, &num, &den);
    const float g = num / den, k = bw / ff;
    const float delta = 1.0f + g * (g + k);
    const float det = (1.0f - g * k + g * g) / delta;   // product of the two SVF poles
    const float htr = (1.0f - g * g) / delta;           // half trace
    const float disc = htr * htr - det;
    float rho = disc > 0.0f ? fabsf(htr) + sqrtf(disc) : sqrtf(fmaxf(det, 0.0f));
    const float o = 1.0f - sm;
    const float a5 = fabsf(o * o * o * o * o);          // low-pass pole :535
    rho = fmaxf(rho, a5);
    if (!(rho < 1.0f)) return 0.0f;                     // no decay (or NaN): warm up from sample 0
    return -logf(fmaxf(rho, 1e-30f));
}

__device__ float seg_decay_rate(const float* ue, uint32_t p, uint32_t n_elems, int fi, float dff)
{
    float x[6], y[6], bl;
    seg_endpoints(ue, p, n_elems, fi, x, y, &bl);
    float r = decay_rate(x[P_FF] - dff, x[P_BW], x[P_SM]);
    r = fminf(r, decay_rate(x[P_FF] + dff, x[P_BW], x[P_SM]));
    r = fminf(r, decay_rate(y[P_FF] - dff, y[P_BW], y[P_SM]));
    r = fminf(r, decay_rate(y[P_FF] + dff, y[P_BW], y[P_SM]));
    r = fminf(r, decay_rate(x[P_FF] - dff, y[P_BW], x[P_SM]));
    r = fminf(r, decay_rate(y[P_FF] + dff, x[P_BW], y[P_SM]));
    return r;
}

// samples of history needed before n0 so that a zero-state start has decayed by exp(-need)
__device__ uint32_t warmup_len(const float* ue, const SegRec* segs, uint32_t n_elems, uint32_t n0, int fi, float dff,
                               float need)
{
    uint32_t p = last_le(n_elems, [&](uint32_t i) { return segs[i].start; }, (int64_t)n0 - 1);
    float acc = 0.0f;
    uint32_t hi = n0;
    for (;;) {
        const uint32_t lo = segs[p].start;
        const float rate = seg_decay_rate(ue, p, n_elems, fi, fabsf(dff));
        const float span = (float)(hi - lo);
        if (rate > 0.0f && acc + rate * span >= need) {
            const float extra = ceilf((need - acc) / rate);
            uint32_t w = (n0 - hi) + (uint32_t)fminf(extra, span);
            w = (w + 7u) & ~7u;
            return min(w, n0);
        }
        acc += rate * span;
        hi = lo;
        if (p == 0) return n0;
        --p;
    }
}

// LCG^8 as compile-time constants (the 8 formant draws of one value-noise refill are 8 apart, src/lib.rs:301)
constexpr uint32_t lcg_pow_a(int n) { uint32_t a = 1; for (int i = 0; i < n; ++i) a *= LCG_A; return a; }
constexpr uint32_t lcg_pow_c(int n) { uint32_t c = 0; for (int i = 0; i < n; ++i) c = c * LCG_A + LCG_C; return c; }
constexpr uint32_t LCG8_A = lcg_pow_a(8), LCG8_C = lcg_pow_c(8);

// One formant of one lane.  Every per-sample parameter is one or two FMAs of the two slowly varying scalars
//   alpha (Sequencer blend position, src/lib.rs:899) and jph (value-noise phase, :291):
//   formant_freq = ff0 + alpha*ff1 + jph*ff2       (blend :407 + jitter :764, the jitter lerp folded in)
//   formant_amp  = (am0 + alpha*am1) * (aj0 + jph*aj1)                          (blend :412 + jitter :768-773)
//   1 - smooth, bw, breath, turb = x0 + alpha*x1                                        (blend :408-411)
// The folded constants change only at a phoneme hand-over or a value-noise wrap.
struct FormantLane {
    float ff0, ff1, ff2, xff;
    float bw0, bw1, om0, om1, br0, br1, tb0, tb1, am0, am1;
    float aj0, aj1;
    uint32_t s_ff, s_amp;   // LCG states of the "next" draws of formant_freq_noise[fi] / formant_amp_noise[fi]
    float a, b, c;          // low-pass state, SVF ic1eq, ic2eq
    int fi;                 // formant index, -1 if this slot is unused
};

// samples between two CTA-wide reductions of the partial sums (a power of two, at least 32)
#ifndef KF_TILE
#define KF_TILE 32
#endif
#ifndef KF_ALG2
#define KF_ALG2 0
#endif
#ifndef KF_RING
#define KF_RING 0      // measured slower (see k_formant): kept as a build option for the record
#endif
template <int FPT> struct FormantCfg { static constexpr int warps_per_sm = 24; };   // 80 registers per lane
#ifndef KF_WARPS2
#define KF_WARPS2 12
#endif
template <> struct FormantCfg<2> { static constexpr int warps_per_sm = KF_WARPS2; };       // 12: up to 168 registers per lane

template <int NW, int FPT>
__global__ void __launch_bounds__(NW * 32, FormantCfg<FPT>::warps_per_sm / NW)
k_formant(PlanDev P, void* __restrict__ out, int format)
{
    // partial sums: [formant group][chunk row][32 samples], rows padded to 36 floats so that both the per-lane
    // 128-bit row writes and the 8-lanes-per-row 128-bit reads of the reduction are bank-conflict free
    // (static shared memory ends at 48 KB: instantiations with many formant groups fall back to 32-sample tiles)
    constexpr int KT = (NW * 32 * (KF_TILE + 4) * 4 + 1024 <= 49152) ? KF_TILE : 32;
    // KF_RING (a build option, OFF: measured 1.66 ms against 1.45 ms at config 2 and 5.06 against 4.17 ms on a config-4
    // slice -- the reducer warp alone is slower than all warps sharing the rows between two barriers).
    // In instantiations with 2-4 warps the CTA-wide reduction of the partial sums is taken off the common path.
    // The LAST warp is the reducer: the others (producers) only park their partial sums in a two-deep ring of tiles
    // and signal an mbarrier per 32-sample batch; the reducer waits for them, adds the tiles in formant order and
    // stores.  No __syncthreads in the loop, and the warm-up imbalance between the warps (the first warp holds the
    // formants that ring longest: 3 590 against 1 450 samples of warm-up with the default voice) pays for the
    // reducer's extra work instead of being waited out at a barrier.
    constexpr bool RING = (KF_RING != 0) && NW >= 2 && NW <= 4 && KT == 32;
    constexpr int RB = RING ? 2 : 1;
    __shared__ __align__(16) float part_all[RB][NW][32][KT + 4];
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
    float (*part)[32][KT + 4] = part_all[0];
    __shared__ unsigned long long row_out[32];
    __shared__ uint32_t row_len[32];

    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const uint32_t item_id = blockIdx.x * 32u + lane;
    const bool have = item_id < P.n_items;
    ItemDev it;
    it.utt = 0; it.n0 = 0; it.len = 0; it.pad = 0;
    if (have) it = P.items[item_id];
    const UttDev& U = P.utts[it.utt];
    if (w == 0) {
        row_out[lane] = U.out_off + it.n0;
        row_len[lane] = it.len;
    }
    const float* ue = P.elems + (size_t)U.elem_first * SEQ_WORDS;
    const SegRec* segs = P.segs + U.elem_first;
    const uint32_t n_elems = U.n_elems;
    const uint32_t CL = P.chunk_len;

    const float dt = sdiv(1.0f, U.voice.sample_rate);
    const float ndt = -dt;
    const float jinc = U.voice.jitter_frequency;
    const float dff = U.voice.jitter_delta_formant_frequency;
    const float hda = 0.5f * U.voice.jitter_delta_amplitude;   // :769
    const float quiet_t = 17.0f * dt, quiet_j = 1.0f - 17.0f * jinc;   // no hand-over / wrap within the next 16 samples
    const float quiet_j8 = 1.0f - 9.0f * jinc;                         // no wrap within the next 8
    const uint32_t jseed = U.voice.jitter_seed;

    const bool has_init = U.has_init != 0u;     // a continued stream (read once: the loop below must not reload it)
    FormantLane L[FPT];
    bool on = false;
#pragma unroll
    for (int j = 0; j < FPT; ++j) {
        const uint32_t slot = (uint32_t)(w * FPT + j);
        L[j].fi = (have && it.len > 0 && slot < U.n_active) ? (int)U.active[slot] : -1;
        on |= L[j].fi >= 0;
        // an unused slot runs SynthesisElem::silent() parameters with zero amplitude: it contributes exactly 0
        L[j].ff0 = 0.25f; L[j].ff1 = 0.f; L[j].ff2 = 0.f; L[j].xff = 0.25f;
        L[j].bw0 = 0.25f; L[j].bw1 = 0.f; L[j].om0 = 0.75f; L[j].om1 = 0.f;
        L[j].br0 = L[j].br1 = L[j].tb0 = L[j].tb1 = L[j].am0 = L[j].am1 = 0.f;
        L[j].aj0 = 1.f; L[j].aj1 = 0.f;
        L[j].s_ff = L[j].s_amp = 0u;
        L[j].a = L[j].b = L[j].c = 0.f;
    }

    // ---- warm-up depth: per formant slot, maximised over the warp so the warp walks the same rows.  The slot with the
    // longer warm-up goes first: the other one (FPT = 2) joins only `w1` rows before the chunk, so the rows before
    // that run one formant instead of two (formant 0 of the default voice needs 3 590 samples, formant 1 1 600).
    uint32_t wslot[FPT];
#pragma unroll
    for (int j = 0; j < FPT; ++j) {
        uint32_t wl = 0;
        if (on && it.n0 > 0 && L[j].fi >= 0) {
            wl = warmup_len(ue, segs, n_elems, it.n0, L[j].fi, dff, P.warmup_nepers);
            // a formant that rings longer than the utterance has lasted so far: this lane recomputes it from sample 0
            // (exact, but the time parallelism is gone for it) -- counted, so that the cliff shows in the plan's stats
            if (wl >= it.n0 && P.pstats) atomicAdd(P.pstats + 8, 1u);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
        wslot[j] = (wl + 15u) & ~15u;             // whole 16-sample interpolation blocks
    }
    if (FPT == 2 && wslot[FPT - 1] > wslot[0]) {  // warp-uniform
        const int f = L[0].fi; L[0].fi = L[FPT - 1].fi; L[FPT - 1].fi = f;
        const uint32_t t = wslot[0]; wslot[0] = wslot[FPT - 1]; wslot[FPT - 1] = t;
    }
    const uint32_t wmax = wslot[0];
    const bool keep_slot1 = has_init && wslot[FPT - 1] >= it.n0;   // the second slot starts at the window's sample 0 with the carried state
    const int r_join = (FPT == 2) ? -(int)wslot[FPT - 1] : (int)0x80000000;   // first row of the second slot
    const uint32_t wmine = min(wmax, it.n0);      // multiple of 16 (n0 is a multiple of 256)
    const uint32_t ns = it.n0 - wmine;            // first sample this lane computes

    // ---- shared (per lane) clocks and noise at sample ns
    float time = 0.f, jph = 0.f, inv_bl = 0.f;
    uint32_t p = 0, jw = 0, s_noise = 0;

    // (re)load the blend endpoints of phoneme p for every formant of this lane  (:891-931)
    auto load_segment = [&]() {
        float bl = 1.0f;
#pragma unroll
        for (int j = 0; j < FPT; ++j) {
            if (L[j].fi < 0) continue;
            float x[6], y[6];
            seg_endpoints(ue, p, n_elems, L[j].fi, x, y, &bl);
            L[j].ff0 += x[P_FF] - L[j].xff;            // keep the folded jitter term, swap the base
            L[j].xff = x[P_FF];
            L[j].ff1 = y[P_FF] - x[P_FF];
            L[j].bw0 = x[P_BW]; L[j].bw1 = y[P_BW] - x[P_BW];
            L[j].om0 = 1.0f - x[P_SM]; L[j].om1 = x[P_SM] - y[P_SM];
            L[j].br0 = x[P_BR]; L[j].br1 = y[P_BR] - x[P_BR];
            L[j].tb0 = x[P_TB]; L[j].tb1 = y[P_TB] - x[P_TB];
            L[j].am0 = x[P_AMP]; L[j].am1 = y[P_AMP] - x[P_AMP];
        }
        inv_bl = 1.0f / bl;
    };
    // fold the value-noise (current, next) pairs into the per-sample constants  (:289-306, :764-773)
    auto fold_jitter = [&](int j, float ffc, float ffn, float amc, float amn) {
        L[j].ff0 = fmaf(dff, ffc, L[j].xff);
        L[j].ff2 = dff * (ffn - ffc);
        L[j].aj0 = (1.0f - hda) - hda * amc;           // 1 - (n + 1) * (0.5 * delta_amplitude)
        L[j].aj1 = -hda * (amn - amc);
    };

    if (on) {
        p = last_le(n_elems, [&](uint32_t i) { return segs[i].start; }, (int64_t)ns);
        time = clock_desc_run(segs[p].time0, dt, ns - segs[p].start).x;
        const JitSchedDev& JS = P.jscheds[U.jit_sched];
        const JitRec* recs = P.jrecs + JS.rec_first;
        jw = last_le(JS.n_recs, [&](uint32_t i) { return recs[i].n; }, (int64_t)ns);
        jph = clock_asc_run(recs[jw].phase, jinc, (uint64_t)((int64_t)ns - recs[jw].n)).x;
        jw += U.jw0;                                    // wraps since the Jitter was built, not since sample 0
        s_noise = lcg_jump(U.voice.synth_seed, U.sample0 + ns);   // noise of stream sample n is draw n+1  (:528)
#pragma unroll
        for (int j = 0; j < FPT; ++j) {
            if (L[j].fi < 0) continue;
            const float c0 = lcg_float(lcg_jump(jseed, jit_arr_cur_idx(0, L[j].fi, jw)));
            L[j].s_ff = lcg_jump(jseed, jit_arr_next_idx(0, L[j].fi, jw));
            const float c1 = lcg_float(lcg_jump(jseed, jit_arr_cur_idx(1, L[j].fi, jw)));
            L[j].s_amp = lcg_jump(jseed, jit_arr_next_idx(1, L[j].fi, jw));
            L[j].xff = 0.f;
            fold_jitter(j, c0, lcg_float(L[j].s_ff), c1, lcg_float(L[j].s_amp));
            L[j].xff = 0.f; L[j].ff0 = dff * c0;      // base added by load_segment below
            // A continued stream: the Synthesize filter states carry over.  They apply to every slot whose run starts at
            // the window's sample 0 -- the first chunk, and any later chunk whose warm-up reaches back to sample 0 (there
            // a start from rest would ignore the carried state; its error decays only with the warm-up actually available).
            if (has_init && wslot[j] >= it.n0) {
                const float* st0 = P.utt_init + (size_t)it.utt * 32;
                L[j].a = st0[L[j].fi]; L[j].b = st0[8 + L[j].fi]; L[j].c = st0[16 + L[j].fi];
            }
        }
        load_segment();
    }

    // Per-sample coefficients of one formant at clock values (alpha, jph): the Sequencer blend (:404-414), the
    // Jitter perturbation (:764-773), exp_approx (:75-82) and the SVF coefficients (:555-562).
    // lp = 1 - exp_approx(smooth); turbulence and amplitude are folded: v0 = a * (amp0 + amp1 * noise) with
    // amp0 = amp (1 - turb), amp1 = amp turb   (1 * (1 - turb) + noise * turb, times amp: :544-550)
    struct Coef { float a1, g, lp, amp0, amp1, br; };
    auto coeffs = [&](const FormantLane& F, float alpha, float jp) -> Coef {
        Coef c;
        const float ff = fmaf(jp, F.ff2, fmaf(alpha, F.ff1, F.ff0));
        const float bw = fmaf(alpha, F.bw1, F.bw0);
        const float o = fmaf(alpha, F.om1, F.om0);                            // 1 - smooth
        c.br = fmaf(alpha, F.br1, F.br0);
        const float tb = fmaf(alpha, F.tb1, F.tb0);
        const float amp = fmaf(alpha, F.am1, F.am0) * fmaf(jp, F.aj1, F.aj0);
        c.amp1 = amp * tb;
        c.amp0 = amp - c.amp1;
        const float o2 = o * o;
        c.lp = fmaf(-o2 * o2, o, 1.0f);
        float num, den;
        tan_nd(ff, &num, &den);
        const float g = num * frcp(den);
        const float kq = bw * frcp(ff);
        c.a1 = frcp(fmaf(g, g + kq, 1.0f));
        c.g = g;
        return c;
    };
    // One filter step of one formant (:531-571): breath mix, one-pole low-pass, turbulence, amplitude, SVF tick.
    auto tick = [&](FormantLane& F, const Coef& c, float saw, float d1, float nz) -> float {
        const float nw = fmaf(c.br, d1, saw);                                 // :531
        F.a = fmaf(c.lp, nw - F.a, F.a);                                      // :538
        const float v0 = F.a * fmaf(c.amp1, nz, c.amp0);                      // :544-550
        const float v3 = v0 - F.c;                                            // :565
        const float v1 = c.a1 * fmaf(c.g, v3, F.b);                           // a1 b + a2 v3 with a2 = g a1  (:561, :566)
        const float v2 = fmaf(c.g, v1, F.c);                                  // c + a2 b + a3 v3 with a3 = g a2  (:562, :567)
        F.b = fmaf(2.0f, v1, -F.b);
        F.c = fmaf(2.0f, v2, -F.c);
        return v1;
    };
    // aspiration noise of the next sample (:528) and the two differences every formant uses
    auto noise = [&](float saw, float& d1, float& nz) {
        s_noise = s_noise * LCG_A + LCG_C;                                    // :40
        nz = fmaf(__uint_as_float((s_noise >> 9) | 0x3F800000u), 2.0f, -3.0f);
        d1 = nz - saw;
    };
    // one sample with exact per-sample coefficients: all formants of this lane; returns their v1 sum (:566)
    auto sample = [&](float saw) -> float {
        const float alpha = fminf(time * inv_bl, 1.0f);                       // :899
        float d1, nz;
        noise(saw, d1, nz);
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < FPT; ++j) acc += tick(L[j], coeffs(L[j], alpha, jph), saw, d1, nz);
        return acc;
    };

    // value-noise wrap: current <- next, draw the new next, refold  (:294-301).  Small, so it is inlined after
    // every sample of the unrolled block; lanes hit it at different samples (one lane in ~2756 per sample).
    auto jitter_wrap = [&]() {
        jph = __fadd_rn(jph, -1.0f);
        ++jw;
#pragma unroll
        for (int j = 0; j < FPT; ++j) {
            if (L[j].fi < 0) continue;
            const float nf = lcg_float(L[j].s_ff), na_ = lcg_float(L[j].s_amp);   // old next becomes current
            if (jw == 1) {
                L[j].s_ff = lcg_jump(jseed, jit_arr_next_idx(0, L[j].fi, 1));
                L[j].s_amp = lcg_jump(jseed, jit_arr_next_idx(1, L[j].fi, 1));
            } else {
                L[j].s_ff = LCG8_A * L[j].s_ff + LCG8_C;                     // 8 draws per refill :301
                L[j].s_amp = LCG8_A * L[j].s_amp + LCG8_C;
            }
            fold_jitter(j, nf, lcg_float(L[j].s_ff), na_, lcg_float(L[j].s_amp));
        }
    };
    // phoneme hand-over (:864-888): rare (once per phoneme per lane) and large, handled in the per-sample loop
    auto handover = [&]() {
        ++p;
        if (p < n_elems) {
            time = __fadd_rn(time, ue[(size_t)p * SEQ_WORDS + SE_LEN]);      // :873
            load_segment();
        }
    };

    if (RING && threadIdx.x == 0) {
        mbar_init(&s_full[0], NW - 1); mbar_init(&s_full[1], NW - 1);
        mbar_init(&s_empty[0], 1); mbar_init(&s_empty[1], 1);
    }
    __syncthreads();   // row_out / row_len (and the barriers) visible
    uint32_t lmax = 0;
#pragma unroll 1
    for (int r = 0; r < 32; ++r) lmax = max(lmax, row_len[r]);
    lmax = (lmax + (uint32_t)(KT - 1)) & ~(uint32_t)(KT - 1);      // whole KT-sample batches: the reduction runs at the end of each

    // saw source: walks the tiled layout from sample ns; crossing into the next chunk of the utterance (and,
    // at r == 0, into this lane's own chunk) is a pointer reset every CL samples
    const uint32_t item_stride = U.item_stride;
    uint32_t src_item = U.item_first + (ns / CL) * item_stride, src_j = ns % CL;
    const float4* sp = reinterpret_cast<const float4*>(P.saw + saw_index(src_item, src_j, CL));
    // saw values are fetched one iteration ahead (register double buffer) so the L2 latency of the
    // coalesced 128-bit loads is covered by a whole block of arithmetic
    float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na;
    auto fetch = [&]() {
        na = __ldg(sp);
        nb = __ldg(sp + 1);
        src_j += 8;
        if (src_j == CL) {
            src_j = 0;
            src_item += item_stride;
            sp = reinterpret_cast<const float4*>(P.saw + saw_index(src_item, 0, CL));
        } else {
            sp += 64;
        }
    };
    // this lane computes r in [r_lo, r_hi): its warm-up then its chunk (empty for an idle lane)
    const int r_lo = on ? -(int)wmine : 0x7fffffff;
    const int r_hi = on ? (int)it.len : (int)0x80000000;
    if (-(int)wmax >= r_lo && -(int)wmax < r_hi) fetch();
    unsigned part_a = (unsigned)__cvta_generic_to_shared(&part_all[0][w][lane][0]);   // (RING: moves to the batch's tile)
    const unsigned part_a0 = part_a;
    constexpr unsigned SLOT_BYTES = (unsigned)(NW * 32 * (KT + 4) * 4);
    const unsigned full_a = (unsigned)__cvta_generic_to_shared(&s_full[0]), empty_a = (unsigned)__cvta_generic_to_shared(&s_empty[0]);

    Coef cend[FPT];        // coefficients at the current clock values (the next block's start point)
    bool c_valid = false;
#pragma unroll
    for (int j = 0; j < FPT; ++j) cend[j] = Coef{ 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };

    // ---- one loop over [-wmax, lmax) in steps of 8: warm-up (r < 0, no output) then the chunk itself.
    // Blocks of 16 samples (two iterations, half = 0 / 1) are classified once, at their first half.
    // hand: a phoneme hand-over may fall inside (per lane, rare: generic per-sample loop).
    // exact: a value-noise wrap or the alpha clip falls inside, so the parameters have a kink in this block; if
    // ANY lane of the warp is in that class the whole warp takes the exact per-sample path (a uniform branch).
    // Otherwise every parameter is linear in time over the block and the 6 per-sample coefficients are
    // interpolated between the block's end points (error ~ h^2/8 c'' < 1e-7 relative, the size of f32 rounding),
    // which removes the blend / tan_approx / reciprocal work from 15 of every 16 samples.
    // (the ragged last block of a chunk also goes through the per-sample loop, so that the lane stops exactly
    //  after its last sample: a continued stream picks the filter states up from there)
    Coef c0[FPT], dc[FPT];
#pragma unroll
    for (int j = 0; j < FPT; ++j) { c0[j] = cend[j]; dc[j] = cend[j]; }
    struct Coef2 { f2_t a1, g, lp, amp0, amp1, br, m1; };
    Coef2 p0 = { 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull }, pd = p0;   // packed block start / block delta (FPT = 2)
    bool hand = false, warp_exact = false, split = false;
    // (two halves per trip, unrolled: `half` is a constant in each copy and the saw double buffer alternates between two
    //  register sets instead of being copied every 8 samples; wmax and lmax are multiples of 16)
    for (int r0 = -(int)wmax; r0 < (int)lmax; r0 += 16) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = r0 + half * 8;
        const bool act = r >= r_lo && r < r_hi;
        const float4 sa = na, sb = nb;
        if (r + 8 >= r_lo && r + 8 < r_hi) fetch();
        float v[8];
        if (half == 0) {
            hand = act && (!(time > quiet_t) || r + 16 > r_hi);
            const float a_now = time * inv_bl, a_end = fmaf(16.0f, ndt, time) * inv_bl;
            const bool kink = act && !hand && (!(jph < quiet_j) || ((a_now > 1.0f) != (a_end > 1.0f)));
            split = __any_sync(0xffffffffu, kink);
        }
        // A block with a kink in some lane is taken as two 8-sample halves, each classified on its own from the clocks at
        // its start: the half with the kink runs the exact per-sample loop, the other one interpolates over 8 samples
        // (with a voice per utterance a fifth of the blocks have a kink in one of the 32 lanes; config-4 slice: k_formant 4.33 -> 4.03 ms)
        warp_exact = false;
        if (split) {
            const float a_now = time * inv_bl, a_end = fmaf(8.0f, ndt, time) * inv_bl;
            const bool kink = act && !hand && (!(jph < quiet_j8) || ((a_now > 1.0f) != (a_end > 1.0f)));
            warp_exact = __any_sync(0xffffffffu, kink);
        }
        if (FPT == 2 && r == r_join && r_join != -(int)wmax) {
            // the second slot starts here, from rest (whatever the exact / hand-over rows before may have left in it)
            // (... unless it starts at the window's sample 0 of a continued stream: it then holds the carried state)
            if (act && !keep_slot1) { L[FPT - 1].a = 0.f; L[FPT - 1].b = 0.f; L[FPT - 1].c = 0.f; }
            c_valid = false;
        }
        if (act) {
            if (!hand && !warp_exact) {
                const float s8[8] = { sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w };
                // NJ = number of slots that run in this block (a compile-time constant in each instantiation)
                auto interp_block = [&](auto nj_tag) {
                    constexpr int NJ = decltype(nj_tag)::value;
                    if (half == 0 || split) {
                        if (!c_valid) {
                            const float alpha = fminf(time * inv_bl, 1.0f);
#pragma unroll
                            for (int j = 0; j < NJ; ++j) cend[j] = coeffs(L[j], alpha, jph);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {  // the clocks stay literal f32 chains  (:861, :291)
                            time = __fadd_rn(time, ndt);
                            jph = __fadd_rn(jph, jinc);
                        }
                        if (!split) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                time = __fadd_rn(time, ndt);
                                jph = __fadd_rn(jph, jinc);
                            }
                        }
                        const float alpha = fminf(time * inv_bl, 1.0f);
                        const float ds = split ? 2.0f : 1.0f;      // dc is the change per 16 samples
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            c0[j] = cend[j];
                            cend[j] = coeffs(L[j], alpha, jph);
                            dc[j].a1 = (cend[j].a1 - c0[j].a1) * ds; dc[j].g = (cend[j].g - c0[j].g) * ds;
                            dc[j].lp = (cend[j].lp - c0[j].lp) * ds; dc[j].amp0 = (cend[j].amp0 - c0[j].amp0) * ds;
                            dc[j].amp1 = (cend[j].amp1 - c0[j].amp1) * ds; dc[j].br = (cend[j].br - c0[j].br) * ds;
                        }
                        c_valid = true;
                    } else {
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {     // second half: start from the block's midpoint
                            c0[j].a1 = fmaf(dc[j].a1, 0.5f, c0[j].a1); c0[j].g = fmaf(dc[j].g, 0.5f, c0[j].g);
                            c0[j].lp = fmaf(dc[j].lp, 0.5f, c0[j].lp); c0[j].amp0 = fmaf(dc[j].amp0, 0.5f, c0[j].amp0);
                            c0[j].amp1 = fmaf(dc[j].amp1, 0.5f, c0[j].amp1); c0[j].br = fmaf(dc[j].br, 0.5f, c0[j].br);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float d1, nz;
                        noise(s8[k], d1, nz);
                        const float t = (float)k * 0.0625f;
                        float acc = 0.0f;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            Coef c;
                            c.a1 = fmaf(dc[j].a1, t, c0[j].a1); c.g = fmaf(dc[j].g, t, c0[j].g);
                            c.lp = fmaf(dc[j].lp, t, c0[j].lp); c.amp0 = fmaf(dc[j].amp0, t, c0[j].amp0);
                            c.amp1 = fmaf(dc[j].amp1, t, c0[j].amp1); c.br = fmaf(dc[j].br, t, c0[j].br);
                            acc += tick(L[j], c, s8[k], d1, nz);
                        }
                        v[k] = acc;
                    }
                };
                // Both slots running (FPT = 2): the same block with the two formants carried as packed pairs (each half
                // of a packed operation rounds exactly like its scalar counterpart).
                auto interp_block2 = [&]() {
                    if (half == 0 || split) {
                        if (!c_valid) {
                            const float alpha = fminf(time * inv_bl, 1.0f);
#pragma unroll
                            for (int j = 0; j < FPT; ++j) cend[j] = coeffs(L[j], alpha, jph);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {  // the clocks stay literal f32 chains  (:861, :291)
                            time = __fadd_rn(time, ndt);
                            jph = __fadd_rn(jph, jinc);
                        }
                        if (!split) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                time = __fadd_rn(time, ndt);
                                jph = __fadd_rn(jph, jinc);
                            }
                        }
                        const float alpha = fminf(time * inv_bl, 1.0f);
                        const Coef s0 = cend[0], s1 = cend[FPT - 1];
#pragma unroll
                        for (int j = 0; j < FPT; ++j) cend[j] = coeffs(L[j], alpha, jph);
                        const Coef e0 = cend[0], e1 = cend[FPT - 1];
#if KF_ALG2
                        p0.a1 = pk(s0.a1, s1.a1); p0.g = pk(2.0f * s0.g, 2.0f * s1.g); p0.lp = pk(s0.lp, s1.lp);
                        p0.m1 = pk(s0.a1 * s0.g, s1.a1 * s1.g);
                        pd.m1 = sub2(pk(e0.a1 * e0.g, e1.a1 * e1.g), p0.m1);
                        p0.amp0 = pk(s0.amp0, s1.amp0); p0.amp1 = pk(s0.amp1, s1.amp1); p0.br = pk(s0.br, s1.br);
                        pd.a1 = sub2(pk(e0.a1, e1.a1), p0.a1); pd.g = sub2(pk(2.0f * e0.g, 2.0f * e1.g), p0.g);
#else
                        p0.a1 = pk(s0.a1, s1.a1); p0.g = pk(s0.g, s1.g); p0.lp = pk(s0.lp, s1.lp);
                        p0.amp0 = pk(s0.amp0, s1.amp0); p0.amp1 = pk(s0.amp1, s1.amp1); p0.br = pk(s0.br, s1.br);
                        pd.a1 = sub2(pk(e0.a1, e1.a1), p0.a1); pd.g = sub2(pk(e0.g, e1.g), p0.g);
#endif
                        pd.lp = sub2(pk(e0.lp, e1.lp), p0.lp); pd.amp0 = sub2(pk(e0.amp0, e1.amp0), p0.amp0);
                        pd.amp1 = sub2(pk(e0.amp1, e1.amp1), p0.amp1); pd.br = sub2(pk(e0.br, e1.br), p0.br);
                        if (split) {       // an 8-sample block: pd is the change per 16 samples
                            pd.a1 = add2(pd.a1, pd.a1); pd.g = add2(pd.g, pd.g); pd.lp = add2(pd.lp, pd.lp);
                            pd.amp0 = add2(pd.amp0, pd.amp0); pd.amp1 = add2(pd.amp1, pd.amp1); pd.br = add2(pd.br, pd.br);
#if KF_ALG2
                            pd.m1 = add2(pd.m1, pd.m1);
#endif
                        }
                        c_valid = true;
                    } else {
                        const f2_t hf = pk(0.5f, 0.5f);                 // second half: start from the block's midpoint
                        p0.a1 = fma2(pd.a1, hf, p0.a1); p0.g = fma2(pd.g, hf, p0.g); p0.lp = fma2(pd.lp, hf, p0.lp);
                        p0.amp0 = fma2(pd.amp0, hf, p0.amp0); p0.amp1 = fma2(pd.amp1, hf, p0.amp1); p0.br = fma2(pd.br, hf, p0.br);
#if KF_ALG2
                        p0.m1 = fma2(pd.m1, hf, p0.m1);
#endif
                    }
                    f2_t A = pk(L[0].a, L[FPT - 1].a), B = pk(L[0].b, L[FPT - 1].b), Cc = pk(L[0].c, L[FPT - 1].c);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float d1, nz;
                        noise(s8[k], d1, nz);
                        const float tk = (float)k * 0.0625f;
                        const f2_t t = pk(tk, tk), saw2 = pk(s8[k], s8[k]), d2 = pk(d1, d1), n2 = pk(nz, nz);
#if KF_ALG2
                        // (p0.g / pd.g hold 2 g and p0.m1 / pd.m1 hold a1 g here: one product and one sum fewer per sample)
                        const f2_t a1 = fma2(pd.a1, t, p0.a1), g2 = fma2(pd.g, t, p0.g), m1 = fma2(pd.m1, t, p0.m1), lp = fma2(pd.lp, t, p0.lp);
                        const f2_t amp0 = fma2(pd.amp0, t, p0.amp0), amp1 = fma2(pd.amp1, t, p0.amp1), br = fma2(pd.br, t, p0.br);
                        const f2_t nw = fma2(br, d2, saw2);                              // :531
                        A = fma2(lp, sub2(nw, A), A);                                    // :538
                        const f2_t v0 = mul2(A, fma2(amp1, n2, amp0));                   // :544-550
                        const f2_t v3 = sub2(v0, Cc);                                    // :565
                        const f2_t v1 = fma2(m1, v3, mul2(a1, B));                       // a1 b + a2 v3   (:566)
                        Cc = fma2(g2, v1, Cc);                                           // 2 v2 - c = c + 2 g v1
                        B = sub2(mul2(v1, pk(2.0f, 2.0f)), B);                           // 2 v1 - b (one FFMA2 with a negated addend)
#else
                        const f2_t a1 = fma2(pd.a1, t, p0.a1), g = fma2(pd.g, t, p0.g), lp = fma2(pd.lp, t, p0.lp);
                        const f2_t amp0 = fma2(pd.amp0, t, p0.amp0), amp1 = fma2(pd.amp1, t, p0.amp1), br = fma2(pd.br, t, p0.br);
                        const f2_t nw = fma2(br, d2, saw2);                              // :531
                        A = fma2(lp, sub2(nw, A), A);                                    // :538
                        const f2_t v0 = mul2(A, fma2(amp1, n2, amp0));                   // :544-550
                        const f2_t v3 = sub2(v0, Cc);                                    // :565
                        // v1 = a1 (g v3 + b) and c' = 2 (c + g v1) - c, regrouped so that the loop-carried chains are
                        // b -> a1 b -> v1 -> 2 v1 - b (4 operations) and c -> v3 -> v1 -> c' (3) instead of 6: the
                        // same operation count, half the dependent latency per sample
                        const f2_t v1 = fma2(mul2(a1, g), v3, mul2(a1, B));
                        Cc = fma2(add2(g, g), v1, Cc);
                        B = sub2(add2(v1, v1), B);
#endif
                        float x0, x1;
                        unpk(v1, x0, x1);
                        v[k] = x0 + x1;
                    }
                    unpk(A, L[0].a, L[FPT - 1].a);
                    unpk(B, L[0].b, L[FPT - 1].b);
                    unpk(Cc, L[0].c, L[FPT - 1].c);
                };
                if (FPT == 2 && r >= r_join) interp_block2();
                else if (FPT == 1) interp_block(std::integral_constant<int, FPT>{});
                else interp_block(std::integral_constant<int, 1>{});
            } else {
                // a kink somewhere in the warp (exact coefficients every sample) or the phoneme's / the chunk's last
                // samples: ONE generic per-sample loop, rolled (results parked in the smem row) -- its code is a tenth of
                // the straight-line version's, which matters more than its loop overhead once many lanes of a warp have
                // kinks of their own (per-utterance voices): the kernel's hot paths then fit the instruction cache
                c_valid = false;
                const int cnt = hand ? min(8, r_hi - r) : 8;
#pragma unroll 1
                for (int k = 0; k < cnt; ++k) {
                    const float4 q = k < 4 ? sa : sb;
                    const int kk = k & 3;
                    const float s = kk == 0 ? q.x : (kk == 1 ? q.y : (kk == 2 ? q.z : q.w));
                    const float vk = sample(s);
                    time = __fadd_rn(time, ndt);
                    if (time < 0.0f) handover();                             // :864
                    jph = __fadd_rn(jph, jinc);
                    if (jph > 1.0f) jitter_wrap();
                    sts32(part_a + ((r & (KT - 1)) + k) * 4, vk);
                }
                if (r >= 0) {
                    const float4 t0 = lds128(part_a + (r & (KT - 1)) * 4), t1 = lds128(part_a + (r & (KT - 1)) * 4 + 16);
                    v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w;
                    v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = 0.0f;
        }
        if (r >= 0) {
            sts128(part_a + (r & (KT - 1)) * 4, v[0], v[1], v[2], v[3]);
            sts128(part_a + (r & (KT - 1)) * 4 + 16, v[4], v[5], v[6], v[7]);
            if (half == 1 && (r0 & (KT - 1)) == KT - 16) {       // (r & (KT - 1)) == KT - 8
                const uint32_t base = (uint32_t)r - (uint32_t)(KT - 8);
                const uint32_t kb = base / (uint32_t)KT, slot = kb & 1u, use = kb >> 1;   // (RING) batch, its tile, the tile's n-th use
                bool reduce_here = true;
                int rc0 = w * 4, rcs = NW * 4;
                if (RING) {
                    __syncwarp();
                    if (w < NW - 1) {
                        // producer: this batch's tile is complete; the next batch goes to the other tile, once the reducer
                        // has emptied it (its previous use was batch kb - 1)
                        if (lane == 0) mbar_arrive(full_a + slot * 8u);
                        if (kb >= 1u) mbar_wait(empty_a + (slot ^ 1u) * 8u, ((kb - 1u) >> 1) & 1u);
                        reduce_here = false;
                    } else {
                        mbar_wait(full_a + slot * 8u, use & 1u);
                        rc0 = 0; rcs = 4;                      // the reducer takes all 32 rows
                    }
                    part = part_all[slot];
                } else {
                    __syncthreads();
                }
                // 8 lanes per row, 4 rows per pass: sum the formant groups in index order (Array::sum, :123),
                // scale (:574), store 16 bytes per lane = 128 contiguous bytes per row
                const int sub = lane >> 3, l8 = lane & 7;
                // The common case -- mono f32, no ring: the passes unrolled (a compile-time trip count), so that their row
                // addresses differ by immediates instead of being rebuilt from the thread index in every pass.
                constexpr int RED_PASSES = (32 * (KT / 32) + NW * 4 - 1) / (NW * 4);
                const bool plain_out = !RING && P.out_channels == 1u && format == GRAIL_F32;
                if (plain_out) {
#pragma unroll
                    for (int i = 0; i < RED_PASSES; ++i) {
                        const int rc = w * 4 + sub + i * NW * 4;
                        if (rc >= 32 * (KT / 32)) break;
                        const int row = rc & 31, cb = (rc >> 5) * 32;
                        const uint32_t rl = row_len[row];
                        const uint32_t s0 = base + cb + l8 * 4;
                        if (s0 < rl) {
                            float4 acc = *reinterpret_cast<const float4*>(&part[0][row][cb + l8 * 4]);
#pragma unroll
                            for (int f = 1; f < NW; ++f) {
                                const float4 t = *reinterpret_cast<const float4*>(&part[f][row][cb + l8 * 4]);
                                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                            }
                            acc.x *= 0.5f; acc.y *= 0.5f; acc.z *= 0.5f; acc.w *= 0.5f;
                            float* op = reinterpret_cast<float*>(out) + (row_out[row] + s0);
                            const uint32_t cnt = rl - s0;
                            if (cnt >= 4u && (reinterpret_cast<uintptr_t>(op) & 15u) == 0) {
                                *reinterpret_cast<float4*>(op) = acc;
                            } else {
                                const float a4[4] = { acc.x, acc.y, acc.z, acc.w };
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    if ((uint32_t)q < cnt) op[q] = a4[q];
                            }
                        }
                    }
                }
#pragma unroll 1
                for (int rc = rc0 + sub; !plain_out && reduce_here && rc < 32 * (KT / 32); rc += rcs) {
                    const int row = rc & 31, cb = (rc >> 5) * 32;          // (row, 32-sample column block) of the tile
                    const uint32_t rl = row_len[row];
                    const uint32_t s0 = base + cb + l8 * 4;
                    if (s0 < rl) {
                        float4 acc = *reinterpret_cast<const float4*>(&part[0][row][cb + l8 * 4]);
#pragma unroll
                        for (int f = 1; f < NW; ++f) {
                            const float4 t = *reinterpret_cast<const float4*>(&part[f][row][cb + l8 * 4]);
                            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                        }
                        acc.x *= 0.5f; acc.y *= 0.5f; acc.z *= 0.5f; acc.w *= 0.5f;
                        const unsigned long long o = row_out[row] + s0;
                        const uint32_t cnt = min(4u, rl - s0);
                        const float a4[4] = { acc.x, acc.y, acc.z, acc.w };
                        const uint32_t nch = P.out_channels;
                        if (nch != 1u) {
                            // channel duplication, `flat_map(|x| repeat(x).take(num_channels))` (examples/cli.rs:229,
                            // interactive.rs:38): each sample nch times, interleaved
                            for (uint32_t q = 0; q < cnt; ++q) {
                                const float sc = a4[q] * 32767.0f;
                                const short qi = (short)((sc != sc) ? 0 : __float2int_rz(fminf(fmaxf(sc, -32768.0f), 32767.0f)));
                                for (uint32_t c = 0; c < nch; ++c) {
                                    if (format == GRAIL_F32) reinterpret_cast<float*>(out)[(o + q) * nch + c] = a4[q];
                                    else reinterpret_cast<short*>(out)[(o + q) * nch + c] = qi;
                                }
                            }
                        } else if (format == GRAIL_F32) {
                            float* op = reinterpret_cast<float*>(out) + o;
                            if (cnt == 4 && (reinterpret_cast<uintptr_t>(op) & 15u) == 0) {   // 128-bit store when the row is 16-byte aligned
                                *reinterpret_cast<float4*>(op) = acc;
                            } else {
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    if ((uint32_t)q < cnt) op[q] = a4[q];
                            }
                        } else {
                            // (x * i16::MAX as f32) as i16: truncating, saturating, NaN -> 0 (examples/cli.rs:50)
                            short* op = reinterpret_cast<short*>(out) + o;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float sc = a4[q] * 32767.0f;
                                const int qi = (sc != sc) ? 0 : __float2int_rz(fminf(fmaxf(sc, -32768.0f), 32767.0f));
                                if ((uint32_t)q < cnt) op[q] = (short)qi;
                            }
                        }
                    }
                }
                if (RING) {
                    if (w == NW - 1) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty_a + slot * 8u);      // the tile may be written again
                    }
                    part_a = part_a0 + (slot ^ 1u) * SLOT_BYTES;              // every warp: the next batch's tile
                } else {
                    __syncthreads();
                }
            }
        }
      }
    }
    // the lane that owns an utterance's last chunk leaves the Synthesize filter states behind (stream state)
    if (on && it.n0 + it.len == U.n_samples) {
        float* fin = P.utt_final + (size_t)it.utt * 32;
#pragma unroll
        for (int j = 0; j < FPT; ++j)
            if (L[j].fi >= 0) { fin[L[j].fi] = L[j].a; fin[8 + L[j].fi] = L[j].b; fin[16 + L[j].fi] = L[j].c; }
    }
}

// ------------------------------------------------------------------------------------------------
// Output gather helper (multi-GPU): segment i of `src` goes to its place in `dst`.  One CTA per (segment, tile of
// SEG_TILE elements); 16-byte copies where source and destination are equally aligned, element copies otherwise.
// ------------------------------------------------------------------------------------------------
struct SegCopy { unsigned long long dst, src, len; };   // byte offsets / byte length
constexpr unsigned long long SEG_TILE_BYTES = 1ull << 16;
__global__ void __launch_bounds__(256) k_copy_segments(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src,
                                                       const SegCopy* __restrict__ segs, const uint32_t* __restrict__ tile_seg,
                                                       const uint32_t* __restrict__ tile_idx, uint32_t elem_bytes)
{
    const SegCopy sg = segs[tile_seg[blockIdx.x]];
    const unsigned long long b0 = (unsigned long long)tile_idx[blockIdx.x] * SEG_TILE_BYTES;
    const unsigned long long b1 = min(b0 + SEG_TILE_BYTES, sg.len);
    unsigned char* d = dst + sg.dst;
    const unsigned char* s = src + sg.src;
    if ((((uintptr_t)(d + b0)) & 15u) == (((uintptr_t)(s + b0)) & 15u)) {
        // head up to the next 16-byte boundary, body in 16-byte words, tail: element by element
        unsigned long long head = (16u - (((uintptr_t)(d + b0)) & 15u)) & 15u;
        head = min(head, b1 - b0);
        for (unsigned long long i = b0 + (unsigned long long)threadIdx.x * elem_bytes; i < b0 + head; i += 256ull * elem_bytes)
            for (uint32_t k = 0; k < elem_bytes; ++k) d[i + k] = s[i + k];
        const unsigned long long body0 = b0 + head, nvec = (b1 - body0) >> 4;
        const uint4* sv = reinterpret_cast<const uint4*>(s + body0);
        uint4* dv = reinterpret_cast<uint4*>(d + body0);
        for (unsigned long long i = threadIdx.x; i < nvec; i += 256) dv[i] = __ldg(sv + i);
        for (unsigned long long i = body0 + (nvec << 4) + (unsigned long long)threadIdx.x * elem_bytes; i < b1; i += 256ull * elem_bytes)
            for (uint32_t k = 0; k < elem_bytes; ++k) d[i + k] = s[i + k];
    } else if (elem_bytes == 4) {
        const uint32_t* sv = reinterpret_cast<const uint32_t*>(s + b0);
        uint32_t* dv = reinterpret_cast<uint32_t*>(d + b0);
        for (unsigned long long i = threadIdx.x; i < ((b1 - b0) >> 2); i += 256) dv[i] = __ldg(sv + i);
    } else {
        const uint16_t* sv = reinterpret_cast<const uint16_t*>(s + b0);
        uint16_t* dv = reinterpret_cast<uint16_t*>(d + b0);
        for (unsigned long long i = threadIdx.x; i < ((b1 - b0) >> 1); i += 256) dv[i] = sv[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Roofline probes: dense FFMA issue rate and MUFU.RCP rate of this device.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_probe_ffma(float* sink, int iters, float seed)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
          a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 0.001f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) k_probe_mufu(float* sink, int iters, float seed)
{
    // x <- rcp(x) + 1 (converges to the golden ratio; never folds): one MUFU.RCP + one FADD per step
    float a0 = seed + 1.5f + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = frcp(a0) + 1.0f; a1 = frcp(a1) + 1.0f; a2 = frcp(a2) + 1.0f; a3 = frcp(a3) + 1.0f;
        }
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

} // namespace grail

#include "grail_phase.cuh"
