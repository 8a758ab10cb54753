// grail_runtime.cu -- host runtime and C ABI (include/grail_cuda.h) of the B200 waveform path.
//
// Host side of the cut: exact per-phoneme schedule (the Sequencer's f32 clock in closed form), work
// planning (time chunks x active formants), device memory, launches, transfers.  "Phoneme scheduling
// stays on the host" (north_star); everything per-sample runs in grail_kernels.cuh.
// Product code: no CPU synthesis path exists here -- without a device every compute entry point fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "grail_kernels.cuh"

using namespace grail;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct PoolBuf {
    void*  ptr;
    size_t size;
    bool   in_use;
};

struct grail_ctx {
    int            device = 0;
    cudaStream_t   stream = nullptr;    // main (API-visible) stream
    cudaStream_t   s_front = nullptr;   // pipelined plans: schedule / frequency / phase kernels of launch k+1 ...
    cudaStream_t   s_back = nullptr;    // ... overlap the formant kernel of launch k
    int            pipeline = 0;        // ctx option "pipeline" (off by default so that launches complete in stream order without a join; +5 % at config 2)
    cudaDeviceProp prop{};
    std::string    err;
    std::vector<PoolBuf> pool;
    // options
    double   warmup_nepers = 11.5;   // exp(-11.5) = 1e-5 of the state at a chunk start: measured 1.9e-6 / 110.5 dB worst case at 2 048-sample
                                     // chunks against 1.1e-6 / 111.9 dB at 13.8 (the f32 noise floor), for 3.3 % less k_formant time
    uint32_t target_items = 0;      // 0 = one resident wave of k_formant CTAs
    uint32_t min_chunk = 2048;
    uint32_t max_chunk = 1u << 22;
    int      debug_taps = 0;
    uint32_t pscan_min = 1u << 18;   // utterances at least this long may get the exact parallel phase scan
    int phase_lean = -1;             // k_phase_pair build: -1 by batch size, 0 latency build, 1 few-register build
    int interleave = 1;              // interleave equally long utterances chunk by chunk in k_formant's CTAs
    int pscan_cost_model = 1;        // 0: scan every utterance >= pscan_min (at most 16), whatever it costs
    int      zero_copy_out = 0;      // 1: k_formant stores straight into pinned host output (measured slower: 23 vs 50 GB/s over PCIe)
    int      formants_per_lane = 2;  // 1 or 2 formants of an utterance share one lane's clocks, noise and saw
    int      phase_mode = 1;         // 1: chunk-parallel exact carrier phase (grail_phase.cuh); 0: serial chains (+ phase scan for long utterances)
    uint32_t phase_chunk = 0;        // samples per phase chunk (multiple of 256; 0: chosen by the planner)
    uint32_t walk_warps_per_sm = 0;  // resident warps per SM of the phase walks (occupancy query, first plan)
    int      phase_rounds = -1;      // repair rounds enqueued after the first proof (-1: by the longest utterance)
    cudaStream_t   s_copy = nullptr;    // one-shot batches: device-to-host copies of finished utterance groups
    int      e2e_groups = -1;        // one-shot batches: utterance groups whose copies overlap the next group's kernels (-1 auto)
    std::vector<grail_plan*> live_plans;   // plans created on this ctx and not yet destroyed (synchronize clears their in_flight)
    // Plan uploads: the host-built tables go through pinned staging blocks on their own stream, so building the next
    // plan neither waits for the kernels already queued on `stream` (a copy from pageable memory synchronises the
    // stream it is issued on) nor delays them.
    cudaStream_t   s_up = nullptr;
    std::vector<PoolBuf> hpool;            // pinned host blocks, reused across plans
};

static int set_err(grail_ctx* ctx, int status, const char* fmt, ...)
{
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return status;
}

#define CU(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return set_err((ctx), e_ == cudaErrorMemoryAllocation ? GRAIL_ERR_OOM : GRAIL_ERR_CUDA, \
                           "%s failed: %s", #call, cudaGetErrorString(e_));                        \
    } while (0)

static int pool_alloc(grail_ctx* ctx, size_t bytes, void** out)
{
    if (bytes == 0) bytes = 256;
    bytes = (bytes + 255) & ~(size_t)255;
    int best = -1;
    for (size_t i = 0; i < ctx->pool.size(); ++i) {
        PoolBuf& b = ctx->pool[i];
        if (!b.in_use && b.size >= bytes && b.size <= bytes * 2 + (1u << 20))
            if (best < 0 || b.size < ctx->pool[best].size) best = (int)i;
    }
    if (best >= 0) {
        ctx->pool[best].in_use = true;
        *out = ctx->pool[best].ptr;
        return GRAIL_OK;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        // drop cached free buffers and retry once
        for (auto& b : ctx->pool)
            if (!b.in_use && b.ptr) { cudaFree(b.ptr); b.ptr = nullptr; b.size = 0; }
        ctx->pool.erase(std::remove_if(ctx->pool.begin(), ctx->pool.end(), [](const PoolBuf& b) { return !b.ptr; }),
                        ctx->pool.end());
        cudaGetLastError();
        e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_err(ctx, GRAIL_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        }
    }
    ctx->pool.push_back({ p, bytes, true });
    *out = p;
    return GRAIL_OK;
}

static void pool_free(grail_ctx* ctx, void* p)
{
    if (!p) return;
    for (auto& b : ctx->pool)
        if (b.ptr == p) { b.in_use = false; return; }
}

// pinned host staging blocks (same discipline as the device pool: smallest free block that fits, else a new one)
static int hpool_alloc(grail_ctx* ctx, size_t bytes, void** out)
{
    bytes = (bytes + 0xFFFFFull) & ~0xFFFFFull;          // whole MiB
    PoolBuf* best = nullptr;
    for (auto& b : ctx->hpool)
        if (!b.in_use && b.size >= bytes && (!best || b.size < best->size)) best = &b;
    if (best) { best->in_use = true; *out = best->ptr; return GRAIL_OK; }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return set_err(ctx, GRAIL_ERR_OOM, "cudaHostAlloc(%zu) for plan staging failed", bytes);
    }
    ctx->hpool.push_back(PoolBuf{ p, bytes, true });
    *out = p;
    return GRAIL_OK;
}
static void hpool_free(grail_ctx* ctx, void* p)
{
    if (!p) return;
    for (auto& b : ctx->hpool)
        if (b.ptr == p) { b.in_use = false; return; }
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct grail_plan {
    void*       h_stage = nullptr;      // pinned staging block of this plan's uploads (back to the ctx's pool with the plan)
    cudaEvent_t ev_up = nullptr;        // uploads done (the main stream waits on it)
    grail_ctx* ctx = nullptr;
    uint32_t n_utts = 0, n_elems = 0, n_items = 0, n_groups = 0, n_jscheds = 0, n_jrecs = 0, nw = 1, fpt = 1;
    uint32_t chunk_len = 0;
    uint64_t total_samples = 0, f_words = 0, saw_words = 0;
    std::vector<UttDev>      utts;       // host copies (original utterance order)
    std::vector<ItemDev>     items;
    std::vector<JitSchedDev> jscheds;
    std::vector<SegRec>      segs;
    std::vector<uint64_t>    out_offsets;
    // device
    float* d_elems = nullptr; SegRec* d_segs = nullptr; UttDev* d_utts = nullptr; ItemDev* d_items = nullptr;
    JitSchedDev* d_jscheds = nullptr; JitRec* d_jrecs = nullptr; float* d_F = nullptr; float* d_saw = nullptr;
    float* d_phase_dbg = nullptr; uint32_t* d_err = nullptr; uint32_t* d_fflags = nullptr;
    float* d_utt_init = nullptr; float* d_utt_final = nullptr;
    void* d_out = nullptr; size_t d_out_bytes = 0; int d_out_format = -1;
    cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    // Pipelined plans (grail_cuda_plan_create with ctx option "pipeline"): two scratch sets, so that the front
    // kernels of launch k+1 (own stream) run under the formant kernel of launch k.  Slot 0 is d_F/d_fflags/d_saw.
    struct Slot {
        float* F = nullptr; uint32_t* fflags = nullptr; float* saw = nullptr;
        cudaEvent_t front_done = nullptr, back_done = nullptr;
        cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
        bool used = false;
    } slot[2];
    uint32_t out_channels = 1;
    bool pipelined = false, in_flight = false;
    uint32_t launch_idx = 0, last_slot = 0;
    cudaEvent_t ev_begin = nullptr;
    bool launched = false;
    std::vector<PScanDev> pscans;          // exact parallel phase scans (one per long utterance)
    std::vector<uint32_t> pscan_utt;
    uint32_t* d_pscan_status = nullptr;
    float* d_pchunks = nullptr; DirtyRec* d_pdrec = nullptr; float* d_ppark = nullptr; uint32_t pc_stride = 0; double* d_bsum = nullptr; uint32_t* d_utt_status = nullptr; uint32_t* d_pstats = nullptr;
    uint32_t n_pchunks = 0, phase_chunk = 0, max_pchunks = 0, pc_per_item = 1;   // phase_chunk == 0: serial chains only
    bool select_on_device = false;   // d_elems was written by k_select from phoneme-level input
    std::vector<void*> pscan_bufs;
    bool jit_on_host = false;   // few distinct jitter increments: schedules computed by the planner
    std::vector<JitRec> jrecs;
    uint32_t last_launches = 0;
};

// exact Sequencer schedule of one utterance (reference src/lib.rs:859-888): per phoneme the index of its
// first sample and the clock value there; returns the sample count, or <0 on unsupported input.
struct SeqKey {
    uint32_t t, len, dt;
    bool operator==(const SeqKey& o) const { return t == o.t && len == o.len && dt == o.dt; }
};
struct SeqKeyHash {
    size_t operator()(const SeqKey& k) const
    {
        uint64_t h = (uint64_t)k.t * 0x9E3779B97F4A7C15ull ^ ((uint64_t)k.len << 32 | k.dt);
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        return (size_t)h;
    }
};
struct SeqVal { uint64_t steps; uint32_t x; };
typedef std::unordered_map<SeqKey, SeqVal, SeqKeyHash> SeqCache;

static const uint64_t MAX_UTT_SAMPLES = (1ull << 31) - 4096;

// How a single-utterance plan continues a stream (grail_stream): the three iterators' carried state.
struct StreamStart {
    bool     cont_phoneme = false;  // elems[0] is a phoneme already in progress ...
    float    time0 = 0.0f;          //   ... whose Sequencer.time at the window's first sample is this
    float    t_neg = 0.0f;          // else: the (negative) Sequencer.time carried into elems[0]'s hand-over
    bool     fresh = true;          // the very first window: t_neg = 0 - delta_time
    float    jitter_phase = 0.0f;   // value-noise phase after the last produced sample
    uint32_t jitter_wraps = 0;      // value-noise wraps so far
    uint64_t sample0 = 0;           // absolute index of the window's first sample
    float    carrier_phase = 0.0f;  // Synthesize.phase
    const float* filter_state = nullptr;   // 24 floats a[8] b[8] c[8], or null (all zero)
    uint64_t max_samples = ~0ull;   // truncate the window here
    bool     finished = true;       // false: the last element is only a look-ahead, its own samples are not due yet
};


// What a plan is built from: Sequencer input records (the path's own boundary), or phoneme-level input that the
// device expands with k_select (Selector, src/lib.rs:987-1005; Intonator stub, :1057-1075).  The host only ever
// needs each element's length, whether it has a sound, and which formant amplitudes are non-zero.
struct ElemInput {
    const grail_seq_elem* full = nullptr;
    const grail_phoneme_elem* ph = nullptr;     // PhonemeElem records, or
    const uint8_t* ids = nullptr;               // bare phoneme ids (Intonator on the device too)
    const float* center = nullptr;              // per utterance, with ids
    const grail_elem* storages = nullptr;       // n_storages x n_sounds
    uint32_t n_sounds = 0, n_storages = 0;
    const uint32_t* utt_storage = nullptr;      // per utterance, null = storage 0

    bool phoneme_level() const { return full == nullptr && (ph != nullptr || ids != nullptr); }
    float length(uint32_t p) const { return full ? full[p].length : (ph ? ph[p].length : 0.5f); }   // :1069
    uint32_t phoneme(uint32_t p) const { return ph ? ph[p].phoneme : (uint32_t)ids[p]; }
    bool has_elem(uint32_t p) const
    {
        if (full) return full[p].has_elem != 0;
        return phoneme(p) >= GRAIL_PHONEME_FIRST_SOUND;                                             // :658
    }
    float amp(uint32_t u, uint32_t p, int i) const
    {
        if (full) return full[p].elem.formant_amp[i];
        const uint32_t st = utt_storage ? utt_storage[u] : 0u;
        return storages[(size_t)st * n_sounds + (phoneme(p) - GRAIL_PHONEME_FIRST_SOUND)].formant_amp[i];
    }
};

template <class LenFn>
static int64_t schedule_utterance_fn(LenFn length_of, uint32_t n_elems, float sample_rate, SegRec* segs,
                                     SeqCache& cache, const StreamStart* ss = nullptr, float* t_neg_out = nullptr)
{
    const float dt = sdiv(1.0f, sample_rate);       // src/lib.rs:944
    float t_neg = ssub(0.0f, dt);                   // first call: time = 0 - delta_time  (:861)
    if (ss && !ss->fresh) t_neg = ss->t_neg;
    uint64_t n = 0;
    for (uint32_t p = 0; p < n_elems; ++p) {
        const float len = length_of(p);
        float time0 = sadd(t_neg, len);             // :873 / :882
        if (p == 0 && ss && ss->cont_phoneme) time0 = ss->time0;   // a phoneme already in progress
        if (segs) { segs[p].start = (uint32_t)n; segs[p].time0 = time0; }
        const SeqKey key = { f2u(time0), 0u, f2u(dt) };
        auto itc = cache.find(key);
        SeqVal v;
        if (itc != cache.end()) {
            v = itc->second;
        } else {
            const ClockRun r = clock_desc_run(time0, dt, MAX_UTT_SAMPLES + 1);
            if (r.stuck || !(r.x < 0.0f)) return -1; // clock cannot reach zero: the reference would never end
            v.steps = r.steps;
            v.x = f2u(r.x);
            if (cache.size() < (1u << 20)) cache.emplace(key, v);
        }
        n += v.steps;
        if (n > MAX_UTT_SAMPLES) return -1;
        t_neg = u2f(v.x);
    }
    if (t_neg_out) *t_neg_out = t_neg;
    return (int64_t)n;
}

static int64_t schedule_utterance(const grail_seq_elem* e, uint32_t n_elems, float sample_rate, SegRec* segs,
                                  SeqCache& cache, const StreamStart* ss = nullptr, float* t_neg_out = nullptr)
{
    return schedule_utterance_fn([e](uint32_t p) { return e[p].length; }, n_elems, sample_rate, segs, cache, ss, t_neg_out);
}

static int validate_inputs(grail_ctx* ctx, const ElemInput& in, const uint32_t* utt_offsets,
                           const grail_voice_params* voices, uint32_t n_utts)
{
    if (!utt_offsets || !voices || (n_utts && !in.full && !in.phoneme_level() && utt_offsets[n_utts] != 0))
        return set_err(ctx, GRAIL_ERR_INVALID_ARG, "null input pointer");
    if (in.phoneme_level()) {
        if (!in.storages || in.n_sounds == 0 || in.n_storages == 0 || (in.ids && !in.center))
            return set_err(ctx, GRAIL_ERR_INVALID_ARG, "phoneme-level input needs a voice storage (and centre frequencies with bare ids)");
        for (uint32_t u = 0; u < n_utts; ++u) {
            if (in.utt_storage && in.utt_storage[u] >= in.n_storages)
                return set_err(ctx, GRAIL_ERR_INVALID_ARG, "utterance %u: voice storage index %u out of range", u, in.utt_storage[u]);
            if (utt_offsets[u + 1] < utt_offsets[u]) continue;          // reported below
            for (uint32_t p = utt_offsets[u]; p < utt_offsets[u + 1]; ++p)
                if (in.phoneme(p) >= GRAIL_PHONEME_FIRST_SOUND + in.n_sounds)
                    return set_err(ctx, GRAIL_ERR_INVALID_ARG, "utterance %u: phoneme id %u has no entry in the voice storage", u, in.phoneme(p));
        }
    }
    for (uint32_t u = 0; u < n_utts; ++u) {
        if (utt_offsets[u + 1] < utt_offsets[u])
            return set_err(ctx, GRAIL_ERR_INVALID_ARG, "utt_offsets not monotone at %u", u);
        const grail_voice_params& v = voices[u];
        if (!(v.sample_rate > 0.0f) || !std::isfinite(v.sample_rate))
            return set_err(ctx, GRAIL_ERR_INVALID_ARG, "utterance %u: sample_rate must be finite and > 0", u);
        const float dt = sdiv(1.0f, v.sample_rate);
        if (!std::isnormal(dt))
            return set_err(ctx, GRAIL_ERR_INVALID_ARG, "utterance %u: 1/sample_rate is not a normal f32", u);
        if (!(v.jitter_frequency >= 0.0f && v.jitter_frequency <= 0.25f))
            return set_err(ctx, GRAIL_ERR_UNSUPPORTED,
                           "utterance %u: jitter_frequency %g outside the supported range [0, 0.25]", u,
                           (double)v.jitter_frequency);
        if (!std::isfinite(v.jitter_delta_frequency) || !std::isfinite(v.jitter_delta_formant_frequency) ||
            !std::isfinite(v.jitter_delta_amplitude))
            return set_err(ctx, GRAIL_ERR_INVALID_ARG, "utterance %u: non-finite jitter scalar", u);
        for (uint32_t p = utt_offsets[u]; p < utt_offsets[u + 1]; ++p)
            if (!std::isfinite(in.length(p)))
                return set_err(ctx, GRAIL_ERR_INVALID_ARG, "utterance %u: non-finite phoneme length", u);
    }
    return GRAIL_OK;
}

template <int NW, int FPT>
static void launch_formant(const PlanDev& P, void* out, int format, cudaStream_t s)
{
    k_formant<NW, FPT><<<P.n_groups, NW * 32, 0, s>>>(P, out, format);
}
template <int NW, int FPT>
static int formant_occupancy()
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_formant<NW, FPT>, NW * 32, 0);
    return nb;
}
// NW = warps per CTA (formant groups), FPT = formants per lane
static int formant_occupancy_of(int nw, int fpt)
{
    if (fpt == 2) {
        switch (nw) {
        case 1: return formant_occupancy<1, 2>();
        case 2: return formant_occupancy<2, 2>();
        case 3: return formant_occupancy<3, 2>();
        default: return formant_occupancy<4, 2>();
        }
    }
    switch (nw) {
    case 1: return formant_occupancy<1, 1>();
    case 2: return formant_occupancy<2, 1>();
    case 3: return formant_occupancy<3, 1>();
    case 4: return formant_occupancy<4, 1>();
    case 5: return formant_occupancy<5, 1>();
    case 6: return formant_occupancy<6, 1>();
    case 7: return formant_occupancy<7, 1>();
    default: return formant_occupancy<8, 1>();
    }
}
static void launch_formant_of(int nw, int fpt, const PlanDev& P, void* out, int format, cudaStream_t s)
{
    if (fpt == 2) {
        switch (nw) {
        case 1: launch_formant<1, 2>(P, out, format, s); break;
        case 2: launch_formant<2, 2>(P, out, format, s); break;
        case 3: launch_formant<3, 2>(P, out, format, s); break;
        default: launch_formant<4, 2>(P, out, format, s); break;
        }
        return;
    }
    switch (nw) {
    case 1: launch_formant<1, 1>(P, out, format, s); break;
    case 2: launch_formant<2, 1>(P, out, format, s); break;
    case 3: launch_formant<3, 1>(P, out, format, s); break;
    case 4: launch_formant<4, 1>(P, out, format, s); break;
    case 5: launch_formant<5, 1>(P, out, format, s); break;
    case 6: launch_formant<6, 1>(P, out, format, s); break;
    case 7: launch_formant<7, 1>(P, out, format, s); break;
    default: launch_formant<8, 1>(P, out, format, s); break;
    }
}

static PlanDev plan_dev(const grail_plan* pl, bool with_dbg, int slot = 0)
{
    PlanDev P;
    P.elems = pl->d_elems; P.segs = pl->d_segs; P.utts = pl->d_utts; P.items = pl->d_items;
    P.jscheds = pl->d_jscheds; P.jrecs = pl->d_jrecs; P.F = pl->slot[slot].F; P.saw = pl->slot[slot].saw;
    P.phase_dbg = with_dbg ? pl->d_phase_dbg : nullptr;
    P.err = pl->d_err;
    P.fflags = pl->slot[slot].fflags;
    P.pscan_status = pl->d_pscan_status;
    P.utt_init = pl->d_utt_init;
    P.utt_final = pl->d_utt_final;
    P.pchunks = pl->phase_chunk ? pl->d_pchunks : nullptr;
    P.pdrec = pl->d_pdrec;
    P.ppark = pl->d_ppark;
    P.pdirty = pl->phase_chunk ? reinterpret_cast<uint32_t*>(pl->d_pchunks + (size_t)pl->pc_stride * PCF_COUNT) : nullptr;
    P.bsum = pl->phase_chunk ? pl->d_bsum : nullptr;
    P.utt_status = pl->d_utt_status;
    P.pstats = pl->d_pstats;
    P.phase_chunk = pl->phase_chunk;
    P.n_pchunks = pl->n_pchunks;
    P.pc_stride = pl->pc_stride;
    P.pc_per_item = pl->pc_per_item;
    P.f_tiled = pl->phase_chunk ? 1u : 0u;
    P.n_utts = pl->n_utts; P.n_items = pl->n_items; P.n_groups = pl->n_groups; P.n_jscheds = pl->n_jscheds;
    P.chunk_len = pl->chunk_len;
    P.out_channels = pl->out_channels;
    P.warmup_nepers = (float)pl->ctx->warmup_nepers;
    return P;
}

static void plan_release(grail_plan* pl)
{
    if (!pl) return;
    grail_ctx* ctx = pl->ctx;
    ctx->live_plans.erase(std::remove(ctx->live_plans.begin(), ctx->live_plans.end(), pl), ctx->live_plans.end());
    void* bufs[] = { pl->d_elems, pl->d_segs, pl->d_utts, pl->d_items, pl->d_jscheds, pl->d_jrecs, pl->d_F, pl->d_saw,
                     pl->d_phase_dbg, pl->d_err, pl->d_out, pl->d_fflags, pl->d_utt_init, pl->d_utt_final,
                     pl->d_pchunks, pl->d_bsum, pl->d_utt_status, pl->d_pstats, pl->d_pdrec, pl->d_ppark };
    for (void* b : bufs) pool_free(ctx, b);
    for (void* b : pl->pscan_bufs) pool_free(ctx, b);
    pool_free(ctx, pl->slot[1].F); pool_free(ctx, pl->slot[1].fflags); pool_free(ctx, pl->slot[1].saw);
    for (auto& sl : pl->slot) {
        if (sl.front_done) cudaEventDestroy(sl.front_done);
        if (sl.back_done) cudaEventDestroy(sl.back_done);
        for (auto& e : sl.ev) if (e) cudaEventDestroy(e);
    }
    if (pl->ev_begin) cudaEventDestroy(pl->ev_begin);
    if (pl->ev_up) { cudaEventSynchronize(pl->ev_up); cudaEventDestroy(pl->ev_up); }   // (the staging block is free to reuse after this)
    hpool_free(ctx, pl->h_stage);
    pool_free(ctx, pl->d_pscan_status);
    for (auto& e : pl->ev)
        if (e) cudaEventDestroy(e);
    delete pl;
}

static int plan_build(grail_ctx* ctx, const ElemInput& in, const uint32_t* utt_offsets,
                      const grail_voice_params* voices, uint32_t n_utts, grail_plan** out_plan,
                      const StreamStart* ss = nullptr, bool pipelined = false)
{
    // (ss: one StreamStart per utterance -- the windows of n_utts concurrent streams share one plan and one launch)
    int rc = validate_inputs(ctx, in, utt_offsets, voices, n_utts);
    if (rc) return rc;
    CU(ctx, cudaSetDevice(ctx->device));
    grail_plan* pl = new (std::nothrow) grail_plan();
    if (!pl) return set_err(ctx, GRAIL_ERR_OOM, "host allocation failed");
    pl->ctx = ctx;
    pl->pipelined = pipelined && ctx->pipeline;
    pl->n_utts = n_utts;
    pl->n_elems = n_utts ? utt_offsets[n_utts] : 0;
    pl->utts.resize(n_utts);
    pl->segs.resize(pl->n_elems ? pl->n_elems : 1);
    pl->out_offsets.assign(n_utts + 1, 0);

    static const bool ptrace = getenv("GRAIL_PLAN_TRACE") != nullptr;
    const auto pt0 = std::chrono::steady_clock::now();
    auto pmark = [&](const char* what) {
        if (ptrace) fprintf(stderr, "[grail plan] %-28s at %.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - pt0).count());
    };
    pmark("validated");
    // ---- exact schedule + active formants
    SeqCache cache;
    std::unordered_map<uint64_t, uint32_t> jmap;
    std::vector<float> init_states;                  // continued streams: 32 floats per utterance
    if (ss) init_states.assign((size_t)std::max<uint32_t>(n_utts, 1) * 32, 0.0f);
    const StreamStart* const ss_all = ss;
    uint32_t nw = 1;
    uint64_t total = 0, f_words = 0;
    uint32_t n_max = 0;
    for (uint32_t u = 0; u < n_utts; ++u) {
        UttDev& U = pl->utts[u];
        memset(&U, 0, sizeof U);
        U.elem_first = utt_offsets[u];
        U.n_elems = utt_offsets[u + 1] - utt_offsets[u];
        const uint32_t e0 = U.elem_first;
        const StreamStart* ss = ss_all ? ss_all + u : nullptr;      // this utterance's carried state
        if (ss && ss->filter_state) memcpy(init_states.data() + (size_t)u * 32, ss->filter_state, 24 * sizeof(float));
        int64_t n = schedule_utterance_fn([&in, e0](uint32_t p) { return in.length(e0 + p); }, U.n_elems, voices[u].sample_rate,
                                          pl->segs.data() + U.elem_first, cache, ss);
        if (n >= 0 && ss) {
            // an unfinished stream holds its last element back as look-ahead: only the phonemes before it are due
            if (!ss->finished && U.n_elems > 0) n = (int64_t)pl->segs[U.elem_first + U.n_elems - 1].start;
            if ((uint64_t)n > ss->max_samples) n = (int64_t)ss->max_samples;
        }
        if (n < 0) {
            plan_release(pl);
            return set_err(ctx, GRAIL_ERR_UNSUPPORTED,
                           "utterance %u: more than %llu samples or a phoneme clock that never reaches zero", u,
                           (unsigned long long)MAX_UTT_SAMPLES);
        }
        U.n_samples = (uint32_t)n;
        U.voice = voices[u];
        U.init_phase = ss ? ss->carrier_phase : 0.0f;
        U.pscan = -1;
        U.jw0 = ss ? ss->jitter_wraps : 0u;
        U.sample0 = ss ? ss->sample0 : 0ull;
        U.has_init = (ss && ss->filter_state) ? 1u : 0u;
        U.out_off = total;
        U.f_off = f_words;
        total += (uint64_t)n;
        f_words += ((uint64_t)n + 255) & ~255ull;   // 256-aligned: k_frequency's 128-sample runs index the flag words
        n_max = std::max(n_max, U.n_samples);
        pl->out_offsets[u + 1] = total;
        // a formant whose amplitude is zero in every element contributes exactly 0 (v0 = 0 keeps the SVF at rest)
        for (int i = 0; i < NF; ++i) {
            bool act = false;
            for (uint32_t p = 0; p < U.n_elems && !act; ++p)
                act = in.has_elem(e0 + p) && !(in.amp(u, e0 + p, i) == 0.0f);
            // a continued stream: a formant that is still ringing stays active whatever the new elements say
            if (ss && ss->filter_state)
                act = act || ss->filter_state[i] != 0.0f || ss->filter_state[8 + i] != 0.0f || ss->filter_state[16 + i] != 0.0f;
            if (act) U.active[U.n_active++] = (uint8_t)i;
        }
        nw = std::max(nw, U.n_active);
        // jitter schedules are shared by every utterance with the same increment (and, for streams, the same phase)
        const uint64_t key = (uint64_t)f2u(voices[u].jitter_frequency) | ((uint64_t)f2u(ss ? ss->jitter_phase : 0.0f) << 32);
        auto jt = jmap.find(key);
        if (jt == jmap.end()) {
            JitSchedDev js;
            memset(&js, 0, sizeof js);
            js.inc = voices[u].jitter_frequency;
            js.phase0 = ss ? ss->jitter_phase : 0.0f;
            js.n_max = U.n_samples;
            jmap.emplace(key, (uint32_t)pl->jscheds.size());
            U.jit_sched = (uint32_t)pl->jscheds.size();
            pl->jscheds.push_back(js);
        } else {
            U.jit_sched = jt->second;
            pl->jscheds[jt->second].n_max = std::max(pl->jscheds[jt->second].n_max, U.n_samples);
        }
    }
    pl->total_samples = total;
    pl->f_words = f_words + 256;
    pl->fpt = (uint32_t)ctx->formants_per_lane;
    pl->nw = (nw + pl->fpt - 1) / pl->fpt;   // warps per CTA: ceil(active formants / formants per lane)
    nw = pl->nw;
    uint64_t n_jrecs = 0;
    for (auto& js : pl->jscheds) {
        js.rec_first = (uint32_t)n_jrecs;
        // (the f32 clock's effective step can exceed `inc` by up to a third when inc is near the phase's ulp: round-to-nearest
        //  inflates it -- so a generous bound, not the real-number rate)
        js.rec_cap = (uint32_t)((double)js.n_max * (double)js.inc * 1.5) + 16;
        n_jrecs += js.rec_cap;
    }
    if (n_jrecs > 0xFFFFFFF0ull) {
        plan_release(pl);
        return set_err(ctx, GRAIL_ERR_UNSUPPORTED, "jitter schedule too large");
    }
    pl->n_jscheds = (uint32_t)pl->jscheds.size();
    pl->n_jrecs = (uint32_t)n_jrecs;
    // a handful of distinct increments (the common case: one voice) is cheaper on the host than a
    // latency-bound single-lane kernel; thousands (per-utterance voices) go to k_jitter_schedule
    if (pl->n_jscheds <= 64 || ss_all) {   // (stream windows: always, the host needs the schedule for the carried state)
        pl->jit_on_host = true;
        pl->jrecs.resize(std::max<uint32_t>(pl->n_jrecs, 1));
        for (auto& js : pl->jscheds) {
            js.n_recs = jitter_schedule_walk(js.inc, js.n_max, pl->jrecs.data() + js.rec_first, js.rec_cap, js.phase0);
            if (js.n_recs == 0) {
                plan_release(pl);
                return set_err(ctx, GRAIL_ERR_UNSUPPORTED, "jitter schedule overflow");
            }
        }
    }

    pmark("schedule + active formants");
    // ---- chunking: aim for one resident wave of k_formant CTAs (32 chunks each)
    uint64_t target = ctx->target_items;
    if (target == 0) {
        int occ = formant_occupancy_of((int)nw, (int)pl->fpt);
        if (occ < 1) occ = 1;
        {
            // (pipelined plans use the same target: the next launch's phase kernel needs two CTAs of 25 K registers per
            // SM at most and finds them once the first formant CTAs retire; reserving room for it up front was
            // measured slower -- 3.22 ms per step against 2.95 at the full target, 3.03 without pipelining)
            // Fewer, longer chunks beat full occupancy (the warm-up is paid per chunk): aim at 12 warps per SM with two
            // formants per lane (what 167 registers allow; measured at config 2: 6 CTAs/SM 1.69 ms, 5 -> 1.91, 4 -> 1.78)
            // and 3/4 of the resident warps with one, and keep the warps per SM a multiple of 4 so the four
            // sub-partitions stay evenly loaded.
            int c = pl->fpt == 2 ? std::min(occ, std::max(1, KF_WARPS2 / (int)nw)) : std::max(1, (3 * occ) / 4);
            while (c > 1 && ((c * (int)nw) % 4) != 0) --c;
            occ = c;
        }
        target = (uint64_t)ctx->prop.multiProcessorCount * (uint64_t)occ * 32ull;
    }
    // smallest chunk length (multiple of 32) whose work items still fit in `target` lanes: the grid is then a
    // single full wave with no tail; when even one chunk per utterance does not fit, utterances stay whole
    auto items_at = [&](uint64_t c) {
        uint64_t n = 0;
        for (uint32_t u = 0; u < n_utts; ++u) n += (pl->utts[u].n_samples + c - 1) / c;
        return n;
    };
    const uint64_t G = 256;             // chunk granularity (k_phase tiles, k_frequency runs, k_formant batches)
    const uint64_t cl_hi = ((uint64_t)std::max<uint32_t>(n_max, 1u) + G - 1) / G * G;
    uint64_t lo = 1, hi = cl_hi / G;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) / 2;
        if (items_at(mid * G) <= target) hi = mid; else lo = mid + 1;
    }
    uint64_t cl = lo * G;
    cl = std::max<uint64_t>(cl, (ctx->min_chunk + G - 1) / G * G);
    cl = std::min<uint64_t>(cl, std::max<uint64_t>((ctx->max_chunk + G - 1) / G * G, G));
    cl = std::min<uint64_t>(cl, cl_hi);
    pl->chunk_len = (uint32_t)cl;

    pmark("chunk length");
    // ---- work items: utterances in descending length so a CTA's 32 chunks are of similar size
    std::vector<uint32_t> order(n_utts);
    for (uint32_t u = 0; u < n_utts; ++u) order[u] = u;
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t b) { return pl->utts[a].n_samples > pl->utts[b].n_samples; });
    // A CTA of k_formant is 32 consecutive items.  When 32 neighbours in this order have the same number of chunks
    // they are interleaved chunk by chunk (item = first + c * 32): a CTA then holds the SAME chunk of 32 utterances,
    // so with a shared voice the value-noise wraps, alpha clips and phoneme hand-overs fall on the same samples in
    // every lane and the warp leaves its interpolating fast path 32 times less often.  Otherwise: consecutive.
    auto chunks_of = [&](uint32_t u) { return (pl->utts[u].n_samples + pl->chunk_len - 1) / pl->chunk_len; };
    for (size_t o0 = 0; o0 < order.size();) {
        bool same = ctx->interleave && o0 + 32 <= order.size() && chunks_of(order[o0]) > 0;
        for (size_t k = 1; same && k < 32; ++k) same = chunks_of(order[o0 + k]) == chunks_of(order[o0]);
        if (same) {
            while (pl->items.size() & 31u) {             // start on a CTA boundary: empty items (idle lanes)
                ItemDev pad;
                pad.utt = order[o0]; pad.n0 = 0; pad.len = 0; pad.pad = 0;
                pl->items.push_back(pad);
            }
            const uint32_t base = (uint32_t)pl->items.size(), nc = chunks_of(order[o0]);
            for (uint32_t c = 0; c < nc; ++c)
                for (uint32_t k = 0; k < 32; ++k) {
                    UttDev& U = pl->utts[order[o0 + k]];
                    const uint32_t n0 = c * pl->chunk_len;
                    ItemDev it;
                    it.utt = order[o0 + k]; it.n0 = n0; it.len = std::min(pl->chunk_len, U.n_samples - n0); it.pad = 0;
                    pl->items.push_back(it);
                }
            for (uint32_t k = 0; k < 32; ++k) {
                UttDev& U = pl->utts[order[o0 + k]];
                U.item_first = base + k; U.item_stride = 32; U.n_items = nc;
            }
            o0 += 32;
            continue;
        }
        UttDev& U = pl->utts[order[o0]];
        U.item_first = (uint32_t)pl->items.size();
        U.item_stride = 1;
        for (uint32_t n0 = 0; n0 < U.n_samples; n0 += pl->chunk_len) {
            ItemDev it;
            it.utt = order[o0]; it.n0 = n0; it.len = std::min(pl->chunk_len, U.n_samples - n0); it.pad = 0;
            pl->items.push_back(it);
        }
        U.n_items = (uint32_t)pl->items.size() - U.item_first;
        ++o0;
    }
    pl->n_items = (uint32_t)pl->items.size();
    pl->n_groups = (pl->n_items + 31) / 32;
    pl->saw_words = (uint64_t)pl->n_groups * 32ull * pl->chunk_len;

    pmark("work items");
    // ---- chunk-parallel exact phase (grail_phase.cuh): every work item is cut into K phase chunks of PC samples, a
    //      multiple of 256; an utterance's chunks are numbered in time order.  PC "auto": enough chunks for about 16
    //      warps of walks per SM (the walks are streaming kernels), at least 1024 samples (a chunk without a carrier
    //      wrap has no anchor: shorter chunks fail their proofs more often), at most 4096.
    pl->phase_chunk = 0;
    // (A handful of very long utterances -- config 3 -- took the fixed-point phase scan of grail_kernels.cuh while the
    //  per-utterance scans of this path (k_phase_guess / _scan_a / _fix) ran as ONE warp per utterance, the wrong shape
    //  for 25 000 chunks of a single utterance: 5.8 ms against 3.3 ms.  With a CTA of 32 warps per utterance for such
    //  plans the chunk-parallel path takes 0.92 ms for the same 26 457 161 samples, bit-identical output; the
    //  fixed-point scan stays selectable: phase_mode 3.)
    const bool few_long = ctx->phase_mode == 3 && n_utts <= 16 && n_max >= ctx->pscan_min;
    if (ctx->phase_mode && !few_long && pl->n_items) {
        uint64_t pc = ctx->phase_chunk;
        if (pc == 0) {
            // a warp of the walks is one sub-range of one group of 32 items: K sub-ranges per item make n_groups * K warps.
            // Largest K whose warps are all resident at once (one wave: a second, partial wave doubles the kernel's
            // time) with chunks of at least 1024 samples; batches too large for that get 4096-sample chunks.
            if (ctx->walk_warps_per_sm == 0) {
                int nb = 0;
                cudaFuncSetAttribute(k_phase_a, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_phase_b, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_phase_b, 128, 0) != cudaSuccess || nb < 1) nb = 1;
                cudaGetLastError();
                ctx->walk_warps_per_sm = (uint32_t)nb * 4u;
            }
            const uint64_t slots = (uint64_t)ctx->prop.multiProcessorCount * ctx->walk_warps_per_sm;
            const uint64_t cl = pl->chunk_len;
            auto pc_of = [&](uint64_t k) { return ((cl + k - 1) / k + 255) & ~255ull; };
            // Cost model (in units of one chunk walk of one sample): round 0 streams every chunk, wave after wave of
            // `slots` resident warps, so it lasts about ceil(waves) * PC (a partial second wave costs a whole one);
            // every repair round is latency-bound and lasts one chunk walk, about 4x slower per sample than the
            // bandwidth-bound round 0 (measured), two to three rounds per launch.  Chunks shorter than 1024 samples lose
            // their anchor (a chunk needs a carrier wrap) and fail their proofs far more often; 4096 is the upper end.
            uint64_t best_k = 1;
            double best = 1e300;
            for (uint64_t k = 1; k <= std::max<uint64_t>(1, cl / 1024) + 1; ++k) {
                const uint64_t p_ = pc_of(k);
                if (k > 1 && p_ < 1024) break;
                if (p_ > 4096 && pc_of(k + 1) >= 1024) continue;
                const uint64_t warps = (uint64_t)pl->n_groups * ((cl + p_ - 1) / p_);
                const double waves = std::ceil((double)warps / (double)slots);
                const double cost = waves * (double)p_ + 2.5 * 4.0 * (double)p_ / 8.0;
                if (cost < best) { best = cost; best_k = k; }
            }
            pc = pc_of(best_k);
        }
        // long forms: the repairs dominate (five or six rounds, each re-walking most of what lies downstream of a failed
        // boundary), and they cost per chunk -- measured 1.20 ms at 1 024 against 0.92 ms at 2 048 for config 3
        if (ctx->phase_chunk == 0 && (uint64_t)n_max / std::max<uint64_t>(pc, 1) >= 4096) pc = std::max<uint64_t>(pc, 2048);
        pc = std::max<uint64_t>(256, (pc + 255) & ~255ull);
        pc = std::min<uint64_t>(pc, pl->chunk_len);
        pl->phase_chunk = (uint32_t)pc;
        pl->pc_per_item = (pl->chunk_len + pl->phase_chunk - 1) / pl->phase_chunk;
        uint64_t npc = 0;
        for (uint32_t u = 0; u < n_utts; ++u) {
            UttDev& U = pl->utts[u];
            U.pc_first = (uint32_t)npc;
            U.pc_count = 0;
            if (U.n_samples) {
                const uint32_t full = (U.n_samples - 1) / pl->chunk_len;              // items before the last one
                const uint32_t tail = U.n_samples - full * pl->chunk_len;              // samples of the last item
                U.pc_count = full * pl->pc_per_item + (tail + pl->phase_chunk - 1) / pl->phase_chunk;
            }
            npc += U.pc_count;
            pl->max_pchunks = std::max(pl->max_pchunks, U.pc_count);
        }
        if (npc > 0xFFFFFFF0ull || (uint64_t)pl->n_groups * pl->pc_per_item > 0x7FFFFFFull) {
            plan_release(pl);
            return set_err(ctx, GRAIL_ERR_UNSUPPORTED, "too many phase chunks");
        }
        pl->n_pchunks = (uint32_t)npc;
        // F_t shares the saw's tiled layout; the tiled offsets are noted in units of 8 floats in 32 bits
        if (pl->saw_words >= (1ull << 35)) {
            plan_release(pl);
            return set_err(ctx, GRAIL_ERR_UNSUPPORTED, "batch too large for one plan (2^35 scratch words)");
        }
    }

    // ---- long utterances: exact parallel phase scan instead of the serial chain.  The chains of k_phase_pair run
    //      concurrently, so that kernel lasts as long as its longest utterance (4.7 ns/sample measured); a scan costs
    //      about 0.40 ms of launches plus 0.125 ns/sample and scans run one after another.  Take the k longest
    //      utterances (k <= 16) that minimise scans + longest remaining chain.
    {
        std::vector<uint32_t> longs;
        for (uint32_t u = 0; u < n_utts && !pl->phase_chunk; ++u)   // (the chunk-parallel path needs no special case for them)
            if (pl->utts[u].n_samples >= ctx->pscan_min) longs.push_back(u);
        std::sort(longs.begin(), longs.end(), [&](uint32_t a, uint32_t b) { return pl->utts[a].n_samples > pl->utts[b].n_samples; });
        uint32_t rest_max = 0;                           // longest utterance below the threshold
        for (uint32_t u = 0; u < n_utts; ++u)
            if (pl->utts[u].n_samples < ctx->pscan_min) rest_max = std::max(rest_max, pl->utts[u].n_samples);
        auto chain_ms = [](uint32_t n) { return 4.7e-6 * (double)n; };
        auto scan_ms = [](uint32_t n) { return 0.40 + 0.125e-6 * (double)n; };
        size_t best_k = 0;
        double best = chain_ms(longs.empty() ? rest_max : std::max(rest_max, pl->utts[longs[0]].n_samples)), scans = 0.0;
        for (size_t k = 1; k <= longs.size() && k <= 16; ++k) {
            scans += scan_ms(pl->utts[longs[k - 1]].n_samples);
            const uint32_t next = k < longs.size() ? pl->utts[longs[k]].n_samples : 0u;
            const double t = scans + chain_ms(std::max(rest_max, next));
            if (t < best || !ctx->pscan_cost_model) { best = t; best_k = k; }
        }
        for (size_t i = 0; i < best_k; ++i) {
            pl->utts[longs[i]].pscan = (int32_t)pl->pscan_utt.size();
            pl->pscan_utt.push_back(longs[i]);
        }
    }

    pmark("phase chunks");
    // ---- device buffers
    auto fail = [&](int code) { plan_release(pl); return code; };
#define PA(ptr, bytes)                                                         \
    do {                                                                       \
        void* p_ = nullptr;                                                    \
        int rc_ = pool_alloc(ctx, (bytes), &p_);                               \
        if (rc_) return fail(rc_);                                             \
        (ptr) = reinterpret_cast<decltype(ptr)>(p_);                           \
    } while (0)
#define CUF(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            set_err(ctx, GRAIL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));      \
            return fail(GRAIL_ERR_CUDA);                                                       \
        }                                                                                      \
    } while (0)
    PA(pl->d_elems, (size_t)std::max<uint32_t>(pl->n_elems, 1) * sizeof(grail_seq_elem) + 256);
    PA(pl->d_segs, pl->segs.size() * sizeof(SegRec));
    PA(pl->d_utts, std::max<size_t>(n_utts, 1) * sizeof(UttDev));
    PA(pl->d_items, std::max<size_t>(pl->n_items, 1) * sizeof(ItemDev));
    PA(pl->d_jscheds, std::max<size_t>(pl->n_jscheds, 1) * sizeof(JitSchedDev));
    PA(pl->d_jrecs, std::max<size_t>(pl->n_jrecs, 1) * sizeof(JitRec));
    // (F_t of a chunk-parallel plan is tiled like the saw)
    const uint64_t F_words = pl->phase_chunk ? std::max<uint64_t>(pl->saw_words, 8) + 256 : pl->f_words;
    PA(pl->d_F, F_words * sizeof(float));
    PA(pl->d_fflags, (pl->f_words / 128 + 2) * sizeof(uint32_t));
    PA(pl->d_saw, std::max<uint64_t>(pl->saw_words, 8) * sizeof(float));
    PA(pl->d_err, 256);
    PA(pl->d_utt_final, std::max<size_t>(n_utts, 1) * 32 * sizeof(float));
    PA(pl->d_utt_init, std::max<size_t>(n_utts, 1) * 32 * sizeof(float));
    PA(pl->d_pscan_status, 256 + 64 * pl->pscan_utt.size());
    PA(pl->d_utt_status, std::max<size_t>(n_utts, 1) * sizeof(uint32_t));
    PA(pl->d_pstats, 256);
    if (pl->phase_chunk) {
        pl->pc_stride = (pl->n_pchunks + 64u) & ~31u;
        PA(pl->d_pchunks, (size_t)pl->pc_stride * (PCF_COUNT + 2) * sizeof(float));   // + the two dirty lists
        PA(pl->d_pdrec, (size_t)pl->pc_stride * sizeof(DirtyRec));
        PA(pl->d_ppark, (size_t)pl->pc_stride * ((pl->phase_chunk + 7) / 8) * sizeof(float));
        PA(pl->d_bsum, (pl->f_words / 256 + 2) * sizeof(double));
    }
    for (uint32_t u : pl->pscan_utt) {
        const UttDev& U = pl->utts[u];
        PScanDev S;
        memset(&S, 0, sizeof S);
        const uint64_t n = U.n_samples, nb = (n + 1 + SCAN_TILE - 1) / SCAN_TILE;
        void *p = nullptr, *q = nullptr, *b = nullptr;
        int rc1 = pool_alloc(ctx, (n + 1) * 8, &p);
        if (!rc1) { pl->pscan_bufs.push_back(p); rc1 = pool_alloc(ctx, n * 8 + 8, &q); }
        void *fl = nullptr, *bp = nullptr, *cl = nullptr, *bi = nullptr, *stp = nullptr, *lb = nullptr;
        if (!rc1) { pl->pscan_bufs.push_back(q); rc1 = pool_alloc(ctx, nb * 8 + 8, &b); }
        if (!rc1) { pl->pscan_bufs.push_back(b); rc1 = pool_alloc(ctx, n + 16, &fl); }
        if (!rc1) { pl->pscan_bufs.push_back(fl); rc1 = pool_alloc(ctx, n / PS_BLOCK + 16, &bp); }
        if (!rc1) { pl->pscan_bufs.push_back(bp); rc1 = pool_alloc(ctx, n + 16, &cl); }
        if (!rc1) { pl->pscan_bufs.push_back(cl); rc1 = pool_alloc(ctx, n / PS_BLOCK + 16, &bi); }
        if (!rc1) { pl->pscan_bufs.push_back(bi); rc1 = pool_alloc(ctx, (n / PS_BLOCK + 16) * 4, &stp); }
        if (!rc1) { pl->pscan_bufs.push_back(stp); rc1 = pool_alloc(ctx, (n / PS_BLOCK + 16) * 4, &lb); }
        if (rc1) return fail(rc1);
        pl->pscan_bufs.push_back(lb);
        S.P = (unsigned long long*)p; S.inc = (unsigned long long*)q; S.bsum = (unsigned long long*)b;
        S.sflag = (unsigned char*)fl; S.bpar = (unsigned char*)bp; S.cls = (unsigned char*)cl;
        S.bpin = (unsigned char*)bi; S.stamp = (uint32_t*)stp; S.lbk = (uint32_t*)lb;
        S.n = U.n_samples;
        S.p0 = (unsigned long long)(U.init_phase * 1099511627776.0f);
        pl->pscans.push_back(S);
    }
    cudaStream_t s = ctx->stream;
    // Uploads: one pinned staging block per plan, filled here, copied on the upload stream; the main stream waits on
    // an event.  Nothing in this function waits for work already queued on the main stream.
    const bool staged = !in.phoneme_level();
    size_t stage_bytes = 0;
    auto stage_room = [&](size_t bytes) { const size_t o = stage_bytes; stage_bytes += (bytes + 255) & ~(size_t)255; return o; };
    const size_t so_elems = stage_room((size_t)pl->n_elems * sizeof(grail_seq_elem));
    const size_t so_segs = stage_room(pl->segs.size() * sizeof(SegRec));
    const size_t so_utts = stage_room((size_t)n_utts * sizeof(UttDev));
    const size_t so_items = stage_room((size_t)pl->n_items * sizeof(ItemDev));
    const size_t so_js = stage_room((size_t)pl->n_jscheds * sizeof(JitSchedDev));
    const size_t so_jr = stage_room((size_t)pl->n_jrecs * sizeof(JitRec));
    const size_t so_init = stage_room(init_states.size() * sizeof(float));
    cudaStream_t su = staged ? ctx->s_up : s;
    char* hs = nullptr;
    if (staged) {
        void* hp = nullptr;
        int rch = hpool_alloc(ctx, stage_bytes + 256, &hp);
        if (rch) return fail(rch);
        pl->h_stage = hp;
        hs = (char*)hp;
    }
    auto upload = [&](void* dst, const void* src, size_t bytes, size_t off) -> cudaError_t {
        if (!bytes) return cudaSuccess;
        if (!staged) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
        memcpy(hs + off, src, bytes);
        return cudaMemcpyAsync(dst, hs + off, bytes, cudaMemcpyHostToDevice, su);
    };
    if (pl->n_elems && in.full) CUF(upload(pl->d_elems, in.full, (size_t)pl->n_elems * sizeof(grail_seq_elem), so_elems));
    CUF(upload(pl->d_segs, pl->segs.data(), pl->segs.size() * sizeof(SegRec), so_segs));
    if (n_utts) CUF(upload(pl->d_utts, pl->utts.data(), n_utts * sizeof(UttDev), so_utts));
    if (pl->n_elems && in.phoneme_level()) {
        // Selector (and, for bare ids, the Intonator stub) on the device: 1-16 bytes per phoneme cross PCIe instead of 208
        SelectDev S;
        memset(&S, 0, sizeof S);
        auto up = [&](const void* host, size_t bytes, const void** dev) -> int {
            void* d = nullptr;
            int rc_ = pool_alloc(ctx, bytes + 16, &d);
            if (rc_) return rc_;
            pl->pscan_bufs.push_back(d);                       // freed with the plan
            if (cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) return GRAIL_ERR_CUDA;
            *dev = d;
            return 0;
        };
        int rcs = 0;
        if (in.ph) rcs = up(in.ph, (size_t)pl->n_elems * sizeof(grail_phoneme_elem), (const void**)&S.ph);
        else {
            rcs = up(in.ids, (size_t)pl->n_elems, (const void**)&S.ids);
            if (!rcs) rcs = up(in.center, (size_t)n_utts * 4, (const void**)&S.center);
        }
        if (!rcs) rcs = up(in.storages, (size_t)in.n_storages * in.n_sounds * sizeof(grail_elem), (const void**)&S.storages);
        if (!rcs && in.utt_storage) rcs = up(in.utt_storage, (size_t)n_utts * 4, (const void**)&S.utt_storage);
        if (rcs) return fail(rcs == GRAIL_ERR_CUDA ? set_err(ctx, GRAIL_ERR_CUDA, "upload of phoneme-level input failed") : rcs);
        S.n_sounds = in.n_sounds;
        S.n_elems = pl->n_elems;
        S.n_utts = n_utts;
        const uint64_t words = (uint64_t)pl->n_elems * SELECT_WORDS;
        k_select<<<(unsigned)((words + 255) / 256), 256, 0, s>>>((uint32_t*)pl->d_elems, pl->d_utts, S);
        CUF(cudaGetLastError());
        pl->select_on_device = true;
    }
    if (pl->n_items) CUF(upload(pl->d_items, pl->items.data(), pl->n_items * sizeof(ItemDev), so_items));
    if (pl->n_jscheds) CUF(upload(pl->d_jscheds, pl->jscheds.data(), pl->n_jscheds * sizeof(JitSchedDev), so_js));
    if (pl->jit_on_host && pl->n_jrecs) CUF(upload(pl->d_jrecs, pl->jrecs.data(), (size_t)pl->n_jrecs * sizeof(JitRec), so_jr));
    {
        // (these buffers come from the pool: whatever used them before was queued on `stream` or has been drained, and
        //  the upload stream is ordered before `stream` below -- so the memsets go first on the stream that writes them)
        CUF(cudaMemsetAsync(pl->d_utt_init, 0, std::max<size_t>(n_utts, 1) * 32 * sizeof(float), su));
        CUF(cudaMemsetAsync(pl->d_utt_final, 0, std::max<size_t>(n_utts, 1) * 32 * sizeof(float), su));
        CUF(cudaMemsetAsync(pl->d_utt_status, 0, std::max<size_t>(n_utts, 1) * sizeof(uint32_t), su));
        CUF(cudaMemsetAsync(pl->d_pstats, 0, 256, su));
        if (ss) {   // formants nobody touches in this window keep their carried state
            const size_t ib = init_states.size() * sizeof(float);
            if (staged) {
                memcpy(hs + so_init, init_states.data(), ib);
                CUF(cudaMemcpyAsync(pl->d_utt_init, hs + so_init, ib, cudaMemcpyHostToDevice, su));
                CUF(cudaMemcpyAsync(pl->d_utt_final, hs + so_init, ib, cudaMemcpyHostToDevice, su));
            } else {
                CUF(cudaMemcpyAsync(pl->d_utt_init, init_states.data(), ib, cudaMemcpyHostToDevice, s));
                CUF(cudaMemcpyAsync(pl->d_utt_final, init_states.data(), ib, cudaMemcpyHostToDevice, s));
            }
        }
    }
    for (auto& e : pl->ev) CUF(cudaEventCreate(&e));
    pl->slot[0].F = pl->d_F; pl->slot[0].fflags = pl->d_fflags; pl->slot[0].saw = pl->d_saw;
    if (pl->pipelined) {
        PA(pl->slot[1].F, F_words * sizeof(float));
        PA(pl->slot[1].fflags, (pl->f_words / 128 + 2) * sizeof(uint32_t));
        PA(pl->slot[1].saw, std::max<uint64_t>(pl->saw_words, 8) * sizeof(float));
        CUF(cudaEventCreateWithFlags(&pl->ev_begin, cudaEventDisableTiming));
        for (auto& sl : pl->slot) {
            CUF(cudaEventCreateWithFlags(&sl.front_done, cudaEventDisableTiming));
            CUF(cudaEventCreateWithFlags(&sl.back_done, cudaEventDisableTiming));
            for (auto& e : sl.ev) CUF(cudaEventCreate(&e));
        }
    }
    if (staged) {
        // the copies read the plan's own pinned block: nothing to wait for here; the kernels wait on the device
        CUF(cudaEventCreateWithFlags(&pl->ev_up, cudaEventDisableTiming));
        CUF(cudaEventRecord(pl->ev_up, su));
        CUF(cudaStreamWaitEvent(s, pl->ev_up, 0));
    } else {
        // (phoneme-level input) the host vectors are read by the async copies above: make them safe to outlive this call
        CUF(cudaStreamSynchronize(s));
    }
#undef PA
#undef CUF
    ctx->live_plans.push_back(pl);
    *out_plan = pl;
    return GRAIL_OK;
}

// make the ctx's main stream wait for everything this plan has in flight on the pipeline streams
static int plan_join(grail_plan* pl)
{
    grail_ctx* ctx = pl->ctx;
    if (pl->pipelined && pl->in_flight) {
        const grail_plan::Slot& sl = pl->slot[pl->last_slot];
        CU(ctx, cudaStreamWaitEvent(ctx->stream, sl.front_done, 0));
        CU(ctx, cudaStreamWaitEvent(ctx->stream, sl.back_done, 0));
        pl->in_flight = false;
    }
    return GRAIL_OK;
}

static int plan_enqueue(grail_plan* pl, void* d_out, int format, bool with_dbg, bool formant)
{
    grail_ctx* ctx = pl->ctx;
    if (format != GRAIL_F32 && format != GRAIL_I16) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "unknown sample format");
    CU(ctx, cudaSetDevice(ctx->device));
    // Stream layout.  Plain plans run the whole path on the main stream.  Pipelined plans alternate between two
    // scratch sets: the front kernels (schedule, frequency, phase) go to s_front, the formant kernel to s_back, so
    // the latency-bound phase chains of launch k+1 execute in the shadow of launch k's formant kernel.
    const bool pipe = pl->pipelined && !with_dbg && formant;
    int slot = 0;
    cudaStream_t sf = ctx->stream, sb = ctx->stream;
    cudaEvent_t* ev = pl->ev;
    if (pipe) {
        slot = (int)(pl->launch_idx++ & 1u);
        grail_plan::Slot& sl = pl->slot[slot];
        sf = ctx->s_front;
        sb = ctx->s_back;
        ev = sl.ev;
        if (!pl->in_flight) {   // first launch of a burst: order it after what the caller queued on the main stream
            CU(ctx, cudaEventRecord(pl->ev_begin, ctx->stream));
            CU(ctx, cudaStreamWaitEvent(sf, pl->ev_begin, 0));
        }
        if (sl.used) CU(ctx, cudaStreamWaitEvent(sf, sl.back_done, 0));   // the scratch set is free again
    } else if (pl->pipelined) {
        int rc = plan_join(pl);    // a debug / partial launch on a pipelined plan: drain first, then run in order
        if (rc) return rc;
    }
    const PlanDev P = plan_dev(pl, with_dbg, slot);
    cudaStream_t s = sf;
    pl->last_launches = 0;
    CU(ctx, cudaMemsetAsync(pl->d_err, 0, 4, s));
    CU(ctx, cudaMemsetAsync(pl->d_pstats, 0, 256, s));
    CU(ctx, cudaEventRecord(ev[0], s));
    if (pl->n_items && !pl->jit_on_host) {
        k_jitter_schedule<<<(pl->n_jscheds + 63) / 64, 64, 0, s>>>(P);
        pl->last_launches++;
    }
    CU(ctx, cudaEventRecord(ev[1], s));
    if (pl->n_items) {
        const uint32_t run_len = pl->n_jscheds == 1 ? (uint32_t)FREQ_RUN_MAX : 256u;   // (see k_frequency)
        const uint32_t runs = (pl->chunk_len + run_len - 1) / run_len;
        const uint64_t threads = (uint64_t)pl->n_groups * 32ull * runs;
        k_frequency<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(P, runs, run_len);
        pl->last_launches++;
    }
    CU(ctx, cudaEventRecord(ev[2], s));
    for (size_t i = 0; i < pl->pscans.size(); ++i) {
        PScanDev S = pl->pscans[i];
        const uint32_t u = pl->pscan_utt[i];
        S.F = P.F + pl->utts[u].f_off;
        S.status = pl->d_pscan_status + 16 * i;
        CU(ctx, cudaMemsetAsync(S.status, 0, 64, s));
        const uint32_t n = S.n, nb = (uint32_t)(((uint64_t)n + 1 + SCAN_TILE - 1) / SCAN_TILE);
        const uint32_t g256 = (n + 255) / 256;
        auto scan = [&](int verify, uint32_t next_round) {
            k_ps_scan_reduce<<<nb, SCAN_THREADS, 0, s>>>(S);
            k_ps_scan_spine<<<1, 1024, 0, s>>>(S, nb);
            k_ps_scan_apply<<<nb, SCAN_THREADS, 0, s>>>(S, verify, next_round);
        };
        k_ps_init<<<g256, 256, 0, s>>>(S);
        scan(0, 1);                                        // round 0: unrounded prefix sum; every block stamped for round 1
        pl->last_launches += 4;
        // measured: 1-2 rounds at 44 k samples, 3 at 0.4-1.3 M, 5 at 26 M; enqueue about twice that
        int rounds = 4;
        for (uint64_t m = 1ull << 18; m < n && rounds < PS_MAX_ROUNDS; m <<= 1) ++rounds;
        for (int r = 0; r < rounds; ++r) {                 // every kernel returns at once after convergence
            const uint32_t nblk = (n + PS_BLOCK - 1) / PS_BLOCK;
            k_ps_replay<<<(nblk + 127) / 128, 128, 0, s>>>(S, (uint32_t)r + 1u);
            k_ps_parity_spine<<<1, 1024, 0, s>>>(S, nblk);
            k_ps_parity_fix<<<(nblk + 127) / 128, 128, 0, s>>>(S);
            scan(1, (uint32_t)r + 2u);
            k_ps_check<<<1, 1, 0, s>>>(S);
            pl->last_launches += 7;
        }
        k_ps_saw<<<((n + 7) / 8 + 255) / 256, 256, 0, s>>>(S, P, u);
        pl->last_launches++;
    }
    if (pl->n_items && pl->phase_chunk) {
        // chunk-parallel exact phase (grail_phase.cuh): guess, round A, scan, round B, then proof / repair rounds
        const unsigned wg = (unsigned)(((uint64_t)pl->n_utts * 32 + 127) / 128);
        const unsigned cg = (unsigned)(((uint64_t)pl->n_groups * pl->pc_per_item + 3) / 4);   // a warp = one sub-range of one group of 32 items
        // repair rounds walk a dense list of dirty chunks whose length only the device knows: the grid covers the
        // worst case (every chunk dirty) and the CTAs past the list's end leave at once
        // repair rounds walk a dense list of dirty chunks whose length only the device knows: fixed grids stride over it
        const unsigned dgc = std::min<unsigned>((pl->n_pchunks + 31) / 32, (unsigned)ctx->prop.multiProcessorCount * 6u);
        const dim3 dgs(std::max(1u, std::min(128u, (pl->n_pchunks + 4095u) / 4096u)), (pl->phase_chunk + 7u) / 8u);
        // the per-utterance scans: a warp per utterance, or -- long forms -- a CTA of 32 warps per utterance
        const bool cta_scans = pl->max_pchunks >= 2048;
        auto guess = [&]() { if (cta_scans) k_phase_guess<32><<<pl->n_utts, 1024, 0, s>>>(P); else k_phase_guess<1><<<wg, 128, 0, s>>>(P); };
        auto scan_a = [&]() { if (cta_scans) k_phase_scan_a<32><<<pl->n_utts, 1024, 0, s>>>(P); else k_phase_scan_a<1><<<wg, 128, 0, s>>>(P); };
        auto fix = [&](uint32_t r) { if (cta_scans) k_phase_fix<32><<<pl->n_utts, 1024, 0, s>>>(P, r); else k_phase_fix<1><<<wg, 128, 0, s>>>(P, r); };
        guess();
        pl->last_launches++;
        if (pl->max_pchunks > 1) {
            k_phase_a<<<cg, 128, 0, s>>>(P);
            scan_a();
            pl->last_launches += 2;
        }
        k_phase_b<<<cg, 128, 0, s>>>(P, 0u);
        pl->last_launches++;
        // repair rounds: a handful for seconds of speech, a few more for long forms (the guesses drift with the square
        // root of the length, and every round's corrections are an order of magnitude smaller than the last)
        int rounds = 4;
        for (uint32_t m = 256; m < pl->max_pchunks && rounds < PH_MAX_ROUNDS; m <<= 1) ++rounds;
        if (ctx->phase_rounds >= 0) rounds = std::min(ctx->phase_rounds, (int)PH_MAX_ROUNDS);
        if (pl->max_pchunks <= 1) rounds = 0;
        for (int r = 1; r <= rounds; ++r) {
            fix((uint32_t)r);
            k_phase_chain<<<dgc, 32, 0, s>>>(P, (uint32_t)r);
            k_phase_saw<<<dgs, 512, 0, s>>>(P, (uint32_t)r);   // (y: the blocks of a chunk)
            pl->last_launches += 3;
        }
        fix((uint32_t)PH_MAX_ROUNDS + 1u);   // the final proof: sets the status bits
        pl->last_launches++;
    }
    if (pl->n_items) {
        const uint32_t ctas = (pl->n_utts + PH_UTTS - 1) / PH_UTTS;
        const bool lean = ctx->phase_lean < 0 ? ctas > 2u * (uint32_t)ctx->prop.multiProcessorCount : ctx->phase_lean != 0;
        if (lean) k_phase_pair<true><<<ctas, PH_UTTS * 64, 0, s>>>(P);
        else k_phase_pair<false><<<ctas, PH_UTTS * 64, 0, s>>>(P);
        pl->last_launches++;
    }
    CU(ctx, cudaEventRecord(ev[3], s));
    if (pipe) {
        CU(ctx, cudaEventRecord(pl->slot[slot].front_done, sf));
        CU(ctx, cudaStreamWaitEvent(sb, pl->slot[slot].front_done, 0));
        // ev[3] was recorded on the front stream; the formant kernel's own start is the later of that and the end of
        // the previous formant kernel, so its launch time is measured between two events on the back stream
        CU(ctx, cudaEventRecord(ev[3], sb));
    }
    s = sb;
    if (pl->n_items && formant) {
        launch_formant_of((int)pl->nw, (int)pl->fpt, P, d_out, format, s);
        pl->last_launches++;
    }
    CU(ctx, cudaEventRecord(ev[4], s));
    if (pipe) {
        CU(ctx, cudaEventRecord(pl->slot[slot].back_done, sb));
        pl->slot[slot].used = true;
        pl->in_flight = true;
        pl->last_slot = (uint32_t)slot;
    }
    CU(ctx, cudaGetLastError());
    pl->launched = true;
    return GRAIL_OK;
}

static int plan_check_device_errors(grail_plan* pl)
{
    grail_ctx* ctx = pl->ctx;
    uint32_t e = 0;
    int rcj = plan_join(pl);
    if (rcj) return rcj;
    CU(ctx, cudaMemcpyAsync(&e, pl->d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (e & DEV_ERR_JIT_OVERFLOW) return set_err(ctx, GRAIL_ERR_CUDA, "device: jitter schedule overflow");
    if (e) return set_err(ctx, GRAIL_ERR_CUDA, "device error word 0x%x", e);
    return GRAIL_OK;
}

static size_t format_bytes(int format) { return format == GRAIL_I16 ? 2 : 4; }

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int grail_cuda_abi_version(void) { return GRAIL_ABI_VERSION; }

int grail_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* grail_cuda_status_string(int status)
{
    switch (status) {
    case GRAIL_OK: return "ok";
    case GRAIL_ERR_INVALID_ARG: return "invalid argument";
    case GRAIL_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
    case GRAIL_ERR_CUDA: return "CUDA error";
    case GRAIL_ERR_OOM: return "out of memory";
    case GRAIL_ERR_COUNT_MISMATCH: return "output offsets disagree with the exact sample counts";
    case GRAIL_ERR_UNSUPPORTED: return "input outside the supported domain";
    default: return "unknown status";
    }
}

int grail_cuda_create(int device, grail_ctx** out_ctx)
{
    if (!out_ctx) return GRAIL_ERR_INVALID_ARG;
    *out_ctx = nullptr;
    int n = grail_cuda_device_count();
    if (n <= 0 || device < 0 || device >= n) return GRAIL_ERR_NO_DEVICE;
    grail_ctx* ctx = new (std::nothrow) grail_ctx();
    if (!ctx) return GRAIL_ERR_OOM;
    ctx->device = device;
    int prio_lo = 0, prio_hi = 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess ||
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->s_front, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->s_back, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_up, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        grail_cuda_destroy(ctx);             // destroys whichever streams were created
        return GRAIL_ERR_CUDA;
    }
    if (ctx->prop.major < 10) {
        // built for sm_100a only; an older device cannot run the cubin
        grail_cuda_destroy(ctx);
        return GRAIL_ERR_NO_DEVICE;
    }
    *out_ctx = ctx;
    return GRAIL_OK;
}

void grail_cuda_destroy(grail_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->s_front) { cudaStreamSynchronize(ctx->s_front); cudaStreamDestroy(ctx->s_front); }
    if (ctx->s_back) { cudaStreamSynchronize(ctx->s_back); cudaStreamDestroy(ctx->s_back); }
    if (ctx->s_copy) { cudaStreamSynchronize(ctx->s_copy); cudaStreamDestroy(ctx->s_copy); }
    if (ctx->s_up) { cudaStreamSynchronize(ctx->s_up); cudaStreamDestroy(ctx->s_up); }
    for (auto& b : ctx->pool)
        if (b.ptr) cudaFree(b.ptr);
    for (auto& b : ctx->hpool)
        if (b.ptr) cudaFreeHost(b.ptr);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
}

const char* grail_cuda_last_error(const grail_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

void* grail_cuda_stream_handle(grail_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int grail_cuda_synchronize(grail_ctx* ctx)
{
    if (!ctx) return GRAIL_ERR_INVALID_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->s_front));
    CU(ctx, cudaStreamSynchronize(ctx->s_back));
    CU(ctx, cudaStreamSynchronize(ctx->s_copy));
    CU(ctx, cudaStreamSynchronize(ctx->s_up));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    // everything has drained: the next pipelined launch of any plan is again "the first of a burst" and orders itself
    // after whatever the caller queues on the main stream from now on
    for (grail_plan* pl : ctx->live_plans) pl->in_flight = false;
    return GRAIL_OK;
}

int grail_cuda_set_option(grail_ctx* ctx, const char* key, double value)
{
    if (!ctx || !key) return GRAIL_ERR_INVALID_ARG;
    if (!strcmp(key, "warmup_nepers")) {
        if (!(value >= 0.0 && value <= 200.0)) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "warmup_nepers out of range");
        ctx->warmup_nepers = value;
    } else if (!strcmp(key, "target_lanes")) {
        ctx->target_items = value < 0 ? 0u : (uint32_t)value;
    } else if (!strcmp(key, "min_chunk")) {
        ctx->min_chunk = std::max(32u, (uint32_t)value);
    } else if (!strcmp(key, "max_chunk")) {
        ctx->max_chunk = std::max(32u, (uint32_t)value);
    } else if (!strcmp(key, "formants_per_lane")) {
        if (value != 1.0 && value != 2.0) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "formants_per_lane must be 1 or 2");
        ctx->formants_per_lane = (int)value;
    } else if (!strcmp(key, "pipeline")) {
        ctx->pipeline = value != 0.0;
    } else if (!strcmp(key, "phase_lean")) {
        ctx->phase_lean = value < 0.0 ? -1 : (value != 0.0);
    } else if (!strcmp(key, "e2e_groups")) {
        ctx->e2e_groups = value < 0.0 ? -1 : (int)value;
    } else if (!strcmp(key, "phase_mode")) {
        ctx->phase_mode = value == 0.0 ? 0 : (value == 3.0 ? 3 : (value == 2.0 ? 2 : 1));   // 3: fixed-point scan for a few long utterances
    } else if (!strcmp(key, "phase_chunk")) {
        if (value != 0.0 && !(value >= 256.0 && value <= 1048576.0)) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "phase_chunk out of range [256, 2^20] (0 = auto)");
        ctx->phase_chunk = (uint32_t)value;
    } else if (!strcmp(key, "phase_rounds")) {
        ctx->phase_rounds = value < 0.0 ? -1 : (value > PH_MAX_ROUNDS ? PH_MAX_ROUNDS : (int)value);
    } else if (!strcmp(key, "interleave")) {
        ctx->interleave = value != 0.0;
    } else if (!strcmp(key, "pscan_cost_model")) {
        ctx->pscan_cost_model = value != 0.0;
    } else if (!strcmp(key, "pscan_min_samples")) {
        ctx->pscan_min = value < 1.0 ? 1u : (value > 4.0e9 ? 0xFFFFFFFFu : (uint32_t)value);
    } else if (!strcmp(key, "zero_copy_out")) {
        ctx->zero_copy_out = value != 0.0;
    } else if (!strcmp(key, "debug_taps")) {
        ctx->debug_taps = value != 0.0;
    } else {
        return set_err(ctx, GRAIL_ERR_INVALID_ARG, "unknown option '%s'", key);
    }
    return GRAIL_OK;
}

int grail_cuda_host_alloc(grail_ctx* ctx, size_t bytes, void** out_ptr)
{
    if (!ctx || !out_ptr) return GRAIL_ERR_INVALID_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostAlloc(out_ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(ctx, GRAIL_ERR_OOM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return GRAIL_OK;
}

void grail_cuda_host_free(grail_ctx* ctx, void* ptr)
{
    (void)ctx;
    if (ptr) cudaFreeHost(ptr);
}

int grail_cuda_count_samples(const grail_seq_elem* elems, const uint32_t* utt_offsets,
                             const grail_voice_params* voices, uint32_t n_utts, uint64_t* counts)
{
    if (!counts) return GRAIL_ERR_INVALID_ARG;
    ElemInput in;
    in.full = elems;
    int rc = validate_inputs(nullptr, in, utt_offsets, voices, n_utts);
    if (rc) return rc;
    SeqCache cache;
    for (uint32_t u = 0; u < n_utts; ++u) {
        const int64_t n = schedule_utterance(elems + utt_offsets[u], utt_offsets[u + 1] - utt_offsets[u],
                                             voices[u].sample_rate, nullptr, cache);
        if (n < 0) return GRAIL_ERR_UNSUPPORTED;
        counts[u] = (uint64_t)n;
    }
    return GRAIL_OK;
}

int grail_cuda_plan_create(grail_ctx* ctx, const grail_seq_elem* elems, const uint32_t* utt_offsets,
                           const grail_voice_params* voices, uint32_t n_utts, grail_plan** out_plan)
{
    if (!ctx || !out_plan) return GRAIL_ERR_INVALID_ARG;
    *out_plan = nullptr;
    ElemInput in;
    in.full = elems;
    return plan_build(ctx, in, utt_offsets, voices, n_utts, out_plan, nullptr, true);
}

int grail_cuda_plan_join(grail_plan* plan)
{
    if (!plan) return GRAIL_ERR_INVALID_ARG;
    return plan_join(plan);
}

int grail_cuda_plan_create_phoneme_elems(grail_ctx* ctx, const grail_phoneme_elem* phonemes, const uint32_t* utt_offsets,
                                         const grail_elem* storages, uint32_t n_sounds, uint32_t n_storages,
                                         const uint32_t* utt_storage, const grail_voice_params* voices, uint32_t n_utts,
                                         grail_plan** out_plan)
{
    if (!ctx || !out_plan) return GRAIL_ERR_INVALID_ARG;
    *out_plan = nullptr;
    if (!phonemes && n_utts && utt_offsets && utt_offsets[n_utts] != 0) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "null input pointer");
    ElemInput in;
    static const grail_phoneme_elem none = { 0u, 0.0f, 0.0f, 0.0f };
    in.ph = phonemes ? phonemes : &none;
    in.storages = storages; in.n_sounds = n_sounds; in.n_storages = n_storages; in.utt_storage = utt_storage;
    return plan_build(ctx, in, utt_offsets, voices, n_utts, out_plan, nullptr, true);
}

int grail_cuda_plan_create_phonemes(grail_ctx* ctx, const uint8_t* phoneme_ids, const uint32_t* utt_offsets,
                                    const float* center_frequency, const grail_elem* storages, uint32_t n_sounds,
                                    uint32_t n_storages, const uint32_t* utt_storage, const grail_voice_params* voices,
                                    uint32_t n_utts, grail_plan** out_plan)
{
    if (!ctx || !out_plan) return GRAIL_ERR_INVALID_ARG;
    *out_plan = nullptr;
    if (!phoneme_ids && n_utts && utt_offsets && utt_offsets[n_utts] != 0) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "null input pointer");
    ElemInput in;
    static const uint8_t none = 0;
    in.ids = phoneme_ids ? phoneme_ids : &none;
    in.center = center_frequency;
    in.storages = storages; in.n_sounds = n_sounds; in.n_storages = n_storages; in.utt_storage = utt_storage;
    return plan_build(ctx, in, utt_offsets, voices, n_utts, out_plan, nullptr, true);
}

void grail_cuda_plan_destroy(grail_plan* plan)
{
    if (!plan) return;
    cudaSetDevice(plan->ctx->device);
    cudaStreamSynchronize(plan->ctx->s_front);
    cudaStreamSynchronize(plan->ctx->s_back);
    cudaStreamSynchronize(plan->ctx->stream);
    plan_release(plan);
}

uint64_t grail_cuda_plan_total_samples(const grail_plan* plan) { return plan ? plan->total_samples : 0; }

int grail_cuda_plan_out_offsets(const grail_plan* plan, uint64_t* out_offsets)
{
    if (!plan || !out_offsets) return GRAIL_ERR_INVALID_ARG;
    memcpy(out_offsets, plan->out_offsets.data(), plan->out_offsets.size() * sizeof(uint64_t));
    return GRAIL_OK;
}

int grail_cuda_plan_launch(grail_plan* plan, void* d_out, int format)
{
    if (!plan) return GRAIL_ERR_INVALID_ARG;
    if (!d_out && plan->total_samples) return set_err(plan->ctx, GRAIL_ERR_INVALID_ARG, "null output pointer");
    plan->out_channels = 1;
    return plan_enqueue(plan, d_out, format, false, true);
}

int grail_cuda_plan_launch_interleaved(grail_plan* plan, void* d_out, int format, uint32_t channels)
{
    if (!plan) return GRAIL_ERR_INVALID_ARG;
    if (channels < 1 || channels > 32) return set_err(plan->ctx, GRAIL_ERR_INVALID_ARG, "channels must be 1..32");
    if (!d_out && plan->total_samples) return set_err(plan->ctx, GRAIL_ERR_INVALID_ARG, "null output pointer");
    plan->out_channels = channels;
    const int rc = plan_enqueue(plan, d_out, format, false, true);
    plan->out_channels = 1;
    return rc;
}

int grail_cuda_plan_device_output(grail_plan* plan, int format, void** out_dptr)
{
    if (!plan || !out_dptr) return GRAIL_ERR_INVALID_ARG;
    grail_ctx* ctx = plan->ctx;
    const size_t need = (size_t)std::max<uint64_t>(plan->total_samples, 1) * format_bytes(format);
    if (!plan->d_out || plan->d_out_bytes < need) {
        if (plan->d_out) {
            int rcj = plan_join(plan);           // a pipelined formant kernel may still be writing the old buffer
            if (rcj) return rcj;
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            pool_free(ctx, plan->d_out);
            plan->d_out = nullptr;
        }
        int rc = pool_alloc(ctx, need, &plan->d_out);
        if (rc) return rc;
        plan->d_out_bytes = need;
    }
    plan->d_out_format = format;
    *out_dptr = plan->d_out;
    return GRAIL_OK;
}

int grail_cuda_plan_read_output(grail_plan* plan, int format, void* host_out)
{
    if (!plan || (!host_out && plan->total_samples)) return GRAIL_ERR_INVALID_ARG;
    grail_ctx* ctx = plan->ctx;
    if (!plan->d_out || plan->d_out_format != format)
        return set_err(ctx, GRAIL_ERR_INVALID_ARG, "plan has no device output in this format; launch into "
                                                   "grail_cuda_plan_device_output first");
    const size_t bytes = (size_t)plan->total_samples * format_bytes(format);
    int rcj = plan_join(plan);
    if (rcj) return rcj;
    if (bytes) CU(ctx, cudaMemcpyAsync(host_out, plan->d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return plan_check_device_errors(plan);
}

int grail_cuda_plan_timings(const grail_plan* plan, grail_timings* out)
{
    if (!plan || !out) return GRAIL_ERR_INVALID_ARG;
    memset(out, 0, sizeof *out);
    if (!plan->launched) return GRAIL_OK;
    grail_ctx* ctx = plan->ctx;
    const bool piped = plan->pipelined && plan->slot[plan->last_slot].used;
    const cudaEvent_t* ev = piped ? plan->slot[plan->last_slot].ev : plan->ev;
    CU(ctx, cudaEventSynchronize(ev[4]));
    CU(ctx, cudaEventElapsedTime(&out->schedule_ms, ev[0], ev[1]));
    CU(ctx, cudaEventElapsedTime(&out->frequency_ms, ev[1], ev[2]));
    if (piped) {   // ev[3] sits on the back stream there: phase = front-stream span minus the two kernels before it
        float front = 0.f;
        CU(ctx, cudaEventElapsedTime(&front, ev[2], ev[3]));
        out->phase_ms = front;   // upper bound: includes waiting for the previous formant kernel
    } else {
        CU(ctx, cudaEventElapsedTime(&out->phase_ms, ev[2], ev[3]));
    }
    CU(ctx, cudaEventElapsedTime(&out->formant_ms, ev[3], ev[4]));
    CU(ctx, cudaEventElapsedTime(&out->total_ms, ev[0], ev[4]));
    out->n_launches = plan->last_launches;
    return GRAIL_OK;
}

int grail_cuda_plan_phase_scan_stats(grail_plan* plan, uint32_t* stats)
{
    if (!plan || !stats) return GRAIL_ERR_INVALID_ARG;
    grail_ctx* ctx = plan->ctx;
    stats[0] = (uint32_t)plan->pscans.size();
    stats[1] = stats[2] = stats[3] = 0;
    if (plan->pscans.empty() || !plan->launched) return GRAIL_OK;
    int rcj = plan_join(plan);
    if (rcj) return rcj;
    std::vector<uint32_t> st(16 * plan->pscans.size());
    CU(ctx, cudaMemcpyAsync(st.data(), plan->d_pscan_status, st.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < plan->pscans.size(); ++i) {
        stats[1] += st[16 * i + 1] ? 1u : 0u;
        stats[2] = std::max(stats[2], st[16 * i + 2]);
        stats[3] += st[16 * i + 3] ? 1u : 0u;
        if (getenv("GRAIL_PSCAN_DEBUG")) {
            fprintf(stderr, "pscan %zu: done %u rounds %u refused %u mismatches per round:", i, st[16 * i + 1], st[16 * i + 2], st[16 * i + 3]);
            for (int r = 0; r < 11; ++r) fprintf(stderr, " %u", st[16 * i + 4 + r]);
            fprintf(stderr, "\n");
        }
    }
    return GRAIL_OK;
}

int grail_cuda_plan_phase_stats(grail_plan* plan, uint32_t* stats)
{
    if (!plan || !stats) return GRAIL_ERR_INVALID_ARG;
    grail_ctx* ctx = plan->ctx;
    memset(stats, 0, 8 * sizeof(uint32_t));
    stats[0] = plan->n_pchunks;
    stats[5] = plan->phase_chunk;
    if (!plan->launched) return GRAIL_OK;
    int rcj = plan_join(plan);
    if (rcj) return rcj;
    uint32_t st[16];
    CU(ctx, cudaMemcpyAsync(st, plan->d_pstats, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    stats[1] = st[PSTAT_WALKS];
    stats[2] = st[PSTAT_UNPROVEN];
    stats[3] = st[PSTAT_ROUNDS];
    stats[4] = st[PSTAT_MISMATCH];
    stats[6] = st[PSTAT_WARMUP_FROM_ZERO];
    return GRAIL_OK;
}

int grail_cuda_plan_read_intermediates(grail_plan* plan, float* frequency, float* carrier_phase, float* saw)
{
    if (!plan) return GRAIL_ERR_INVALID_ARG;
    grail_ctx* ctx = plan->ctx;
    if (!plan->d_phase_dbg) {
        void* p = nullptr;
        int rc = pool_alloc(ctx, plan->f_words * sizeof(float), &p);
        if (rc) return rc;
        plan->d_phase_dbg = (float*)p;
    }
    int rc = plan_enqueue(plan, nullptr, GRAIL_F32, true, false);
    if (rc) return rc;
    rc = plan_check_device_errors(plan);
    if (rc) return rc;
    std::vector<float> tmp(std::max<uint64_t>(plan->f_words, plan->saw_words));
    auto gather_linear = [&](const float* dsrc, float* dst) -> int {
        CU(ctx, cudaMemcpy(tmp.data(), dsrc, plan->f_words * sizeof(float), cudaMemcpyDeviceToHost));
        for (uint32_t u = 0; u < plan->n_utts; ++u) {
            const UttDev& U = plan->utts[u];
            memcpy(dst + U.out_off, tmp.data() + U.f_off, (size_t)U.n_samples * sizeof(float));
        }
        return GRAIL_OK;
    };
    auto gather_tiled = [&](const float* dsrc, float* dst) -> int {
        CU(ctx, cudaMemcpy(tmp.data(), dsrc, plan->saw_words * sizeof(float), cudaMemcpyDeviceToHost));
        const uint32_t CL = plan->chunk_len;
        for (uint32_t u = 0; u < plan->n_utts; ++u) {
            const UttDev& U = plan->utts[u];
            for (uint32_t n = 0; n < U.n_samples; ++n) {
                const uint32_t item = U.item_first + (n / CL) * U.item_stride, j = n % CL;
                const size_t idx = ((size_t)(item >> 5) * (CL >> 3) + (j >> 3)) * 256u + (item & 31u) * 8u + (j & 7u);
                dst[U.out_off + n] = tmp[idx];
            }
        }
        return GRAIL_OK;
    };
    if (frequency && (rc = plan->phase_chunk ? gather_tiled(plan->d_F, frequency) : gather_linear(plan->d_F, frequency))) return rc;
    if (carrier_phase && (rc = gather_linear(plan->d_phase_dbg, carrier_phase))) return rc;
    if (saw && (rc = gather_tiled(plan->slot[0].saw, saw))) return rc;
    return GRAIL_OK;
}

} // extern "C"

// One-shot batches with HOST output: the batch is cut into a few utterance groups of geometrically growing size
// (1 : 4 : 16), each its own plan; group k's device-to-host copy runs on the copy stream while group k+1's kernels run,
// and the host builds plan k+1 while the device works on k.  PCIe is the bound of this call (0.9 GB at ~52 GB/s = 17 ms
// against 3 ms of kernels at config 2), so what matters is that the first copy starts early and the copy engine never
// waits: the small first group gets the copy stream going after a fraction of a millisecond.  Groups are contiguous
// in utterance order, so every group's samples are one contiguous range of `out`.
static void split_groups(const uint64_t* out_offsets, uint32_t n_utts, int want, size_t elem_bytes, std::vector<uint32_t>& bounds)
{
    bounds.clear();
    bounds.push_back(0);
    const uint64_t total = out_offsets[n_utts] - out_offsets[0];
    int G = want;
    if (G < 0) G = (n_utts >= 96 && total * elem_bytes >= (48ull << 20)) ? 3 : 1;
    G = std::max(1, std::min(G, 8));
    if (G > 1) {
        double denom = 0.0, w = 1.0;
        for (int g = 0; g < G; ++g) { denom += w; w *= 4.0; }
        double acc = 0.0;
        w = 1.0;
        for (int g = 0; g + 1 < G; ++g) {
            acc += w; w *= 4.0;
            const uint64_t target = out_offsets[0] + (uint64_t)((double)total * acc / denom);
            uint32_t u = (uint32_t)(std::lower_bound(out_offsets, out_offsets + n_utts + 1, target) - out_offsets);
            u = std::min(std::max(u, bounds.back() + 1), n_utts);
            if (u > bounds.back() && u < n_utts) bounds.push_back(u);
        }
    }
    bounds.push_back(n_utts);
}

static int synthesize_batch_impl(grail_ctx* ctx, const grail_seq_elem* elems, const uint32_t* utt_offsets,
                                 const grail_voice_params* voices, uint32_t n_utts, int format, void* out,
                                 const uint64_t* out_offsets, int out_is_device)
{
    if (!ctx) return GRAIL_ERR_INVALID_ARG;
    if (!out_offsets) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "null out_offsets");
    if (!utt_offsets || (n_utts && !voices)) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "null input pointer");
    for (uint32_t u = 0; u < n_utts; ++u)
        if (utt_offsets[u + 1] < utt_offsets[u]) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "utt_offsets not monotone at %u", u);
    for (uint32_t u = 0; u < n_utts; ++u)
        if (out_offsets[u + 1] < out_offsets[u]) return set_err(ctx, GRAIL_ERR_COUNT_MISMATCH, "out_offsets not monotone at %u", u);
    const size_t eb = format_bytes(format);
    // A pinned (page-locked, device-mapped) host buffer can be written by k_formant directly (ctx option
    // "zero_copy_out"): measured on B200 39 ms vs 21 ms per config-2 step for the copy-engine path (16-byte-per-lane
    // stores use PCIe poorly), so it is opt-in.
    char* base = out ? (char*)out + out_offsets[0] * eb : nullptr;
    void* mapped = nullptr;
    if (!out_is_device && base && ctx->zero_copy_out) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, base) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            mapped = at.devicePointer;
        else
            cudaGetLastError();
    }
    const bool direct = out_is_device || mapped;            // the kernels write the caller's buffer themselves
    std::vector<uint32_t> bounds;
    split_groups(out_offsets, n_utts, direct ? 1 : ctx->e2e_groups, eb, bounds);
    std::vector<grail_plan*> plans;
    std::vector<cudaEvent_t> done;
    std::vector<uint32_t> offs;
    int rc = GRAIL_OK;
    // GRAIL_E2E_TRACE=1: per-group timeline on stderr (host time at enqueue; device times of kernels-done and copy-done)
    static const bool trace = getenv("GRAIL_E2E_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    std::vector<double> thost;
    const auto t_begin = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    if (trace) {
        cudaEvent_t e0; cudaEventCreate(&e0); cudaEventRecord(e0, ctx->stream); tev.push_back(e0);
    }
    for (size_t g = 0; g + 1 < bounds.size() && !rc; ++g) {
        const uint32_t u0 = bounds[g], u1 = bounds[g + 1], nu = u1 - u0;
        offs.resize(nu + 1);
        for (uint32_t u = 0; u <= nu; ++u) offs[u] = utt_offsets[u0 + u] - utt_offsets[u0];
        grail_plan* pl = nullptr;
        ElemInput in;
        in.full = elems ? elems + utt_offsets[u0] : nullptr;
        rc = plan_build(ctx, in, offs.data(), voices + u0, nu, &pl);
        if (rc) break;
        plans.push_back(pl);
        // the caller's layout must be the exact counts, packed in utterance order from out_offsets[0]
        for (uint32_t u = 0; u < nu && !rc; ++u) {
            if (out_offsets[u0 + u + 1] - out_offsets[u0 + u] != (uint64_t)pl->utts[u].n_samples) {
                const unsigned long long want = pl->utts[u].n_samples, got = out_offsets[u0 + u + 1] - out_offsets[u0 + u];
                rc = set_err(ctx, GRAIL_ERR_COUNT_MISMATCH, "utterance %u yields %llu samples, out_offsets leave room for %llu",
                             u0 + u, want, got);
            }
        }
        if (rc) break;
        if (pl->total_samples && !out) { rc = set_err(ctx, GRAIL_ERR_INVALID_ARG, "null output pointer"); break; }
        char* gbase = out ? (char*)out + out_offsets[u0] * eb : nullptr;
        if (direct) {
            rc = plan_enqueue(pl, mapped ? (void*)((char*)mapped + (gbase - base)) : (void*)gbase, format, false, true);
        } else {
            void* d = nullptr;
            rc = grail_cuda_plan_device_output(pl, format, &d);
            if (!rc) rc = plan_enqueue(pl, d, format, false, true);
            if (!rc && pl->total_samples) {
                cudaEvent_t ev = nullptr;
                cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
                if (e == cudaSuccess) { done.push_back(ev); e = cudaEventRecord(ev, ctx->stream); }
                if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_copy, ev, 0);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(gbase, d, (size_t)pl->total_samples * eb, cudaMemcpyDeviceToHost, ctx->s_copy);
                if (trace) {
                    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                    cudaEventRecord(a, ctx->stream); cudaEventRecord(b, ctx->s_copy);
                    tev.push_back(a); tev.push_back(b); thost.push_back(host_ms());
                }
                if (e != cudaSuccess) rc = set_err(ctx, GRAIL_ERR_CUDA, "device-to-host copy failed: %s", cudaGetErrorString(e));
            }
        }
    }
    // drain (also on the error paths: the plans' buffers go back to the pool)
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    cudaError_t ec = cudaStreamSynchronize(ctx->s_copy);
    if (!rc && (es != cudaSuccess || ec != cudaSuccess))
        rc = set_err(ctx, GRAIL_ERR_CUDA, "one-shot batch failed: %s", cudaGetErrorString(es != cudaSuccess ? es : ec));
    if (trace && !tev.empty()) {
        const double t_sync = host_ms();
        for (size_t g = 0; 2 * g + 2 < tev.size(); ++g) {
            float k = 0.f, c = 0.f;
            cudaEventElapsedTime(&k, tev[0], tev[2 * g + 1]);
            cudaEventElapsedTime(&c, tev[0], tev[2 * g + 2]);
            fprintf(stderr, "[grail e2e] group %zu (%u utts): enqueued at host %.2f ms, kernels done at %.2f ms, copy done at %.2f ms\n", g,
                    bounds[g + 1] - bounds[g], thost[g], k, c);
        }
        fprintf(stderr, "[grail e2e] drained at host %.2f ms\n", t_sync);
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    for (grail_plan* pl : plans) {
        if (!rc) rc = plan_check_device_errors(pl);
        plan_release(pl);
    }
    for (cudaEvent_t ev : done) cudaEventDestroy(ev);
    return rc;
}

extern "C" {

int grail_cuda_synthesize_batch(grail_ctx* ctx, const grail_seq_elem* elems, const uint32_t* utt_offsets,
                                const grail_voice_params* voices, uint32_t n_utts, float* out,
                                const uint64_t* out_offsets, int out_is_device)
{
    return synthesize_batch_impl(ctx, elems, utt_offsets, voices, n_utts, GRAIL_F32, out, out_offsets, out_is_device);
}

int grail_cuda_synthesize_batch_i16(grail_ctx* ctx, const grail_seq_elem* elems, const uint32_t* utt_offsets,
                                    const grail_voice_params* voices, uint32_t n_utts, int16_t* out,
                                    const uint64_t* out_offsets, int out_is_device)
{
    return synthesize_batch_impl(ctx, elems, utt_offsets, voices, n_utts, GRAIL_I16, out, out_offsets, out_is_device);
}

// ---- streaming ---------------------------------------------------------------------------------
// A grail_stream is the Copy-able state of the reference's three iterators: Sequencer
// {cur, next, time}, Jitter {value-noise phase, wrap count, seed}, Synthesize {carrier phase, a, b, c, noise draw
// index}.  Each pull synthesizes one window of the unbounded utterance with the batch kernels, entering them with
// that state and reading it back afterwards, so a stream pulled in any window sizes equals the one-shot result
// (bit-exact clocks / phase; filters continue from their exact states, no warm-up at the window edge).
} // extern "C"

struct grail_stream {
    grail_ctx* ctx = nullptr;
    grail_voice_params voice{};
    std::vector<grail_seq_elem> pending;   // [0] = the phoneme in progress (or the next one to start)
    bool finished = false, ended = false;
    StreamStart st;                        // state entering the next window
    float filter[24];
};

extern "C" {

int grail_cuda_stream_new(grail_ctx* ctx, const grail_voice_params* voice, grail_stream** out_stream)
{
    if (!ctx || !voice || !out_stream) return GRAIL_ERR_INVALID_ARG;
    *out_stream = nullptr;
    const uint32_t offs[2] = { 0, 0 };
    int rc = validate_inputs(ctx, ElemInput(), offs, voice, 1);
    if (rc) return rc;
    grail_stream* s = new (std::nothrow) grail_stream();
    if (!s) return set_err(ctx, GRAIL_ERR_OOM, "host allocation failed");
    s->ctx = ctx;
    s->voice = *voice;
    memset(s->filter, 0, sizeof s->filter);
    *out_stream = s;
    return GRAIL_OK;
}

int grail_cuda_stream_push(grail_stream* s, const grail_seq_elem* elems, uint32_t n_elems)
{
    if (!s || (n_elems && !elems)) return GRAIL_ERR_INVALID_ARG;
    if (s->finished) return set_err(s->ctx, GRAIL_ERR_INVALID_ARG, "stream already finished");
    for (uint32_t i = 0; i < n_elems; ++i)
        if (!std::isfinite(elems[i].length)) return set_err(s->ctx, GRAIL_ERR_INVALID_ARG, "non-finite phoneme length");
    s->pending.insert(s->pending.end(), elems, elems + n_elems);
    return GRAIL_OK;
}

int grail_cuda_stream_finish(grail_stream* s)
{
    if (!s) return GRAIL_ERR_INVALID_ARG;
    s->finished = true;
    return GRAIL_OK;
}

// The windows of n streams as ONE plan and one launch per kernel: utterance k of the plan is stream k's next window.
// A server that runs many voices at once (examples/interactive.rs duplicates one stream over the output channels; a
// speech service runs one per client) pays the launch and synchronisation latency once per tick, not once per stream.
static int streams_pull_impl(grail_stream* const* streams, uint32_t n_streams, float* const* outs, const uint64_t* max_samples,
                             uint64_t* n_written)
{
    grail_ctx* ctx = streams[0]->ctx;
    std::vector<uint32_t> live;                  // streams that have something to synthesize in this call
    for (uint32_t k = 0; k < n_streams; ++k) {
        grail_stream* s = streams[k];
        n_written[k] = 0;
        if (s->ctx != ctx) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "streams of one call must share a ctx");
        if (s->ended || max_samples[k] == 0 || s->pending.empty()) {
            if (s->finished && s->pending.empty()) s->ended = true;
            continue;
        }
        if (!s->finished && s->pending.size() < 2) continue;   // the only element is still just a look-ahead
        if (!outs[k]) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "null output pointer");
        live.push_back(k);
    }
    if (live.empty()) return GRAIL_OK;
    const uint32_t nl = (uint32_t)live.size();
    std::vector<grail_seq_elem> elems;
    std::vector<uint32_t> offs(nl + 1, 0);
    std::vector<grail_voice_params> voices(nl);
    std::vector<StreamStart> ss(nl);
    // Only the phonemes a window can reach go into the plan (a stream that was pushed a whole sentence holds dozens of
    // 208-byte records; a 10 ms window needs one or two): the phoneme in progress counts for nothing, every further one
    // for a little less than length * sample_rate, until the window is covered -- plus one more as the look-ahead.  A
    // shortened list is planned as "unfinished" (its last record is held back like any look-ahead).  Should a window
    // still come out short (lengths that are not numbers), the call is planned again with everything.
    auto reach = [&](const grail_stream* s, uint64_t want) -> size_t {
        const size_t np = s->pending.size();
        const double rate = (double)s->voice.sample_rate;
        double acc = 0.0;
        size_t k = 1;
        for (; k < np && acc < (double)want; ++k) {
            const double n = (double)s->pending[k].length * rate * 0.999 - 4.0;
            if (n > 0.0) acc += n;             // (NaN compares false: counts for nothing)
        }
        return std::min(np, k + 1);
    };
    grail_plan* pl = nullptr;
    int rc = GRAIL_OK;
    std::vector<uint8_t> cut(nl, 0);
    for (int attempt = 0; attempt < 2; ++attempt) {
        elems.clear();
        for (uint32_t i = 0; i < nl; ++i) {
            grail_stream* s = streams[live[i]];
            const uint64_t want = std::min<uint64_t>(max_samples[live[i]], MAX_UTT_SAMPLES);
            const size_t take = attempt == 0 ? reach(s, want) : s->pending.size();
            cut[i] = take < s->pending.size();
            elems.insert(elems.end(), s->pending.begin(), s->pending.begin() + take);
            offs[i + 1] = (uint32_t)elems.size();
            voices[i] = s->voice;
            ss[i] = s->st;
            ss[i].filter_state = s->st.fresh ? nullptr : s->filter;
            ss[i].max_samples = want;
            ss[i].finished = s->finished && !cut[i];
        }
        ElemInput in;
        in.full = elems.data();
        rc = plan_build(ctx, in, offs.data(), voices.data(), nl, &pl, ss.data());
        if (rc) return rc;
        bool short_window = false;
        for (uint32_t i = 0; i < nl && !short_window; ++i)
            short_window = cut[i] && pl->utts[i].n_samples < ss[i].max_samples;
        if (!short_window) break;
        cudaStreamSynchronize(ctx->stream);
        plan_release(pl);
        pl = nullptr;
    }
    if (!pl) return set_err(ctx, GRAIL_ERR_CUDA, "stream window could not be planned");
    std::vector<float> fin((size_t)nl * 32, 0.0f);
    if (pl->total_samples) {
        void* d = nullptr;
        rc = grail_cuda_plan_device_output(pl, GRAIL_F32, &d);
        if (!rc) rc = plan_enqueue(pl, d, GRAIL_F32, false, true);
        // Read-back: ONE copy of the filter states and ONE of the packed windows into a pinned block, then host copies
        // into the callers' buffers.  (A cudaMemcpyAsync per stream into pageable memory is a synchronous staged copy
        // each: 16-20 us per stream and tick, which capped a tick of 1 024 streams at 20 ms.)
        const size_t fin_bytes = fin.size() * sizeof(float), out_bytes = (size_t)pl->total_samples * sizeof(float);
        void* hst = nullptr;
        if (!rc) rc = hpool_alloc(ctx, fin_bytes + out_bytes + 256, &hst);
        cudaError_t e = cudaSuccess;
        if (!rc) e = cudaMemcpyAsync(hst, pl->d_utt_final, fin_bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (!rc && e == cudaSuccess) e = cudaMemcpyAsync((char*)hst + fin_bytes, d, out_bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (!rc && e != cudaSuccess) rc = set_err(ctx, GRAIL_ERR_CUDA, "stream read-back failed: %s", cudaGetErrorString(e));
        if (!rc) rc = plan_check_device_errors(pl);             // synchronizes the stream
        if (rc) {
            cudaStreamSynchronize(ctx->stream);
            hpool_free(ctx, hst);
            plan_release(pl);
            return rc;
        }
        memcpy(fin.data(), hst, fin_bytes);
        const float* hout = reinterpret_cast<const float*>((const char*)hst + fin_bytes);
        for (uint32_t i = 0; i < nl; ++i) {
            const uint64_t n = pl->utts[i].n_samples;
            if (n) memcpy(outs[live[i]], hout + pl->out_offsets[i], n * sizeof(float));
        }
        hpool_free(ctx, hst);
    }
    // ---- every stream's iterator state after its last sample, from the exact host-side schedules
    for (uint32_t i = 0; i < nl; ++i) {
        grail_stream* s = streams[live[i]];
        const UttDev& U = pl->utts[i];
        const uint64_t n = U.n_samples;
        if (n == 0) {
            if (s->finished) { s->ended = true; s->pending.clear(); }
            continue;
        }
        const SegRec* segs = pl->segs.data() + U.elem_first;
        const float dt = sdiv(1.0f, s->voice.sample_rate);
        const uint32_t last = (uint32_t)(n - 1);
        uint32_t p = 0;
        while (p + 1 < U.n_elems && segs[p + 1].start <= last) ++p;
        const float time_last = clock_desc_run(segs[p].time0, dt, last - segs[p].start).x;
        const JitSchedDev& js = pl->jscheds[U.jit_sched];
        uint32_t w = 0;
        while (w + 1 < js.n_recs && pl->jrecs[js.rec_first + w + 1].n <= (int32_t)last) ++w;
        const JitRec& jr = pl->jrecs[js.rec_first + w];
        const float* f = fin.data() + (size_t)i * 32;
        StreamStart nx;
        nx.fresh = false;
        nx.jitter_phase = clock_asc_run(jr.phase, s->voice.jitter_frequency, (uint64_t)((int64_t)last - jr.n)).x;
        nx.jitter_wraps = s->st.jitter_wraps + w;
        nx.sample0 = s->st.sample0 + n;
        nx.carrier_phase = f[24];
        memcpy(s->filter, f, sizeof s->filter);
        // Sequencer: what the next call to next() will do with `time` (src/lib.rs:861-888)
        const float t_next = ssub(time_last, dt);
        uint32_t consumed;
        if (t_next < 0.0f) {            // the hand-over happens on the next sample: phoneme p is done
            nx.cont_phoneme = false;
            nx.t_neg = t_next;
            consumed = p + 1;
        } else {                        // phoneme p continues
            nx.cont_phoneme = true;
            nx.time0 = t_next;
            consumed = p;
        }
        s->pending.erase(s->pending.begin(), s->pending.begin() + consumed);
        s->st = nx;
        if (s->finished && s->pending.empty()) s->ended = true;
        n_written[live[i]] = n;
    }
    plan_release(pl);
    return GRAIL_OK;
}

int grail_cuda_stream_pull(grail_stream* s, float* out, uint64_t max_samples, uint64_t* n_written)
{
    if (!s || !n_written || (max_samples && !out)) return GRAIL_ERR_INVALID_ARG;
    return streams_pull_impl(&s, 1, &out, &max_samples, n_written);
}

int grail_cuda_streams_pull(grail_stream* const* streams, uint32_t n_streams, float* const* outs, const uint64_t* max_samples,
                            uint64_t* n_written)
{
    if (n_streams == 0) return GRAIL_OK;
    if (!streams || !outs || !max_samples || !n_written) return GRAIL_ERR_INVALID_ARG;
    for (uint32_t k = 0; k < n_streams; ++k)
        if (!streams[k]) return GRAIL_ERR_INVALID_ARG;
    return streams_pull_impl(streams, n_streams, outs, max_samples, n_written);
}

void grail_cuda_stream_free(grail_stream* s) { delete s; }

// ---- roofline probes ---------------------------------------------------------------------------
int grail_cuda_copy_segments(grail_ctx* ctx, void* dst, const void* src, const uint64_t* dst_off, const uint64_t* src_off,
                             const uint64_t* len, uint64_t n_segments, uint32_t elem_bytes)
{
    if (!ctx) return GRAIL_ERR_INVALID_ARG;
    if (elem_bytes != 2 && elem_bytes != 4) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "elem_bytes must be 2 or 4");
    if (n_segments == 0) return GRAIL_OK;
    if (!dst || !src || !dst_off || !src_off || !len) return set_err(ctx, GRAIL_ERR_INVALID_ARG, "null pointer");
    CU(ctx, cudaSetDevice(ctx->device));
    std::vector<SegCopy> segs(n_segments);
    std::vector<uint32_t> tile_seg, tile_idx;
    for (uint64_t i = 0; i < n_segments; ++i) {
        segs[i].dst = dst_off[i] * elem_bytes; segs[i].src = src_off[i] * elem_bytes; segs[i].len = len[i] * elem_bytes;
        const uint64_t tiles = (segs[i].len + SEG_TILE_BYTES - 1) / SEG_TILE_BYTES;
        if (i > 0xFFFFFFF0ull || tiles > 0xFFFFFFF0ull || tile_seg.size() + tiles > 0x7FFFFFF0ull)
            return set_err(ctx, GRAIL_ERR_UNSUPPORTED, "too many segments / tiles for one call");
        for (uint64_t t = 0; t < tiles; ++t) { tile_seg.push_back((uint32_t)i); tile_idx.push_back((uint32_t)t); }
    }
    if (tile_seg.empty()) return GRAIL_OK;
    void *d_segs = nullptr, *d_ts = nullptr, *d_ti = nullptr;
    int rc = pool_alloc(ctx, segs.size() * sizeof(SegCopy), &d_segs);
    if (!rc) rc = pool_alloc(ctx, tile_seg.size() * 4, &d_ts);
    if (!rc) rc = pool_alloc(ctx, tile_idx.size() * 4, &d_ti);
    if (!rc) {
        cudaStream_t s = ctx->stream;
        cudaError_t e = cudaMemcpyAsync(d_segs, segs.data(), segs.size() * sizeof(SegCopy), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ts, tile_seg.data(), tile_seg.size() * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ti, tile_idx.data(), tile_idx.size() * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) {
            k_copy_segments<<<(unsigned)tile_seg.size(), 256, 0, s>>>((unsigned char*)dst, (const unsigned char*)src,
                                                                       (const SegCopy*)d_segs, (const uint32_t*)d_ts,
                                                                       (const uint32_t*)d_ti, elem_bytes);
            e = cudaGetLastError();
        }
        // the tables are pageable host vectors: the copies above have been staged by the time the calls return, but the
        // scratch buffers go back to the pool, so the launch must have consumed them first
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = set_err(ctx, GRAIL_ERR_CUDA, "segment copy failed: %s", cudaGetErrorString(e));
    }
    pool_free(ctx, d_segs); pool_free(ctx, d_ts); pool_free(ctx, d_ti);
    return rc;
}

int grail_cuda_probe_fp32_peak(grail_ctx* ctx, double* ffma_flops, double* mufu_ops, double* sm_mhz_effective)
{
    if (!ctx) return GRAIL_ERR_INVALID_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->prop.multiProcessorCount * 8, threads = 256;
    float* sink = nullptr;
    void* p = nullptr;
    int rc = pool_alloc(ctx, (size_t)blocks * threads * sizeof(float), &p);
    if (rc) return rc;
    sink = (float*)p;
    cudaEvent_t e0, e1;
    CU(ctx, cudaEventCreate(&e0));
    CU(ctx, cudaEventCreate(&e1));
    cudaStream_t s = ctx->stream;
    float ms = 0.f;
    double best_f = 0, best_m = 0;
    for (int rep = 0; rep < 4; ++rep) {
        const int iters = 4096;
        CU(ctx, cudaEventRecord(e0, s));
        k_probe_ffma<<<blocks, threads, 0, s>>>(sink, iters, 1.0f);
        CU(ctx, cudaEventRecord(e1, s));
        CU(ctx, cudaEventSynchronize(e1));
        CU(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double fl = (double)blocks * threads * iters * 16.0 * 8.0 * 2.0 / (ms * 1e-3);
        if (rep > 0) best_f = std::max(best_f, fl);
        const int miters = 1024;
        CU(ctx, cudaEventRecord(e0, s));
        k_probe_mufu<<<blocks, threads, 0, s>>>(sink, miters, 1.0f);
        CU(ctx, cudaEventRecord(e1, s));
        CU(ctx, cudaEventSynchronize(e1));
        CU(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double mo = (double)blocks * threads * miters * 16.0 * 4.0 / (ms * 1e-3);
        if (rep > 0) best_m = std::max(best_m, mo);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    pool_free(ctx, sink);
    if (ffma_flops) *ffma_flops = best_f;
    if (mufu_ops) *mufu_ops = best_m;
    if (sm_mhz_effective) // FFMA peak = SMs x 128 lanes x 2 flop x f  =>  the clock the probe actually ran at
        *sm_mhz_effective = best_f / ((double)ctx->prop.multiProcessorCount * 128.0 * 2.0) * 1e-6;
    return GRAIL_OK;
}

// ---- debug hooks for the exact-clock tests (host only) -------------------------------------------
void grail_cuda_debug_clock_desc(float x, float d, uint64_t max_steps, float* x_out, uint64_t* steps, int* stuck)
{
    const ClockRun r = clock_desc_run(x, d, max_steps);
    *x_out = r.x; *steps = r.steps; *stuck = r.stuck;
}
void grail_cuda_debug_clock_asc(float x, float d, uint64_t max_steps, float* x_out, uint64_t* steps, int* stuck)
{
    const ClockRun r = clock_asc_run(x, d, max_steps);
    *x_out = r.x; *steps = r.steps; *stuck = r.stuck;
}
uint32_t grail_cuda_debug_lcg_jump(uint32_t seed, uint64_t n) { return lcg_jump(seed, n); }
int grail_cuda_debug_div_check(grail_ctx* ctx, uint32_t seed, uint64_t n_pairs, uint64_t* mismatches)
{
    if (!ctx || !mismatches) return GRAIL_ERR_INVALID_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    unsigned long long* d = nullptr;
    CU(ctx, cudaMalloc(&d, 8));
    CU(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
    const uint32_t per_thread = 1024, threads = 256;
    const uint64_t blocks = (n_pairs + (uint64_t)per_thread * threads - 1) / ((uint64_t)per_thread * threads);
    k_debug_div_check<<<(unsigned)std::min<uint64_t>(blocks, 1u << 20), threads, 0, ctx->stream>>>(seed, per_thread, d);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return set_err(ctx, GRAIL_ERR_CUDA, "div check: %s", cudaGetErrorString(e));
    *mismatches = h;
    return GRAIL_OK;
}
uint64_t grail_cuda_debug_jitter_index(int gen, int which /*0 cur, 1 next*/, int i, uint64_t w)
{
    if (gen < 0) return which ? jit_freq_next_idx(w) : jit_freq_cur_idx(w);
    return which ? jit_arr_next_idx(gen, i, w) : jit_arr_cur_idx(gen, i, w);
}

} // extern "C"
