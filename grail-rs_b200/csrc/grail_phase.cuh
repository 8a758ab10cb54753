// grail_phase.cuh -- the carrier phase `phase <- RN(phase + F_t); if phase >= 1 { phase -= 1 }` (src/lib.rs:520-525),
// bit-exact AND parallel in time.  Included by grail_kernels.cuh.
//
// The f32 accumulator is not associative, so a prefix sum of F_t is not the reference's phase (58 dB).
// What IS true of the reference's arithmetic:
//   * every carrier wrap leaves a phase that is a multiple of 2^-23 (the grid of [1, 2) minus one);
//   * from there on, two trajectories that differ by an even multiple of 2^-23 stay exactly that far apart: all
//     later grids are no coarser than 2^-23 and no finer grid's rounding can tell them apart, round-half-even ties
//     included (an odd multiple flips the parity that decides a tie on a wrap step) -- unless one of the two reaches
//     a binade boundary or 1.0 a step earlier than the other, which needs the boundary to fall inside that small gap.
// So an utterance is cut into phase chunks of at most PC samples, one lane each, and
//   k_phase_guess   start-phase guesses from the exact sums of F_t (k_frequency emits one sum per 256 samples); they
//                   ignore the accumulator's rounding and are off by tens of 2^-23 units after a few seconds;
//   k_phase_a       walks every chunk literally from its guess to its first wrap (the anchor), then, as two
//                   trajectories one unit apart, on to the first wrap inside the NEXT chunk, where that chunk's own
//                   anchor is: the difference there is how far apart the two chunks' guesses were;
//   k_phase_scan_a  integer prefix sum of those differences along each utterance (choosing per chunk the trajectory
//                   at an even distance from the truth): every chunk's start phase, exact unless a trajectory
//                   crossed a boundary a step early or late;
//   k_phase_b       walks every chunk literally from that start: the polyBLEP saw, and the chunk's end phase;
//   k_phase_fix     the PROOF: chunk c's end must equal chunk c+1's start bit for bit for every c; by induction from the
//                   exact phase at sample 0 the whole trajectory is then the reference's.  Where it is not, the
//                   starts downstream are shifted by the observed difference and those chunks walked again
//                   (k_phase_b on the dirty chunks); measured 1.0-1.2 walks per chunk.
// An utterance that is still unproven after the last round keeps status bit 0 clear and k_phase_pair runs the
// serial chain for it.  Nothing here assumes anything about F_t: negative, NaN or tiny increments only make the
// guesses worse, the proof decides.
//
// Memory: F_t lives in the SAME tiled layout as the saw, [group of 32 work items][j / 8][lane][8] (saw_index), and a
// warp of the walks is the same sub-range of 32 work items of one k_formant group: lane l reads 32 bytes of item l per
// 8-sample block, the warp 1 KB of consecutive addresses, and round B stores the saw to the very same offsets.  Both
// rounds are plain streaming kernels bound by HBM (round A reads F_t about 1.2 times, round B reads it once and
// writes the saw once); the loads run eight blocks ahead of the arithmetic through a cp.async ring in shared memory.
// (First version, kept in profiles/r2_phase_strided_launches.csv: linear F_t, one lane per CONSECUTIVE chunk of an
//  utterance, i.e. 32 lanes reading 32 places 8 KB apart -- per-lane 128-byte TMA bulk copies (cp.async.bulk +
//  mbarrier complete_tx) into private shared-memory rows as well as LDG.256 / 16-byte cp.async all ended at 0.84 ms
//  for round A and 0.85 ms for round B at config 2, bound by one L1 tag look-up / one TMA request per 32-byte
//  sector: the access pattern, not the copy engine, was the problem.)
#pragma once

namespace grail {

// Per-chunk records, one array per field (the scans read a field of 32 consecutive chunks with one coalesced load):
//   field f of chunk g is P.pchunks[f * P.pc_stride + g] (float or 32-bit integer bits)
enum { PCF_START = 0,   // claimed exact phase before the chunk's first sample (guess, then scan output)
       PCF_END = 1,     // phase after the chunk's last sample, walked from START (k_phase_b)
       PCF_FLAGS = 2,   // PCH_*
       PCF_A = 3,       // k_phase_a: sample whose step wrapped first inside this chunk (-1: none)
       PCF_R = 4,       //            phase after that step
       PCF_E0 = 5,      //            phase at the chunk's end of the two trajectories (R and R + 2^-23 at the anchor)
       PCF_E1 = 6,
       PCF_A20 = 7,     //            first wrapping step at or after the chunk's end (-1: none inside the next chunk)
       PCF_A21 = 8,
       PCF_R20 = 9,     //            phase after it
       PCF_R21 = 10,
       PCF_TIEKEY = 11, // first round-half-even tie on a wrap step inside the chunk: (sample << 2) | 1 lower / 2 upper candidate taken; -1: none
       PCF_COUNT = 12 };
enum : uint32_t { PCH_DIRTY = 1u, PCH_TIE_LOWER = 2u, PCH_TIE_UPPER = 4u };
// pstats words
enum { PSTAT_WALKS = 0,       // chunks walked by k_phase_b, all rounds
       PSTAT_UNPROVEN = 1,    // utterances left to the serial chain
       PSTAT_ROUNDS = 2,      // last repair round that still found a mismatch (0: the first walk proved everything)
       PSTAT_MISMATCH = 3,    // chunk boundaries that failed the proof, all rounds
       PSTAT_WARMUP_FROM_ZERO = 8,   // k_formant: (chunk, formant) pairs whose filter warm-up reached back to sample 0
       PSTAT_PENDING = 16 };  // + r: dirty chunks waiting for round r's k_phase_b

constexpr int PH_MAX_ROUNDS = 14;
constexpr int PH_OCC = 6;      // CTAs of 128 lanes per SM the walks are compiled for (<= 85 registers, 32 KB of ring each)
constexpr float PH_U23 = 1.1920928955078125e-07f;   // 2^-23

__device__ __forceinline__ float* pcf(const PlanDev& P, int field) { return P.pchunks + (size_t)field * P.pc_stride; }
__device__ __forceinline__ int32_t* pci(const PlanDev& P, int field) { return reinterpret_cast<int32_t*>(pcf(P, field)); }

__device__ __forceinline__ void ldg256(const float* p, float (&v)[8])
{
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float (&v)[8])
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// one literal step, src/lib.rs:520-525, without a branch: `if phase >= 1 { phase -= 1 }` as phase - (phase >= 1 ? 1.0 : 0.0);
// x - 0.0 is x bit for bit, and a NaN compares false like the reference's `>=`.  g comes back 1.0 on a wrap.
__device__ __forceinline__ float phase_step(float p, float f, float& g)
{
    const float q = sadd(p, f);
    g = ge_one(q);
    return ssub(q, g);
}

__device__ __forceinline__ double frac_d(double x) { return x - floor(x); }
// x - y for two phases, as the representative in [-0.5, 0.5)
__device__ __forceinline__ double centered_diff(float x, float y)
{
    double d = (double)x - (double)y;
    if (d >= 0.5) d -= 1.0;
    if (d < -0.5) d += 1.0;
    return d;
}

// ------------------------------------------------------------------------------------------------
// Phase chunks.  Work item (time chunk of k_formant, CL samples) i of an utterance is cut into K = ceil(CL / PC)
// phase chunks of PC samples (the last one of an item shorter when PC does not divide CL); an utterance's chunks are
// numbered in time order, chunk c = item * K + q, and stored at pc_first + c.  A warp of the walks is sub-range q of
// the 32 items of one k_formant group.
// ------------------------------------------------------------------------------------------------
struct PChunkLane {
    uint32_t u;          // utterance
    uint32_t g, c, C;    // global chunk index, index within the utterance, chunks of the utterance
    uint32_t n0, n1;     // samples [n0, n1) of the utterance
    uint32_t nn1;        // end of the NEXT chunk of the utterance (n1 when there is none)
    bool     valid;
};
__device__ __forceinline__ PChunkLane pchunk_of_lane(const PlanDev& P)
{
    PChunkLane X;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    const uint32_t K = P.pc_per_item, PC = P.phase_chunk, CL = P.chunk_len;
    const uint32_t group = warp / K, q = warp - group * K;
    const uint32_t item = group * 32u + lane;
    X.valid = false;
    X.u = 0; X.g = X.c = X.C = X.n0 = X.n1 = X.nn1 = 0;
    if (item >= P.n_items) return X;
    const ItemDev it = P.items[item];
    const uint32_t s0 = q * PC;
    if (s0 >= it.len) return X;
    const UttDev& U = P.utts[it.utt];
    X.valid = true;
    X.u = it.utt;
    X.C = U.pc_count;
    X.c = (it.n0 / CL) * K + q;
    X.g = U.pc_first + X.c;
    X.n0 = it.n0 + s0;
    X.n1 = it.n0 + min(min(s0 + PC, CL), it.len);
    const uint32_t n = U.n_samples;
    uint32_t len2 = 0;
    if (X.n1 < n) {
        const uint32_t j1 = X.n1 - (X.n1 / CL) * CL;              // position of the next chunk inside its item
        len2 = min(min(PC, CL - j1), n - X.n1);
    }
    X.nn1 = X.n1 + len2;
    return X;
}

// the same record for global chunk id g (repair rounds: the dirty chunks come as a dense list, in no particular order)
__device__ __forceinline__ PChunkLane pchunk_of_id(const PlanDev& P, uint32_t g)
{
    PChunkLane X;
    uint32_t lo = 0, hi = P.n_utts;    // last utterance with pc_first <= g (utterances without chunks share their successor's pc_first)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (P.utts[mid].pc_first <= g) lo = mid; else hi = mid;
    }
    const UttDev& U = P.utts[lo];
    const uint32_t K = P.pc_per_item, PC = P.phase_chunk, CL = P.chunk_len, n = U.n_samples;
    X.valid = true;
    X.u = lo;
    X.g = g;
    X.C = U.pc_count;
    X.c = g - U.pc_first;
    const uint32_t ci = X.c / K, q = X.c - ci * K;
    X.n0 = ci * CL + q * PC;
    X.n1 = min(ci * CL + min(q * PC + PC, CL), n);
    X.nn1 = X.n1;
    return X;
}

// Cursor over an utterance's F_t / saw in the tiled layout, one 8-sample block at a time.
struct TileCursor {
    size_t   off;        // float offset of the current block (of this utterance's sample `t`)
    uint32_t j;          // position inside the item, multiple of 8
    uint32_t item;
    uint32_t CL, stride;
    __device__ __forceinline__ void seek(const UttDev& U, uint32_t t, uint32_t cl)
    {
        CL = cl; stride = U.item_stride;
        const uint32_t ci = t / cl;
        item = U.item_first + ci * stride;
        j = t - ci * cl;
        off = saw_index(item, j, cl);
    }
    __device__ __forceinline__ void next()
    {
        j += 8u;
        if (j >= CL) { j = 0u; item += stride; off = saw_index(item, 0u, CL); }
        else off += 256u;
    }
};

// fn(blk, off, f) for every 8-sample block of [t0, t1) (both multiples of 8) in order; `off` is the block's offset in
// the tiled arrays; fn returns false to stop early.
// The stream runs through a per-lane ring of PH_DEPTH 32-byte slots in shared memory filled with cp.async (LDGSTS):
// groups complete in order, so `wait_group PH_DEPTH-1` is an exact, per-lane "block i has landed" with PH_DEPTH-1 more
// blocks in flight behind it.  (First version: the next four blocks in registers, LDG.256 -- ncu showed the walks
// parked on the register copies of three of the four ring slots, long_scoreboard 7.6 per issue: ptxas folds the ring's
// loads onto shared scoreboard entries, which turns "four blocks ahead" into "one block ahead", 1 450 cycles per block.)
// Slots are private to their lane (no synchronisation, lanes may diverge and stop at different blocks); the two
// 16-byte halves of a slot are swapped for every other group of four lanes so that the 128-bit reads of a warp fall
// on all 32 banks.
constexpr int PH_DEPTH = 8;
constexpr unsigned PH_STAGE_BYTES = 128u * 32u;             // one block of every lane of the CTA
constexpr unsigned PH_RING_BYTES = PH_DEPTH * PH_STAGE_BYTES;

template <class Fn>
__device__ __forceinline__ void walk_blocks(unsigned char* ring, const float* __restrict__ F, const UttDev& U, uint32_t CL,
                                            uint32_t t0, uint32_t t1, Fn&& fn)
{
    if (t0 >= t1) return;
    TileCursor ld, cur;
    ld.seek(U, t0, CL);
    cur = ld;
    const unsigned swz = ((threadIdx.x >> 2) & 1u) << 4;
    const unsigned slot = (unsigned)__cvta_generic_to_shared(ring) + threadIdx.x * 32u;
    uint32_t tl = t0, is = 0;                       // next block to load and its stage
    auto issue = [&]() {
        if (tl < t1) {
            const unsigned a = slot + is * PH_STAGE_BYTES;
            cp_async16(a + swz, F + ld.off);
            cp_async16(a + (swz ^ 16u), F + ld.off + 4);
            ld.next();
            tl += 8u;
        }
        cp_async_commit();                          // (an empty group when the stream is over: the count stays uniform)
        is = (is + 1u) & (PH_DEPTH - 1u);
    };
#pragma unroll
    for (int i = 0; i < PH_DEPTH - 1; ++i) issue();
    // the block being stepped sits in registers, and the NEXT one is read out of the ring before the stepping starts,
    // so the shared-memory latency hides behind the dependent chain instead of heading every iteration
    issue();
    cp_async_wait<PH_DEPTH - 1>();
    float4 na = lds128(slot + swz), nb = lds128(slot + (swz ^ 16u));
    uint32_t cs = 1;
#pragma unroll 1
    for (uint32_t blk = t0; blk < t1; blk += 8u) {
        const float f[8] = { na.x, na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w };
        if (blk + 8u < t1) {
            issue();
            cp_async_wait<PH_DEPTH - 1>();
            const unsigned a = slot + cs * PH_STAGE_BYTES;
            na = lds128(a + swz);
            nb = lds128(a + (swz ^ 16u));
            cs = (cs + 1u) & (PH_DEPTH - 1u);
        }
        const bool go = fn(blk, cur.off, f);
        cur.next();
        if (!go) break;
    }
    cp_async_wait<0>();                             // nothing of this lane is in flight when the ring is reused or the CTA retires
}

// ------------------------------------------------------------------------------------------------
// guesses: phase before each chunk = init + sum of F_t (mod 1), from the per-run sums.  One warp per utterance.
// ------------------------------------------------------------------------------------------------
// WPU = warps per utterance: 1 (four utterances per CTA of 128 lanes: seconds of speech, hundreds of chunks) or 32 (one
// CTA of 1024 lanes per utterance: long forms, tens of thousands of chunks -- a lone warp would loop over them 32 at a
// time, a dependent global load per trip).  The CTA-wide scans are the warp scan, the warps' totals scanned by warp 0.
template <int WPU>
__global__ void __launch_bounds__(WPU == 1 ? 128 : 32 * WPU) k_phase_guess(PlanDev P)
{
    constexpr bool CTA = WPU > 1;
    constexpr uint32_t TILE = 32u * (uint32_t)WPU;
    __shared__ double sh_d[CTA ? 2 * WPU + 1 : 1];
    const uint32_t u = CTA ? blockIdx.x : (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, w = CTA ? (int)(threadIdx.x >> 5) : 0;
    if (u >= P.n_utts) return;
    const UttDev& U = P.utts[u];
    if (lane == 0 && w == 0) P.utt_status[u] = 0u;
    const uint32_t C = U.pc_count, PC = P.phase_chunk, K = P.pc_per_item, CL = P.chunk_len;
    if (C == 0) return;
    float* start = pcf(P, PCF_START) + U.pc_first;
    int32_t* flags = pci(P, PCF_FLAGS) + U.pc_first;
    const double* bs = P.bsum + (U.f_off >> 8);
    const uint32_t n = U.n_samples;
    auto chunk_sum = [&](uint32_t c) -> double {
        double s = 0.0;
        if (c < C) {
            const uint32_t ci = c / K, q = c - ci * K;
            const uint32_t s0 = ci * CL + q * PC, s1 = min(ci * CL + min(q * PC + PC, CL), n);   // multiples of 256 (but the end)
            const uint32_t r0 = s0 >> 8, r1 = (s1 + 255u) >> 8;
            for (uint32_t r = r0; r < r1; ++r) s = frac_d(s + bs[r]);
        }
        return s;
    };
    double carry = (double)U.init_phase;
    double s_next = chunk_sum((uint32_t)w * 32u + lane);
    for (uint32_t c0 = 0; c0 < C; c0 += TILE) {
        const uint32_t c = c0 + (uint32_t)w * 32u + lane;
        const double s = s_next;
        s_next = chunk_sum(c + TILE);              // the next block's loads fly during this block's scan
        double incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = frac_d(incl + t);
        }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0;
        double tot = __shfl_sync(0xffffffffu, incl, 31);
        if (CTA) {
            if (lane == 31) sh_d[w] = incl;
            __syncthreads();
            if (w == 0) {
                double a = lane < WPU ? sh_d[lane] : 0.0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double t = __shfl_up_sync(0xffffffffu, a, o);
                    if (lane >= o) a = frac_d(a + t);
                }
                double ae = __shfl_up_sync(0xffffffffu, a, 1);
                if (lane == 0) ae = 0.0;
                if (lane < WPU) sh_d[WPU + lane] = ae;
                if (lane == 31) sh_d[2 * WPU] = a;
            }
            __syncthreads();
            excl = frac_d(sh_d[WPU + w] + excl);
            tot = sh_d[2 * WPU];
        }
        if (c < C) {
            float gph = (c == 0) ? U.init_phase : (float)frac_d(carry + excl);
            if (c != 0 && !(gph < 1.0f)) gph = 0.0f;           // 0.99999999 rounds to 1.0f
            start[c] = gph;
            flags[c] = (int32_t)PCH_DIRTY;
        }
        carry = frac_d(carry + tot);
        if (CTA) __syncthreads();                  // sh_d is written again in the next trip
    }
}

// 8 literal steps; returns the number of wraps as a float sum.  (Blocks that reach past the utterance's end read the
// zeros k_frequency pads the row with: p + 0 is p, no wrap -- no special case.)
__device__ __forceinline__ float steps8(float& p, const float (&f)[8])
{
    float gs = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { float g; p = phase_step(p, f[k], g); gs += g; }
    return gs;
}
// the first wrapping step of a block walked from `p`: its index (8: none) and the phase after it (rare path)
__device__ __forceinline__ int first_wrap8(float p, const float (&f)[8], float* after)
{
    float g0, g1, g2, g3, g4, g5, g6, g7;
    const float a0 = phase_step(p, f[0], g0), a1 = phase_step(a0, f[1], g1), a2 = phase_step(a1, f[2], g2),
                a3 = phase_step(a2, f[3], g3), a4 = phase_step(a3, f[4], g4), a5 = phase_step(a4, f[5], g5),
                a6 = phase_step(a5, f[6], g6), a7 = phase_step(a6, f[7], g7);
    int k = 8;
    float r = 0.0f;
    if (g7 != 0.0f) { k = 7; r = a7; }
    if (g6 != 0.0f) { k = 6; r = a6; }
    if (g5 != 0.0f) { k = 5; r = a5; }
    if (g4 != 0.0f) { k = 4; r = a4; }
    if (g3 != 0.0f) { k = 3; r = a3; }
    if (g2 != 0.0f) { k = 2; r = a2; }
    if (g1 != 0.0f) { k = 1; r = a1; }
    if (g0 != 0.0f) { k = 0; r = a0; }
    *after = r;
    return k;
}

// ------------------------------------------------------------------------------------------------
// round A: one lane per chunk, one pass over [n0, first wrap of both trajectories inside the next chunk], the 32
// lanes of a warp in step over the same blocks of their 32 items
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, PH_OCC) k_phase_a(PlanDev P)
{
    __shared__ __align__(16) unsigned char ring[PH_RING_BYTES];
    const PChunkLane X = pchunk_of_lane(P);
    if (!X.valid) return;
    const bool last = X.c + 1 >= X.C;             // the last chunk hands no phase on, but the scan needs its anchor
    if (last && X.c == 0) return;
    const UttDev& U = P.utts[X.u];
    const uint32_t g = X.g, n1 = X.n1;
    float p0 = pcf(P, PCF_START)[g], p1 = p0;
    int32_t a = -1, a20 = -1, a21 = -1;
    float R = 0.0f, R20 = 0.0f, R21 = 0.0f, E0 = p0, E1 = p0;
    bool two = X.c == 0;                          // chunk 0 starts exact: one trajectory, no anchor needed
    // n1 is a multiple of 8 when the chunk has a successor (chunk ends are multiples of 256 but the utterance's)
    const uint32_t t_end = last ? ((n1 + 7u) & ~7u) : ((X.nn1 + 7u) & ~7u);
    walk_blocks(ring, P.F, U, P.chunk_len, X.n0, t_end, [&](uint32_t blk, size_t, const float (&f)[8]) -> bool {
        if (blk < n1) {
            if (two) {
                steps8(p0, f);
                steps8(p1, f);
            } else {
                // stage 1: from the guess towards the first wrap inside the chunk
                const float pb = p0;
                if (steps8(p0, f) != 0.0f) {
                    const int k = first_wrap8(pb, f, &R);
                    a = (int32_t)blk + k;
                    if (last) return false;
                    p0 = R;
                    p1 = sadd(R, PH_U23);         // exact: R is a small multiple of 2^-23
#pragma unroll
                    for (int i = 1; i < 8; ++i) {  // the rest of the anchor's block, both trajectories
                        if (i > k) {
                            float g0, g1;
                            p0 = phase_step(p0, f[i], g0);
                            p1 = phase_step(p1, f[i], g1);
                        }
                    }
                    two = true;
                }
            }
            return true;
        }
        // stage 3: on to the first wrap of each trajectory inside the next chunk
        if (blk == n1) {
            if (!two) p1 = p0;                    // no wrap in the whole chunk: one trajectory, the guess's own
            E0 = p0; E1 = p1;
        }
        const float b0 = p0, b1 = p1;
        const float w0 = steps8(p0, f), w1 = steps8(p1, f);
        if (w0 != 0.0f && a20 < 0) a20 = (int32_t)blk + first_wrap8(b0, f, &R20);
        if (w1 != 0.0f && a21 < 0) a21 = (int32_t)blk + first_wrap8(b1, f, &R21);
        return a20 < 0 || a21 < 0;
    });
    pci(P, PCF_A)[g] = a; pcf(P, PCF_R)[g] = R;
    if (last) return;
    pcf(P, PCF_E0)[g] = E0; pcf(P, PCF_E1)[g] = E1;
    pci(P, PCF_A20)[g] = a20; pci(P, PCF_A21)[g] = a21;
    pcf(P, PCF_R20)[g] = R20; pcf(P, PCF_R21)[g] = R21;
}

// ------------------------------------------------------------------------------------------------
// scan after round A.  Per chunk c the map from kc (how many 2^-23 units the truth is above the chunk's own
// trajectory after its anchor) to the same quantity of chunk c+1 is
//     kc even:  kc + D0           (trajectory 0, an even distance from the truth)
//     kc odd:   kc - 1 + D1       (trajectory 1)
// with Dj = (R2[j] - R_{c+1}) / 2^-23 when trajectory j wraps where chunk c+1's anchor is, and a constant when the
// chunk has no anchor of its own.  Maps of this "add by parity, or constant" form are closed under composition:
// an ordinary warp scan.  One warp per utterance.
// ------------------------------------------------------------------------------------------------
struct ParMap {
    int e, o;       // what an even / odd input becomes (relative: input + e/o; constant map: the value itself)
    int is_const;
};
__device__ __forceinline__ int parmap_apply(const ParMap& m, int k) { return m.is_const ? m.e : k + ((k & 1) ? m.o : m.e); }
// first f, then g
__device__ __forceinline__ ParMap parmap_compose(const ParMap& f, const ParMap& g)
{
    ParMap h;
    if (f.is_const) {
        h.is_const = 1;
        h.e = h.o = parmap_apply(g, f.e);
        return h;
    }
    if (g.is_const) return g;
    h.is_const = 0;
    // an even input becomes f.e (parity of f.e), then g adds by that parity
    h.e = f.e + ((f.e & 1) ? g.o : g.e);
    const int odd_out = 1 + f.o;                  // parity of (odd input + f.o)
    h.o = f.o + ((odd_out & 1) ? g.o : g.e);
    return h;
}
__device__ __forceinline__ ParMap parmap_shfl_up(const ParMap& m, int o)
{
    ParMap r;
    r.e = __shfl_up_sync(0xffffffffu, m.e, o);
    r.o = __shfl_up_sync(0xffffffffu, m.o, o);
    r.is_const = __shfl_up_sync(0xffffffffu, m.is_const, o);
    return r;
}
// exclusive scan of the block's maps applied to `kin`; returns this lane's input and updates kin to the block's output
__device__ __forceinline__ int parmap_block_scan(const ParMap& m, int lane, int& kin)
{
    ParMap incl = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const ParMap t = parmap_shfl_up(incl, o);
        if (lane >= o) incl = parmap_compose(t, incl);
    }
    ParMap excl = parmap_shfl_up(incl, 1);
    if (lane == 0) { excl.e = 0; excl.o = 0; excl.is_const = 0; }
    const int mine = parmap_apply(excl, kin);
    kin = __shfl_sync(0xffffffffu, parmap_apply(incl, kin), 31);
    return mine;
}

// the same over a CTA of WPU warps (sh: 2 * WPU + 1 maps of shared memory; every lane of the CTA must call it)
template <int WPU>
__device__ __forceinline__ int parmap_cta_scan(const ParMap& m, int lane, int w, int& kin, ParMap* sh)
{
    if (WPU == 1) return parmap_block_scan(m, lane, kin);
    ParMap incl = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const ParMap t = parmap_shfl_up(incl, o);
        if (lane >= o) incl = parmap_compose(t, incl);
    }
    ParMap excl = parmap_shfl_up(incl, 1);
    if (lane == 0) { excl.e = 0; excl.o = 0; excl.is_const = 0; }
    if (lane == 31) sh[w] = incl;
    __syncthreads();
    if (w == 0) {
        ParMap a;
        a.e = 0; a.o = 0; a.is_const = 0;
        if (lane < WPU) a = sh[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const ParMap t = parmap_shfl_up(a, o);
            if (lane >= o) a = parmap_compose(t, a);
        }
        ParMap ae = parmap_shfl_up(a, 1);
        if (lane == 0) { ae.e = 0; ae.o = 0; ae.is_const = 0; }
        if (lane < WPU) sh[WPU + lane] = ae;
        if (lane == 31) sh[2 * WPU] = a;
    }
    __syncthreads();
    const ParMap pre = sh[WPU + w], tot = sh[2 * WPU];
    const int mine = parmap_apply(excl, parmap_apply(pre, kin));
    kin = parmap_apply(tot, kin);
    __syncthreads();                               // sh is written again by the next call
    return mine;
}

template <int WPU>
__global__ void __launch_bounds__(WPU == 1 ? 128 : 32 * WPU) k_phase_scan_a(PlanDev P)
{
    constexpr bool CTA = WPU > 1;
    constexpr uint32_t TILE = 32u * (uint32_t)WPU;
    __shared__ ParMap sh_m[CTA ? 2 * WPU + 1 : 1];
    const uint32_t u = CTA ? blockIdx.x : (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, w = CTA ? (int)(threadIdx.x >> 5) : 0;
    if (u >= P.n_utts) return;
    const UttDev& U = P.utts[u];
    const uint32_t C = U.pc_count;
    if (C < 2) return;
    const uint32_t g0 = U.pc_first;
    const int32_t *A = pci(P, PCF_A) + g0, *A20 = pci(P, PCF_A20) + g0, *A21 = pci(P, PCF_A21) + g0;
    const float *Rr = pcf(P, PCF_R) + g0, *R20 = pcf(P, PCF_R20) + g0, *R21 = pcf(P, PCF_R21) + g0;
    const float *E0 = pcf(P, PCF_E0) + g0, *E1 = pcf(P, PCF_E1) + g0;
    float* start = pcf(P, PCF_START) + g0;
    struct In { int32_t a, an, a20, a21; float rn, r20, r21, e0, e1; };
    auto load = [&](uint32_t c) -> In {
        In x;
        x.a = x.an = x.a20 = x.a21 = -1; x.rn = x.r20 = x.r21 = x.e0 = x.e1 = 0.0f;
        if (c + 1 < C) {
            x.a = A[c]; x.an = A[c + 1]; x.a20 = A20[c]; x.a21 = A21[c];
            x.rn = Rr[c + 1]; x.r20 = R20[c]; x.r21 = R21[c]; x.e0 = E0[c]; x.e1 = E1[c];
        }
        return x;
    };
    int kin = 0;                                  // kc of the first chunk of this block of TILE (chunk 0: exact, 0)
    In nx = load((uint32_t)w * 32u + lane);
    for (uint32_t c0 = 0; c0 + 1 < C; c0 += TILE) {
        const uint32_t c = c0 + (uint32_t)w * 32u + lane;
        const bool valid = c + 1 < C;
        const In x = nx;
        nx = load(c + TILE);
        ParMap m;
        m.e = 0; m.o = 0; m.is_const = 0;         // identity for the padding lanes
        const bool anchored = (c == 0) || x.a >= 0;
        if (valid) {
            int D0 = 0, D1 = 0;
            // both operands are multiples of 2^-23 below 1: the difference is exact in f32
            if (x.a20 >= 0 && x.a20 == x.an) D0 = __float2int_rn((x.r20 - x.rn) * 8388608.0f);
            if (x.a21 >= 0 && x.a21 == x.an) D1 = __float2int_rn((x.r21 - x.rn) * 8388608.0f);
            if (c == 0 || !anchored) { m.is_const = 1; m.e = m.o = D0; }   // chunk 0 is exact (kc = 0, trajectory 0)
            else { m.e = D0; m.o = D1 - 1; }
        }
        const int kc = parmap_cta_scan<WPU>(m, lane, w, kin, sh_m);   // this chunk's own offset
        if (valid) {
            int j = 0, base = 0;
            if (c != 0 && anchored) { j = kc & 1; base = kc - j; }
            const float ej = j ? x.e1 : x.e0;
            float sf = (float)frac_d((double)ej + (double)base * (double)PH_U23);
            if (!(sf < 1.0f)) sf = ej;
            start[c + 1] = sf;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// round B: walk a chunk literally from its claimed start: the saw (tiled for k_formant) and the end phase.
// It also notes the first round-half-even TIE on a wrap step: there, and only there, a trajectory shifted by an odd
// number of 2^-23 units does not stay parallel (the tie goes to the even mantissa, and odd shifts swap which of
// the two candidates is even), it ends one unit further or nearer; k_phase_fix uses the note to carry a
// correction through this chunk without walking it twice.
// ------------------------------------------------------------------------------------------------
// the polyBLEP samples next to a wrap, src/lib.rs:503-517 (same operations as saw_edge, inlined)
__device__ __forceinline__ float saw_edge_inl(float phase, float f)
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
__device__ __forceinline__ uint32_t wrap_tie1(float p, float f)
{
    const float q = sadd(p, f);
    if (q >= 1.0f && p >= f) {
        // Fast2Sum: q + err is the exact sum; half an ulp of [1, 2) is a tie
        const float err = ssub(f, ssub(q, p));
        if (fabsf(err) == 5.9604644775390625e-08f) return err > 0.0f ? PCH_TIE_LOWER : PCH_TIE_UPPER;
    }
    return 0u;
}
// Blocks that hold a polyBLEP edge sample (the sample before and the sample after a carrier wrap: 2 in ~370 at
// 120 Hz).  First version (PH_INLINE_EDGES 0): such a block is only NOTED by the walk -- tiled offset and the phase at the
// block's start -- and redone afterwards, one block at a time, so that the divisions of the polyBLEP and the tie test
// cost a lane what its own wraps cost, instead of every lane of the warp paying for every other lane's wraps inside the
// hot loop.  Measured wrong: the kernel has issue slots to spare (38 % busy) but saturates the memory system, and every
// redo waited a full DRAM round trip for the block's eight increments under that load -- about 5 us each, three per
// chunk, a tenth of the kernel (ncu: 7 % of all stall samples on one instruction of the redo).  Keeping the increments
// in the note (local memory) was worse still.  Now the edge samples are fixed in place, while the block's increments
// and phases are in registers: 355 -> 326 us at config 2 (config-4 slice, phase total: 0.96 -> 0.87 ms), same bits.
constexpr int PH_EDGE_BUF = 24;
#ifndef PH_INLINE_EDGES
#define PH_INLINE_EDGES 1
#endif

__device__ __noinline__ uint32_t phase_b_redo_block(const float* __restrict__ F, float* __restrict__ saw, size_t off, float p,
                                                    uint32_t want_tie)
{
    float f[8];
    ldg256(F + off, f);
    float* dst = saw + off;
    uint32_t tie = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float pv = p;
        float g;
        p = phase_step(p, f[k], g);
        if (!((pv >= f[k]) && (pv <= ssub(1.0f, f[k])))) dst[k] = saw_edge_inl(pv, f[k]);
        if (want_tie && g != 0.0f && tie == 0u) tie = wrap_tie1(pv, f[k]);
    }
    return tie;
}

__global__ void __launch_bounds__(128, PH_OCC) k_phase_b(PlanDev P, uint32_t round)
{
    __shared__ __align__(16) unsigned char ring[PH_RING_BYTES];
    // Round 0 walks every chunk, a warp = one sub-range of one group of 32 items (coalesced).  The repair rounds walk the
    // chunks k_phase_fix listed as dirty, densely packed: a few per cent of all chunks, spread over almost every warp
    // of the round-0 mapping (lanes of a warp are 32 different utterances), so walking them in place would cost a
    // full round's instruction issue each time (measured 158 / 123 / 76 us per round at config 2 against 363 us for
    // round 0 itself).
    PChunkLane X;
    if (round == 0) {
        X = pchunk_of_lane(P);
        if (!X.valid) return;
    } else {
        const uint32_t n_dirty = P.pstats[PSTAT_PENDING + round];
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n_dirty) return;                                      // (nothing dirty: the whole grid leaves)
        X = pchunk_of_id(P, P.pdirty[(size_t)(round & 1u) * P.pc_stride + i]);
    }
    const uint32_t g = X.g;
    if (!((uint32_t)pci(P, PCF_FLAGS)[g] & PCH_DIRTY)) return;
    const UttDev& U = P.utts[X.u];
    const uint32_t CL = P.chunk_len, n0 = X.n0, n1 = X.n1;
    const float* F = P.F;
    float* dbg = P.phase_dbg ? P.phase_dbg + U.f_off : nullptr;
    float p = pcf(P, PCF_START)[g];
    uint32_t tie = 0u;
#if PH_INLINE_EDGES
    auto flush = []() {};
#else
    // note pad: the tiled offsets fit 32 bits in units of 8 floats (2^35 floats)
    uint32_t eb_off[PH_EDGE_BUF];
    float eb_p[PH_EDGE_BUF];
    int eb_n = 0;
    auto flush = [&]() {
#pragma unroll 1
        for (int i = 0; i < eb_n; ++i) {
            const uint32_t t = phase_b_redo_block(F, P.saw, (size_t)eb_off[i] << 3, eb_p[i], tie == 0u ? 1u : 0u);
            if (tie == 0u) tie = t;
        }
        eb_n = 0;
    };
#endif
    const uint32_t n1f = n1 & ~7u;                  // whole blocks; only an utterance's last chunk has a ragged tail
    walk_blocks(ring, F, U, CL, n0, n1f, [&](uint32_t blk, size_t off, const float (&f)[8]) -> bool {
        float pv[8], s[8];
        bool edge = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {                       // :520-525
            pv[k] = p;
            float gk;
            p = phase_step(p, f[k], gk);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s[k] = fmaf(2.0f, pv[k], -1.0f);                                    // :517 with polyblep = 0 (2p is exact)
            edge |= !((pv[k] >= f[k]) && (pv[k] <= ssub(1.0f, f[k]))) || !(p >= pv[k]);   // (second test: any wrap inside the block, or a NaN: its tie test)
        }
#if PH_INLINE_EDGES
        if (edge) {                                          // the polyBLEP samples and the tie test, in place
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (!((pv[k] >= f[k]) && (pv[k] <= ssub(1.0f, f[k])))) s[k] = saw_edge(pv[k], f[k]);
                if (tie == 0u) tie = wrap_tie1(pv[k], f[k]);           // (tests for the wrap itself)
            }
        }
        stg256(P.saw + off, s);
#else
        stg256(P.saw + off, s);
        if (edge) {                                          // noted; redone after the walk (or when the note pad is full)
            eb_off[eb_n] = (uint32_t)(off >> 3);
            eb_p[eb_n] = pv[0];
            if (++eb_n == PH_EDGE_BUF) flush();
        }
#endif
        if (dbg) {
#pragma unroll
            for (int k = 0; k < 8; ++k) dbg[blk + k] = pv[k];
        }
        return true;
    });
    flush();
    if (n1f < n1) {                                  // the ragged tail, sample by sample (the rest of its sector is zeroed)
        TileCursor tc;
        tc.seek(U, n1f, CL);
        const float* fsrc = F + tc.off;
        float* dst = P.saw + tc.off;
#pragma unroll 1
        for (uint32_t t = n1f; t < n1f + 8u; ++t) {
            float sv = 0.0f;
            if (t < n1) {
                const float f = fsrc[t - n1f], pv = p;
                float gk;
                p = phase_step(p, f, gk);
                sv = fmaf(2.0f, pv, -1.0f);
                if (!((pv >= f) && (pv <= ssub(1.0f, f)))) sv = saw_edge_inl(pv, f);
                if (gk != 0.0f && tie == 0u) tie = wrap_tie1(pv, f);
                if (dbg) dbg[t] = pv;
            }
            dst[t - n1f] = sv;
        }
    }
    pcf(P, PCF_END)[g] = p;
    pci(P, PCF_FLAGS)[g] = 0;
    pci(P, PCF_TIEKEY)[g] = tie == 0u ? -1 : (tie == PCH_TIE_LOWER ? 1 : 2);   // (only the kind matters to k_phase_fix)
    if (X.c + 1 == X.C) P.utt_final[(size_t)X.u * 32 + 24] = p;   // Synthesize.phase after the last sample (stream state)
    const unsigned act = __activemask();
    if ((threadIdx.x & 31) == (unsigned)(__ffs(act) - 1)) atomicAdd(P.pstats + PSTAT_WALKS, (uint32_t)__popc(act));
}

// ------------------------------------------------------------------------------------------------
// Repair rounds, in two kernels.  A repair round lasts as long as one chunk walk, and k_phase_b's walk is slow when it
// runs alone (few dirty chunks: one warp per scheduler, ~150 instructions per 8-sample block with the saw, the edge
// tests and the stores in the loop: 1 400 - 1 600 cycles per block measured, 190 us for a 2 048-sample chunk).  So a
// dirty chunk is re-walked as
//   k_phase_chain   the bare chain, one lane per dirty chunk: three operations per sample, the phase at the start of
//                   every 8-sample block parked in a scratch array -- the only serial part;
//   k_phase_saw     one lane per dirty BLOCK: replays its 8 steps from the parked phase, forms the saw with its
//                   polyBLEP edges, notes ties -- every block of every dirty chunk at once.
// Same operations per sample as k_phase_b, bit-identical results.  (A warp per dirty chunk, k_phase_pair style, was
// measured too: 196 / 102 / 26 us per round at config 2 -- fine for the last rounds, but the first one re-walks a
// tenth of all chunks and 31 of a warp's 32 lanes repeat the same chain.)
// ------------------------------------------------------------------------------------------------
struct DirtyRec {           // one per dirty chunk of the round in flight (written by k_phase_chain, read by k_phase_saw)
    unsigned long long off0;   // tiled offset of the chunk's first block
    unsigned long long dbg0;   // index of the chunk's first sample in the linear phase tap (f_off + n0)
    uint32_t g, n;             // chunk id, samples in the chunk
};

// One warp per CTA, a lane per dirty chunk.  The chain runs out of shared memory, 256 samples of every lane at a time:
// all 64 cp.async of a stage are issued at once (one DRAM round trip for 32 blocks) and WAITED FOR before the stepping
// starts.  (A ring that keeps copies in flight while the lane steps -- walk_blocks -- is right for round 0, where two
// dozen warps per SM hide each other's waits, but a lone warp gets nothing from it: ncu shows every shared-memory read
// parked behind the cp.async issued just before it, a full DRAM latency per 8-sample block, 1 000 cycles.)
constexpr uint32_t PCH_STAGE = 256;                   // samples per lane and stage
constexpr uint32_t PCH_ROW = PCH_STAGE + 4;           // floats per lane row: conflict-free 128-bit reads

__global__ void __launch_bounds__(32) k_phase_chain(PlanDev P, uint32_t round)
{
    __shared__ __align__(16) float sF[32 * PCH_ROW];
    const uint32_t n_dirty = P.pstats[PSTAT_PENDING + round];
    const uint32_t lane = threadIdx.x;
    const unsigned row_a = (unsigned)__cvta_generic_to_shared(sF) + lane * (PCH_ROW * 4u);
    const size_t stride = P.pc_stride;
    for (uint32_t base = blockIdx.x * 32u; base < n_dirty; base += gridDim.x * 32u) {
        const uint32_t i = base + lane;
        if (i >= n_dirty) continue;
        const PChunkLane X = pchunk_of_id(P, P.pdirty[(size_t)(round & 1u) * P.pc_stride + i]);
        const uint32_t g = X.g;
        DirtyRec rec;
        rec.g = g; rec.n = 0u; rec.off0 = 0ull; rec.dbg0 = 0ull;
        if (!((uint32_t)pci(P, PCF_FLAGS)[g] & PCH_DIRTY)) { P.pdrec[i] = rec; continue; }
        const UttDev& U = P.utts[X.u];
        const uint32_t n0 = X.n0, n1 = X.n1;
        TileCursor tc;                                                  // the chunk lies inside one work item:
        tc.seek(U, n0, P.chunk_len);                                    // its blocks are 256 floats apart in the tiled arrays
        const float* src = P.F + tc.off;
        float* park = P.ppark + i;                                      // block b of dirty chunk i at ppark[b * pc_stride + i]
        // whole blocks, the utterance's ragged last one included: k_frequency pads it with zeros, and p + 0 is p
        const uint32_t nblk = (n1 - n0 + 7u) >> 3;
        float p = pcf(P, PCF_START)[g];
        for (uint32_t b0 = 0; b0 < nblk; b0 += PCH_STAGE / 8u) {
            const uint32_t nb = min(PCH_STAGE / 8u, nblk - b0);
#pragma unroll 4
            for (uint32_t b = 0; b < nb; ++b) {
                cp_async16(row_a + b * 32u, src + (size_t)(b0 + b) * 256u);
                cp_async16(row_a + b * 32u + 16u, src + (size_t)(b0 + b) * 256u + 4u);
            }
            cp_async_commit();
            cp_async_wait<0>();
            float4 fa = lds128(row_a), fb = lds128(row_a + 16u);
#pragma unroll 1
            for (uint32_t b = 0; b < nb; ++b) {
                const float f[8] = { fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w };
                if (b + 1 < nb) { fa = lds128(row_a + (b + 1) * 32u); fb = lds128(row_a + (b + 1) * 32u + 16u); }
                park[(size_t)(b0 + b) * stride] = p;
                steps8(p, f);
            }
        }
        rec.off0 = tc.off;
        rec.n = n1 - n0;
        rec.dbg0 = U.f_off + n0;
        P.pdrec[i] = rec;
        pcf(P, PCF_END)[g] = p;
        pci(P, PCF_FLAGS)[g] = 0;
        pci(P, PCF_TIEKEY)[g] = -1;
        if (X.c + 1 == X.C) P.utt_final[(size_t)X.u * 32 + 24] = p;    // Synthesize.phase after the last sample (stream state)
        atomicAdd(P.pstats + PSTAT_WALKS, 1u);
    }
}

__global__ void __launch_bounds__(512) k_phase_saw(PlanDev P, uint32_t round)
{
    const uint32_t n_dirty = P.pstats[PSTAT_PENDING + round];
    if (n_dirty == 0u) return;
    // grid: y = block index inside the chunk, x strides over the dirty chunks (consecutive lanes: consecutive chunks, so the
    // parked phases and the records are read coalesced; no index arithmetic beyond an add)
    const uint32_t b = blockIdx.y;
    const size_t stride = P.pc_stride;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_dirty; i += gridDim.x * blockDim.x) {
        const DirtyRec rec = P.pdrec[i];
        if (8u * b >= rec.n) continue;
        const uint32_t valid = min(8u, rec.n - 8u * b);
        float q = P.ppark[(size_t)b * stride + i];
        float f[8], sv[8];
        const size_t off = (size_t)rec.off0 + (size_t)b * 256u;
        ldg256(P.F + off, f);
        uint32_t key = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            sv[k] = 0.0f;
            if ((uint32_t)k < valid) {
                const float pv = q;
                float gk;
                q = phase_step(q, f[k], gk);
                sv[k] = fmaf(2.0f, pv, -1.0f);                                       // :517 with polyblep = 0 (2p is exact)
                if (!((pv >= f[k]) && (pv <= ssub(1.0f, f[k])))) sv[k] = saw_edge_inl(pv, f[k]);
                if (gk != 0.0f && key == 0xFFFFFFFFu) {
                    const uint32_t t = wrap_tie1(pv, f[k]);
                    if (t) key = ((8u * b + (uint32_t)k) << 2) | (t == PCH_TIE_LOWER ? 1u : 2u);
                }
                if (P.phase_dbg) P.phase_dbg[(size_t)rec.dbg0 + 8u * b + k] = pv;
            }
        }
        stg256(P.saw + off, sv);
        if (key != 0xFFFFFFFFu) atomicMin(reinterpret_cast<unsigned int*>(pci(P, PCF_TIEKEY)) + rec.g, key);   // the chunk's FIRST tie
    }
}

// ------------------------------------------------------------------------------------------------
// the proof, and the repair of what fails it.  One warp per utterance.
//   phi_c = end_c - start_{c+1} must be 0 for every c.  Where it is not, the truth downstream is the walked
//   trajectory shifted by the accumulated difference, as long as that is a whole number of 2^-23 units
//   (translation invariance); an odd shift changes by one at the chunk's first wrap-step tie (see k_phase_b).  So the
//   shift d_c that chunk c's start needs obeys
//       d_{c+1} = phi_c + d_c                      d_c even, or no such tie in chunk c
//               = phi_c + d_c -+ 1                 d_c odd and the walk took the upper / lower candidate at the tie
//   ("add by parity": the same composable maps as k_phase_scan_a), and a phi that is off the 2^-23 lattice says
//   nothing about what follows: the chain restarts there with 0.  Chunks whose start changed are dirty and are
//   walked again by the k_phase_b that follows; a round that finds every phi = 0 has proven the utterance.
// ------------------------------------------------------------------------------------------------
template <int WPU>
__global__ void __launch_bounds__(WPU == 1 ? 128 : 32 * WPU) k_phase_fix(PlanDev P, uint32_t round /* the k_phase_b round that follows */)
{
    constexpr bool CTA = WPU > 1;
    constexpr uint32_t TILE = 32u * (uint32_t)WPU;
    __shared__ ParMap sh_m[CTA ? 2 * WPU + 1 : 1];
    const uint32_t u = CTA ? blockIdx.x : (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, w = CTA ? (int)(threadIdx.x >> 5) : 0;
    if (u >= P.n_utts) return;
    if (P.utt_status[u] & 1u) return;
    const UttDev& U = P.utts[u];
    const uint32_t C = U.pc_count;
    float* start = pcf(P, PCF_START) + U.pc_first;
    const float* endp = pcf(P, PCF_END) + U.pc_first;
    int32_t* flags = pci(P, PCF_FLAGS) + U.pc_first;
    bool any = false;
    uint32_t n_dirty = 0, n_bad = 0;   // (n_dirty: kept for debugging)
    int din = 0;                                  // d of this block's first chunk (chunk 0: its start is exact)
    const int32_t* tiekey = pci(P, PCF_TIEKEY) + U.pc_first;
    struct In { float e, s; uint32_t fl; };
    auto load = [&](uint32_t c) -> In {
        In x;
        x.e = x.s = 0.0f; x.fl = 0u;
        if (c + 1 < C) {
            x.e = endp[c]; x.s = start[c + 1];
            const int32_t k = tiekey[c];
            x.fl = k < 0 ? 0u : ((k & 3) == 1 ? PCH_TIE_LOWER : PCH_TIE_UPPER);
        }
        return x;
    };
    In nx = load((uint32_t)w * 32u + lane);
    for (uint32_t c0 = 0; c0 + 1 < C; c0 += TILE) {
        const uint32_t c = c0 + (uint32_t)w * 32u + lane;
        const bool valid = c + 1 < C;
        const In x = nx;
        nx = load(c + TILE);
        double phi = 0.0;
        bool bad = false, lat = true, hard = false;
        const float e = x.e, s = x.s;
        const int t = (x.fl & PCH_TIE_UPPER) ? -1 : ((x.fl & PCH_TIE_LOWER) ? 1 : 0);
        ParMap m;
        m.e = 0; m.o = 0; m.is_const = 0;
        if (valid) {
            bad = __float_as_uint(e) != __float_as_uint(s);
            int ki = 0;
            if (bad) {
                phi = centered_diff(e, s);
                if (!(fabs(phi) <= 0.5)) { hard = true; phi = 0.0; }      // NaN / inf: copy the proven value, no shift
                const double k = phi * 8388608.0;
                lat = !hard && (k == rint(k));
                ki = lat ? (int)rint(k) : 0;
            }
            if (lat) { m.e = ki; m.o = ki + t; }
            else { m.is_const = 1; m.e = m.o = 0; }
        }
        const unsigned badmask = __ballot_sync(0xffffffffu, bad);
        any |= badmask != 0u;
        n_bad += __popc(badmask);
        const int dc = parmap_cta_scan<WPU>(m, lane, w, din, sh_m);   // the shift chunk c's own start gets this round
        bool dirty = false;
        if (valid) {
            float ns = s;
            if (hard) ns = e;
            else {
                const int through = dc + ((dc & 1) ? t : 0);             // what is left of it after walking chunk c
                const double v = phi + (double)through * (double)PH_U23;
                if (v != 0.0) {
                    ns = (float)frac_d((double)s + v);
                    if (!(ns < 1.0f)) ns = 0.0f;
                }
            }
            dirty = __float_as_uint(ns) != __float_as_uint(s);
            if (dirty) {
                start[c + 1] = ns;
                flags[c + 1] = (int32_t)PCH_DIRTY;
            }
        }
        // the dirty chunks of this block go to the list the next k_phase_b walks (one atomic per block that has any)
        const unsigned dmask = __ballot_sync(0xffffffffu, dirty);
        if (dmask != 0u && round <= (uint32_t)PH_MAX_ROUNDS) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(P.pstats + PSTAT_PENDING + round, (uint32_t)__popc(dmask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (dirty) P.pdirty[(size_t)(round & 1u) * P.pc_stride + base + __popc(dmask & ((1u << lane) - 1u))] = U.pc_first + c + 1;
        }
        n_dirty += __popc(dmask);
    }
    if (CTA) {
        any = __syncthreads_or(any ? 1 : 0) != 0;
        if (lane == 0 && w != 0 && n_bad) atomicAdd(P.pstats + PSTAT_MISMATCH, n_bad);   // (warp 0 adds its own below)
    }
    if (lane == 0 && w == 0) {
        if (!any) {
            P.utt_status[u] = 1u;                  // proven: k_phase_pair leaves this utterance alone
        } else {
            atomicAdd(P.pstats + PSTAT_MISMATCH, n_bad);
            if (round <= (uint32_t)PH_MAX_ROUNDS) {
                atomicMax(P.pstats + PSTAT_ROUNDS, round);
            } else {
                atomicAdd(P.pstats + PSTAT_UNPROVEN, 1u);
            }
        }
    }
}

} // namespace grail
