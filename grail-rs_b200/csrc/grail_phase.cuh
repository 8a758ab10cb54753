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
// So an utterance is cut into chunks of PC samples, one lane each, and
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
       PCF_COUNT = 11 };
enum : uint32_t { PCH_DIRTY = 1u, PCH_TIE_LOWER = 2u, PCH_TIE_UPPER = 4u };
// pstats words
enum { PSTAT_WALKS = 0,       // chunks walked by k_phase_b, all rounds
       PSTAT_UNPROVEN = 1,    // utterances left to the serial chain
       PSTAT_ROUNDS = 2,      // last repair round that still found a mismatch (0: the first walk proved everything)
       PSTAT_MISMATCH = 3,    // chunk boundaries that failed the proof, all rounds
       PSTAT_PENDING = 16 };  // + r: dirty chunks waiting for round r's k_phase_b

constexpr int PH_MAX_ROUNDS = 14;
#ifndef PH_LOOKAHEAD
#define PH_LOOKAHEAD 2     // 32-sample rows of F_t per lane in shared memory (18 KB each per CTA of 128 lanes)
#endif
#ifndef PH_OCC
#define PH_OCC 6           // CTAs of 128 lanes per SM the walks are compiled for
#endif
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

// utterance that owns global phase chunk g
__device__ __forceinline__ uint32_t pchunk_utt(const UttDev* utts, uint32_t n_utts, uint32_t g)
{
    uint32_t lo = 0, hi = n_utts;    // last utterance with pc_first <= g (utterances without samples share their successor's pc_first)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (utts[mid].pc_first <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// guesses: phase before each chunk = init + sum of F_t (mod 1), from the per-run sums.  One warp per utterance.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_phase_guess(PlanDev P)
{
    const uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= P.n_utts) return;
    const UttDev& U = P.utts[u];
    if (lane == 0) P.utt_status[u] = 0u;
    const uint32_t C = U.pc_count, PC = P.phase_chunk;
    if (C == 0) return;
    float* start = pcf(P, PCF_START) + U.pc_first;
    int32_t* flags = pci(P, PCF_FLAGS) + U.pc_first;
    const double* bs = P.bsum + (U.f_off >> 8);
    const uint32_t runs = (U.n_samples + 255u) >> 8, rpc = PC >> 8;
    double carry = (double)U.init_phase;
    auto chunk_sum = [&](uint32_t c) -> double {
        double s = 0.0;
        if (c < C) {
            const uint32_t r0 = c * rpc, r1 = min(r0 + rpc, runs);
            for (uint32_t r = r0; r < r1; ++r) s = frac_d(s + bs[r]);
        }
        return s;
    };
    double s_next = chunk_sum(lane);
    for (uint32_t c0 = 0; c0 < C; c0 += 32) {
        const uint32_t c = c0 + lane;
        const double s = s_next;
        s_next = chunk_sum(c + 32);                // the next block's loads fly during this block's scan
        double incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = frac_d(incl + t);
        }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0;
        if (c < C) {
            float gph = (c == 0) ? U.init_phase : (float)frac_d(carry + excl);
            if (c != 0 && !(gph < 1.0f)) gph = 0.0f;           // 0.99999999 rounds to 1.0f
            start[c] = gph;
            flags[c] = (int32_t)PCH_DIRTY;
        }
        carry = frac_d(carry + __shfl_sync(0xffffffffu, incl, 31));
    }
}

// ------------------------------------------------------------------------------------------------
// F_t reader of the walks.  A lane streams its own chunk -- 32 lanes of a warp read 32 places 4-8 KB apart, the
// worst case for the L1 (one tag look-up per 32-byte sector: measured, the walks were bound by exactly that with
// LDG.256 and with 16-byte cp.async alike).  So the stream goes through the TMA instead: every lane moves 128-byte
// ROWS (32 samples) of its chunk into a private slice of shared memory with cp.async.bulk (UBLKCP), completion on a
// private mbarrier (complete_tx), PH_ROWS rows per lane, the next row in flight while the current one is stepped.
// Slice, barrier and data are touched by the issuing lane only: no warp- or CTA-level synchronisation, and lanes
// may diverge freely (round A's lanes leave their loops at different samples).  Rows are 144 bytes apart:
// conflict-free 128-bit reads.  fn(blk, f) is called for every 8-sample block of [b0, b1) in order and returns false
// to stop early.  Rows are copied whole; the F_t buffer is padded so that a row may reach past its utterance.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t PH_ROWS = PH_LOOKAHEAD;       // rows per lane (power of two)
constexpr uint32_t PH_LANES = 128;
constexpr uint32_t PH_ROW_STRIDE = 144;          // 128 bytes of samples + 16 of padding
constexpr uint32_t PH_RING_BYTES = PH_ROWS * PH_LANES * PH_ROW_STRIDE;

struct PhRing {
    unsigned data;        // shared-window address of this lane's row 0 (row r at + r * PH_LANES * PH_ROW_STRIDE)
    unsigned bar;         // ... of its mbarrier 0 (8 bytes each)
    uint32_t parity;      // bit r: parity of the phase row r's next completion will end
};
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned a, unsigned bytes)
{
    asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ PhRing ph_ring_init(unsigned char* smem, uint64_t* bars)
{
    PhRing R;
    R.data = (unsigned)__cvta_generic_to_shared(smem) + threadIdx.x * PH_ROW_STRIDE;
    R.bar = (unsigned)__cvta_generic_to_shared(bars + threadIdx.x * PH_ROWS);
    R.parity = 0u;
#pragma unroll
    for (uint32_t r = 0; r < PH_ROWS; ++r) mbar_init(bars + threadIdx.x * PH_ROWS + r, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the barriers are visible to the TMA
    return R;
}

template <class Fn>
__device__ __forceinline__ void stream_blocks(PhRing& R, const float* __restrict__ F, uint32_t b0, uint32_t b1, Fn&& fn)
{
    if (b0 >= b1) return;
    const uint32_t r0 = b0 & ~31u;
    uint32_t inflight = 0u;                        // bit r: row slot r has a copy in flight
    auto issue = [&](uint32_t row) {
        if (row < b1) {
            const uint32_t r = (row >> 5) & (PH_ROWS - 1u);
            mbar_arrive_expect_tx(R.bar + r * 8u, 128u);
            bulk_g2s(R.data + r * (PH_LANES * PH_ROW_STRIDE), F + row, 128u, R.bar + r * 8u);
            inflight |= 1u << r;
        }
    };
    auto land = [&](uint32_t r) {
        mbar_wait(R.bar + r * 8u, (R.parity >> r) & 1u);
        R.parity ^= 1u << r;
        inflight &= ~(1u << r);
    };
#pragma unroll 1
    for (uint32_t i = 0; i + 1u < PH_ROWS; ++i) issue(r0 + 32u * i);
    bool go = true;
#pragma unroll 1
    for (uint32_t row = r0; row < b1 && go; row += 32u) {
        issue(row + 32u * (PH_ROWS - 1u));
        const uint32_t r = (row >> 5) & (PH_ROWS - 1u);
        land(r);
        const unsigned src = R.data + r * (PH_LANES * PH_ROW_STRIDE);
#pragma unroll 1
        for (uint32_t i = 0; i < 4u && go; ++i) {
            const uint32_t blk = row + 8u * i;
            if (blk >= b0 && blk < b1) {
                const float4 fa = lds128(src + i * 32u), fb = lds128(src + i * 32u + 16u);
                const float f[8] = { fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w };
                go = fn(blk, f);
            }
        }
    }
    // rows still in flight (an early stop) must land before the slots and their barriers are used again
#pragma unroll
    for (uint32_t r = 0; r < PH_ROWS; ++r)
        if (inflight & (1u << r)) land(r);
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
// round A: one lane per chunk
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, PH_OCC) k_phase_a(PlanDev P)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.n_pchunks) return;
    const uint32_t u = pchunk_utt(P.utts, P.n_utts, g);
    const UttDev& U = P.utts[u];
    const uint32_t c = g - U.pc_first, C = U.pc_count;
    const bool last = c + 1 >= C;                 // the last chunk hands no phase on, but the scan needs its anchor
    if (last && c == 0) return;
    const uint32_t PC = P.phase_chunk, n = U.n_samples;
    const uint32_t n0 = c * PC, n1 = min(n0 + PC, n), nn1 = min(n1 + PC, n);
    const float* F = P.F + U.f_off;

    __shared__ __align__(16) unsigned char ring_mem[PH_RING_BYTES];
    __shared__ __align__(8) uint64_t ring_bar[PH_LANES * PH_ROWS];
    PhRing ring = ph_ring_init(ring_mem, ring_bar);
    float p0 = pcf(P, PCF_START)[g], p1 = p0;
    int32_t a = -1;
    float R = 0.0f;
    uint32_t t2 = n0;                             // first sample of the two-trajectory walk
    if (c != 0) {
        // stage 1: from the guess to the first wrap inside the chunk
        float p = p0;
        stream_blocks(ring, F, n0, (n1 + 7u) & ~7u, [&](uint32_t blk, const float (&f)[8]) -> bool {
            const float pb = p;
            if (steps8(p, f) != 0.0f) {
                a = (int32_t)blk + first_wrap8(pb, f, &R);
                return false;
            }
            return true;
        });
        if (last) { pci(P, PCF_A)[g] = a; pcf(P, PCF_R)[g] = R; return; }
        if (a >= 0) {
            t2 = (uint32_t)a + 1u;
            p0 = R;
            p1 = sadd(R, PH_U23);                 // exact: R is a small multiple of 2^-23
        } else {
            t2 = n1;                              // no wrap in the whole chunk: one trajectory, the guess's own
            p0 = p1 = p;
        }
    }
    // stage 2: both trajectories to the chunk's end (n1 is a multiple of 8 here: the chunk has a successor)
    if (t2 < n1) {
        stream_blocks(ring, F, t2 & ~7u, n1, [&](uint32_t blk, const float (&f)[8]) -> bool {
            if (blk >= t2) {
                steps8(p0, f);
                steps8(p1, f);
            } else {                              // the rest of the anchor's block
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (blk + k >= t2) {
                        float g0, g1;
                        p0 = phase_step(p0, f[k], g0);
                        p1 = phase_step(p1, f[k], g1);
                    }
                }
            }
            return true;
        });
    }
    const float E0 = p0, E1 = p1;
    // stage 3: on to the first wrap inside the next chunk
    int32_t a20 = -1, a21 = -1;
    float R20 = 0.0f, R21 = 0.0f;
    stream_blocks(ring, F, n1, (nn1 + 7u) & ~7u, [&](uint32_t blk, const float (&f)[8]) -> bool {
        const float b0 = p0, b1 = p1;
        const float w0 = steps8(p0, f), w1 = steps8(p1, f);
        if (w0 != 0.0f && a20 < 0) a20 = (int32_t)blk + first_wrap8(b0, f, &R20);
        if (w1 != 0.0f && a21 < 0) a21 = (int32_t)blk + first_wrap8(b1, f, &R21);
        return a20 < 0 || a21 < 0;
    });
    pci(P, PCF_A)[g] = a; pcf(P, PCF_R)[g] = R;
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

__global__ void __launch_bounds__(128) k_phase_scan_a(PlanDev P)
{
    const uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
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
    int kin = 0;                                  // kc of the first chunk of this block of 32 (chunk 0: exact, 0)
    In nx = load(lane);
    for (uint32_t c0 = 0; c0 + 1 < C; c0 += 32) {
        const uint32_t c = c0 + lane;
        const bool valid = c + 1 < C;
        const In x = nx;
        nx = load(c + 32);
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
        const int kc = parmap_block_scan(m, lane, kin);   // this chunk's own offset
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
// 120 Hz) are only NOTED by the walk -- block index and the phase at its start -- and redone afterwards, one block
// at a time: the divisions of the polyBLEP and the tie test then cost a lane what its own wraps cost, instead of
// every lane of the warp paying for every other lane's wraps inside the hot loop.
constexpr int PH_EDGE_BUF = 24;

struct EdgeCtx {
    const float* F;
    float* saw;
    float* dbg;
    uint32_t item_first, item_stride, CL;
};
__device__ __noinline__ uint32_t phase_b_redo_block(const EdgeCtx X, uint32_t blk, float p, uint32_t want_tie)
{
    float f[8];
    ldg256(X.F + blk, f);
    float* dst = X.saw + saw_index(X.item_first + (blk / X.CL) * X.item_stride, blk % X.CL, X.CL);
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
    if (round != 0 && P.pstats[PSTAT_PENDING + round] == 0u) return;   // nothing is dirty: the whole grid leaves
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.n_pchunks) return;
    if (!((uint32_t)pci(P, PCF_FLAGS)[g] & PCH_DIRTY)) return;
    const uint32_t u = pchunk_utt(P.utts, P.n_utts, g);
    const UttDev& U = P.utts[u];
    const uint32_t c = g - U.pc_first;
    const uint32_t PC = P.phase_chunk, n = U.n_samples, CL = P.chunk_len;
    const uint32_t n0 = c * PC, n1 = min(n0 + PC, n);
    const float* F = P.F + U.f_off;
    float* dbg = P.phase_dbg ? P.phase_dbg + U.f_off : nullptr;
    const uint32_t item_stride = U.item_stride;
    uint32_t dst_item = U.item_first + (n0 / CL) * item_stride, dst_j = n0 % CL;
    EdgeCtx X;
    X.F = F; X.saw = P.saw; X.dbg = dbg; X.item_first = U.item_first; X.item_stride = item_stride; X.CL = CL;
    float p = pcf(P, PCF_START)[g];
    uint32_t tie = 0u;
    uint32_t eb_blk[PH_EDGE_BUF];
    float eb_p[PH_EDGE_BUF];
    int eb_n = 0;
    auto flush = [&]() {
#pragma unroll 1
        for (int i = 0; i < eb_n; ++i) {
            const uint32_t t = phase_b_redo_block(X, eb_blk[i], eb_p[i], tie == 0u ? 1u : 0u);
            if (tie == 0u) tie = t;
        }
        eb_n = 0;
    };
    const uint32_t n1f = n1 & ~7u;                  // whole blocks; only an utterance's last chunk has a ragged tail
    __shared__ __align__(16) unsigned char ring_mem[PH_RING_BYTES];
    __shared__ __align__(8) uint64_t ring_bar[PH_LANES * PH_ROWS];
    PhRing ring = ph_ring_init(ring_mem, ring_bar);
    stream_blocks(ring, F, n0, n1f, [&](uint32_t blk, const float (&f)[8]) -> bool {
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
            edge |= !((pv[k] >= f[k]) && (pv[k] <= ssub(1.0f, f[k]))) || !(p >= pv[k]);
        }
        stg256(P.saw + saw_index(dst_item, dst_j, CL), s);
        if (edge) {                                          // noted; redone after the walk (or when the note pad is full)
            eb_blk[eb_n] = blk;
            eb_p[eb_n] = pv[0];
            if (++eb_n == PH_EDGE_BUF) flush();
        }
        if (dbg) {
#pragma unroll
            for (int k = 0; k < 8; ++k) dbg[blk + k] = pv[k];
        }
        dst_j += 8;
        if (dst_j >= CL) { dst_j = 0; dst_item += item_stride; }
        return true;
    });
    flush();
    if (n1f < n1) {                                  // the ragged tail, sample by sample (the rest of its sector is zeroed)
        float* dst = P.saw + saw_index(dst_item, dst_j, CL);
#pragma unroll 1
        for (uint32_t t = n1f; t < n1f + 8u; ++t) {
            float sv = 0.0f;
            if (t < n1) {
                const float f = F[t], pv = p;
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
    pci(P, PCF_FLAGS)[g] = (int32_t)tie;
    if (c + 1 == U.pc_count) P.utt_final[(size_t)u * 32 + 24] = p;   // Synthesize.phase after the last sample (stream state)
    const unsigned act = __activemask();
    if ((threadIdx.x & 31) == (unsigned)(__ffs(act) - 1)) atomicAdd(P.pstats + PSTAT_WALKS, (uint32_t)__popc(act));
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
__global__ void __launch_bounds__(128) k_phase_fix(PlanDev P, uint32_t round /* the k_phase_b round that follows */)
{
    const uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= P.n_utts) return;
    if (P.utt_status[u] & 1u) return;
    const UttDev& U = P.utts[u];
    const uint32_t C = U.pc_count;
    float* start = pcf(P, PCF_START) + U.pc_first;
    const float* endp = pcf(P, PCF_END) + U.pc_first;
    int32_t* flags = pci(P, PCF_FLAGS) + U.pc_first;
    bool any = false;
    uint32_t n_dirty = 0, n_bad = 0;
    int din = 0;                                  // d of this block's first chunk (chunk 0: its start is exact)
    struct In { float e, s; uint32_t fl; };
    auto load = [&](uint32_t c) -> In {
        In x;
        x.e = x.s = 0.0f; x.fl = 0u;
        if (c + 1 < C) { x.e = endp[c]; x.s = start[c + 1]; x.fl = (uint32_t)flags[c]; }
        return x;
    };
    In nx = load(lane);
    for (uint32_t c0 = 0; c0 + 1 < C; c0 += 32) {
        const uint32_t c = c0 + lane;
        const bool valid = c + 1 < C;
        const In x = nx;
        nx = load(c + 32);
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
        const int dc = parmap_block_scan(m, lane, din);   // the shift chunk c's own start gets this round
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
        n_dirty += __popc(__ballot_sync(0xffffffffu, dirty));
    }
    if (lane == 0) {
        if (!any) {
            P.utt_status[u] = 1u;                  // proven: k_phase_pair leaves this utterance alone
        } else {
            atomicAdd(P.pstats + PSTAT_MISMATCH, n_bad);
            if (round <= (uint32_t)PH_MAX_ROUNDS) {
                atomicAdd(P.pstats + PSTAT_PENDING + round, n_dirty);
                atomicMax(P.pstats + PSTAT_ROUNDS, round);
            } else {
                atomicAdd(P.pstats + PSTAT_UNPROVEN, 1u);
            }
        }
    }
}

} // namespace grail
