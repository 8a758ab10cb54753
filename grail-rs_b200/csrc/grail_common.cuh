// grail_common.cuh -- primitives shared by the host planner and the sm_100a kernels:
//   * strict (never contracted) f32 ops for the bit-exact set,
//   * the reference LCG with O(log n) jump-ahead,
//   * exact closed-form evaluation of the two constant-increment f32 clocks.
// Product code.  Nothing here comes from or calls oracle/.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../include/grail_cuda.h"

#if defined(__CUDACC__)
#define GRAIL_HD __host__ __device__ __forceinline__
#else
#define GRAIL_HD inline
#endif

namespace grail {

constexpr int NF = GRAIL_NUM_FORMANTS;

// ------------------------------------------------------------------------------------------------
// Strict f32.  The reference is Rust: every f32 operation rounds once, nothing is contracted.
// On the device the *_rn intrinsics are never fused by nvcc; on the host this file is compiled
// with -ffp-contract=off and each helper is a single operation anyway.
// ------------------------------------------------------------------------------------------------
GRAIL_HD float sadd(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
GRAIL_HD float ssub(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b;
    return r;
#endif
}
GRAIL_HD float smul(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
GRAIL_HD float sdiv(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b;
    return r;
#endif
}
GRAIL_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
GRAIL_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// ------------------------------------------------------------------------------------------------
// LCG of reference src/lib.rs:40: s <- s*16807 + 1 (mod 2^32), and its float map src/lib.rs:50-54.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t LCG_A = 16807u;
constexpr uint32_t LCG_C = 1u;

GRAIL_HD uint32_t lcg_step(uint32_t s) { return s * LCG_A + LCG_C; }

// ((s >> 9) | 0x3F800000) reinterpreted, minus 1.5, times 2: both float ops are exact.
GRAIL_HD float lcg_float(uint32_t s)
{
    float f = u2f((s >> 9) | 0x3F800000u);
    return smul(ssub(f, 1.5f), 2.0f);
}

// state after n more draws: the affine map composed by binary doubling, (A,C)^2 = (A*A, A*C + C).
GRAIL_HD uint32_t lcg_jump(uint32_t s, uint64_t n)
{
    uint32_t a = LCG_A, c = LCG_C;
    while (n) {
        if (n & 1u) s = a * s + c;
        c = a * c + c;
        a = a * a;
        n >>= 1;
    }
    return s;
}
// LCG^8 as one affine step (used once per jitter wrap: the 8 formant draws are 8 apart)
GRAIL_HD void lcg_pow(uint32_t n, uint32_t* a_out, uint32_t* c_out)
{
    uint32_t ra = 1u, rc = 0u, a = LCG_A, c = LCG_C;
    while (n) {
        if (n & 1u) { rc = a * rc + c; ra = a * ra; }
        c = a * c + c;
        a = a * a;
        n >>= 1;
    }
    *a_out = ra;
    *c_out = rc;
}

// Draw indices (1-based, of the LCG started at the .jitter() seed) that the three value-noise
// generators hold after w wraps.  Derived from src/lib.rs:227-237, 270-286, 301, 786-791: the
// generators are built from one running seed, so their private states start after 2, 18 and 34 draws
// and they read overlapping windows of one stream.
GRAIL_HD uint64_t jit_freq_cur_idx(uint64_t w) { return 1 + w; }
GRAIL_HD uint64_t jit_freq_next_idx(uint64_t w) { return 2 + w; }
GRAIL_HD uint64_t jit_arr_cur_idx(int gen /*0: formant_freq, 1: formant_amp*/, int i, uint64_t w)
{
    const uint64_t base0 = gen ? 18 : 2, base1 = gen ? 34 : 18;
    if (w == 0) return base0 + 2 * (uint64_t)i + 1;
    if (w == 1) return base0 + 2 * (uint64_t)i + 2;
    return base1 + 8 * (w - 2) + (uint64_t)i + 1;
}
GRAIL_HD uint64_t jit_arr_next_idx(int gen, int i, uint64_t w)
{
    const uint64_t base0 = gen ? 18 : 2, base1 = gen ? 34 : 18;
    if (w == 0) return base0 + 2 * (uint64_t)i + 2;
    return base1 + 8 * (w - 1) + (uint64_t)i + 1;
}

// ------------------------------------------------------------------------------------------------
// Exact clocks.
//
// x <- RN(x -+ d) with constant d is, while x stays inside one binade [2^e, 2^(e+1)) with grid
// u = 2^(e-23), the integer update m <- m -+ S on the mantissa m = x/u, where S is d/u rounded to
// nearest (ties handled through the parity of m).  So a whole binade is skipped in O(1) and only the
// steps that change binade are executed as real f32 operations.  Bit-exact by construction; tested
// against literal loops in tests/test_clocks.py.
// ------------------------------------------------------------------------------------------------
struct ClockRun {
    float    x;      // value after the last step taken
    uint64_t steps;  // steps taken
    int      stuck;  // 1 if the clock can never move again (x -+ d rounds back to x)
};

// integer step S for mantissa m at exponent field ex (biased), given d's mantissa md and exponent
// field ed: d/u rounded to nearest, or 0 when no closed form applies here (d >= the whole binade, d < u/2,
// or a round-half-even tie on an odd mantissa) -- the caller then takes one real f32 step.
GRAIL_HD uint32_t clock_binade_step(uint32_t m, int ex, uint32_t md, int ed)
{
    const int sh = ex - ed; // d/u = md * 2^-sh
    if (sh < 0 || sh >= 26) return 0;
    if (sh == 0) return md;
    const uint32_t s = md >> sh;
    const uint32_t rem = md & ((1u << sh) - 1u);
    const uint32_t half = 1u << (sh - 1);
    if (rem < half) return s;
    if (rem > half) return s + 1;
    // exact tie: with m even the step is the even one of {s, s+1}, and m stays even afterwards
    if (m & 1u) return 0;
    return s + (s & 1u);
}

// Descending clock (Sequencer `time -= delta_time`, reference src/lib.rs:861).
// Takes steps x <- RN(x - d) until `max_steps` are done or x < 0 (that step is counted).
GRAIL_HD ClockRun clock_desc_run(float x, float d, uint64_t max_steps)
{
    ClockRun r;
    r.steps = 0;
    r.stuck = 0;
    const uint32_t db = f2u(d);
    const int ed = (int)((db >> 23) & 0xFF);
    const uint32_t md = (db & 0x7FFFFFu) | 0x800000u;
    const bool d_ok = (ed > 0 && ed < 255 && !(db >> 31));
    while (r.steps < max_steps) {
        const uint32_t xb = f2u(x);
        const int ex = (int)((xb >> 23) & 0xFF);
        if (d_ok && !(xb >> 31) && ex > 0 && ex < 255) {
            const uint32_t m = (xb & 0x7FFFFFu) | 0x800000u;
            const uint32_t S = clock_binade_step(m, ex, md, ed);
            if (S != 0 && m > 0x800001u) {
                // every in-binade step needs m_prev - S >= 2^23 + 1 (then the exact difference is
                // still inside this binade, so the rounding grid is u)
                uint64_t k = (m - 0x800001u) / S;   // 32-bit divide
                if (k > max_steps - r.steps) k = max_steps - r.steps;
                if (k > 0) {
                    const uint32_t m2 = m - (uint32_t)(k * S);
                    x = u2f(((uint32_t)ex << 23) | (m2 & 0x7FFFFFu));
                    r.steps += k;
                    continue;
                }
            }
        }
        const float y = ssub(x, d);
        if (y == x) {
            r.stuck = 1;
            break;
        }
        x = y;
        r.steps++;
        if (x < 0.0f) break;
    }
    r.x = x;
    return r;
}

// Ascending clock (value-noise `phase += increment`, reference src/lib.rs:242,291).
// Takes steps x <- RN(x + d) until `max_steps` are done or x > 1.0 (that step is counted; the
// caller applies the `phase -= 1.0` wrap of src/lib.rs:245-246).
GRAIL_HD ClockRun clock_asc_run(float x, float d, uint64_t max_steps)
{
    ClockRun r;
    r.steps = 0;
    r.stuck = 0;
    const uint32_t db = f2u(d);
    const int ed = (int)((db >> 23) & 0xFF);
    const uint32_t md = (db & 0x7FFFFFu) | 0x800000u;
    const bool d_ok = (ed > 0 && ed < 255 && !(db >> 31));
    while (r.steps < max_steps) {
        const uint32_t xb = f2u(x);
        const int ex = (int)((xb >> 23) & 0xFF);
        if (d_ok && !(xb >> 31) && ex > 0 && ex < 127) { // normal, 0 < x < 1
            const uint32_t m = (xb & 0x7FFFFFu) | 0x800000u;
            const uint32_t S = clock_binade_step(m, ex, md, ed);
            if (S != 0 && m < 0xFFFFFFu) {
                // m_prev + S <= 2^24 - 1 keeps the exact sum below 2^(e+1): grid u, no wrap (x < 1)
                uint64_t k = (0xFFFFFFu - m) / S;   // 32-bit divide
                if (k > max_steps - r.steps) k = max_steps - r.steps;
                if (k > 0) {
                    const uint32_t m2 = m + (uint32_t)(k * S);
                    x = u2f(((uint32_t)ex << 23) | (m2 & 0x7FFFFFu));
                    r.steps += k;
                    continue;
                }
            }
        }
        const float y = sadd(x, d);
        if (y == x) {
            r.stuck = 1;
            break;
        }
        x = y;
        r.steps++;
        if (x > 1.0f) break;
    }
    r.x = x;
    return r;
}

// ------------------------------------------------------------------------------------------------
// Per-utterance / per-phoneme schedule records produced by the host planner.
// ------------------------------------------------------------------------------------------------
struct SegRec {          // one per phoneme (Sequencer "cur" element)
    uint32_t start;      // index (within the utterance) of the phoneme's first sample
    float    time0;      // Sequencer.time at that sample, after the hand-over add (src/lib.rs:873,882)
};

struct JitRec {          // one per jitter period (value-noise wrap), shared by all 3 generators
    int32_t  n;          // sample at which this period's wrap happened (-1 for the initial period)
    float    phase;      // value-noise phase after that sample (0 for the initial period)
};


// Wrap schedule of one value-noise phase clock over samples [0, n_max): rec[w] = {sample of the w-th wrap,
// phase after it}; rec[0] is the initial period.  Returns the number of records, or 0 if `cap` is too small.
// Shared by the host planner (few schedules) and k_jitter_schedule (many).
GRAIL_HD uint32_t jitter_schedule_walk(float inc, uint32_t n_max, JitRec* rec, uint32_t cap, float phase0 = 0.0f)
{
    if (cap == 0) return 0;
    int64_t n = -1;
    float ph = phase0;   // value-noise phase before sample 0 (0 for a fresh Jitter, carried for a continued stream)
    uint32_t w = 0;
    rec[0].n = -1;
    rec[0].phase = phase0;
    const int64_t last = (int64_t)n_max - 1;
    while (n < last) {
        const ClockRun r = clock_asc_run(ph, inc, (uint64_t)(last - n));
        n += (int64_t)r.steps;
        if (r.stuck || !(r.x > 1.0f)) break;
        ph = ssub(r.x, 1.0f); // src/lib.rs:246
        ++w;
        if (w >= cap) return 0;
        rec[w].n = (int32_t)n;
        rec[w].phase = ph;
    }
    return w + 1;
}

} // namespace grail
