/*
 * grail_cuda_debug.h -- test hooks of libgrail_cuda.so (host-only unless they take a grail_ctx).
 * They expose the closed-form exact-clock and LCG jump-ahead primitives the planner and the kernels
 * share (grail-rs_b200/csrc/grail_common.cuh) so tests can compare them with literal f32 loops.
 */
#ifndef GRAIL_CUDA_DEBUG_H
#define GRAIL_CUDA_DEBUG_H
#include <stdint.h>
#include "grail_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif
/* x <- RN(x - d) until max_steps or x < 0 (reference src/lib.rs:861) */
void grail_cuda_debug_clock_desc(float x, float d, uint64_t max_steps, float* x_out, uint64_t* steps, int* stuck);
/* x <- RN(x + d) until max_steps or x > 1 (reference src/lib.rs:242-245) */
void grail_cuda_debug_clock_asc(float x, float d, uint64_t max_steps, float* x_out, uint64_t* steps, int* stuck);
/* LCG state after n more draws (reference src/lib.rs:40) */
uint32_t grail_cuda_debug_lcg_jump(uint32_t seed, uint64_t n);
/* 1-based draw index held by a jitter generator after w wraps: gen -1 = freq_noise, 0 = formant_freq_noise,
 * 1 = formant_amp_noise; which 0 = current, 1 = next (reference src/lib.rs:227-301, 786-791) */
uint64_t grail_cuda_debug_jitter_index(int gen, int which, int i, uint64_t w);
/* device check of k_frequency's division by a per-segment constant (correctly rounded reciprocal + one FMA
 * correction) against the IEEE division, on n_pairs pseudo-random (a, b) pairs: *mismatches must come back 0 */
int grail_cuda_debug_div_check(grail_ctx* ctx, uint32_t seed, uint64_t n_pairs, uint64_t* mismatches);
#ifdef __cplusplus
}
#endif
#endif
