/*
 * grail_oracle.c -- CPU ORACLE for the grail-rs waveform-generation hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libgrail_cuda.so) never links, loads or calls anything in this directory.
 *
 * What it is: a plain-C restatement of the reference's per-sample iterator chain
 *     Sequencer -> Jitter -> Synthesize            (reference src/lib.rs)
 * in the reference's own operation order, in strict IEEE-754 binary32
 * (build with -ffp-contract=off -fno-fast-math; x86-64 SSE arithmetic is strict,
 * which is what Rust guarantees for f32).  Every function cites the reference lines
 * it follows.
 *
 * PARITY UNPINNED: the reference ships no golden audio, no known-answer test and no
 * fixture for this path (its only asserting tests pin the text Transcriber,
 * src/lib.rs:1210-1358; synthesize_normalized / synthesize_resampled /
 * jitter_within_bounds are empty, src/lib.rs:603-608,804-805) and it cannot be
 * compiled here (no Rust toolchain).  The authority of this oracle is line-by-line
 * fidelity plus agreement with the independent survey probe's known answers
 * (SURVEY.md Appendix B), which tests/test_oracle_kat.py checks.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NF 8 /* NUM_FORMANTS, src/lib.rs:24 */

/* ------------------------------------------------------------------------------------------
 * Data crossing the Selector -> Sequencer cut.  Same memory layout as include/grail_cuda.h so
 * one numpy buffer can feed both sides, but declared independently here.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float v[NF]; } arr_t; /* Array, src/lib.rs:88 */

typedef struct {           /* SynthesisElem, field order of src/lib.rs:316-337 */
    float frequency;
    arr_t formant_freq, formant_bw, formant_smooth, formant_breath, formant_turb, formant_amp;
} elem_t;

typedef struct {           /* SequenceElem, src/lib.rs:814-824 (Option flattened to a flag) */
    uint32_t has_elem;
    elem_t   elem;
    float    length, blend_length;
} seq_elem_t;

typedef struct {           /* the Voice scalars the hot path reads, src/lib.rs:696-717 */
    float    sample_rate, jitter_frequency, jitter_delta_frequency,
             jitter_delta_formant_frequency, jitter_delta_amplitude;
    uint32_t jitter_seed;  /* argument of .jitter(seed, voice), src/lib.rs:786 */
    uint32_t synth_seed;   /* always 0 in the reference, src/lib.rs:594 */
} voice_params_t;

/* optional per-sample intermediates, any pointer may be NULL */
typedef struct {
    float    *time, *alpha, *jitter_phase, *frequency, *carrier_phase;
    uint32_t *phoneme_index;
    float    *state_a, *state_b, *state_c; /* NF floats per sample, after the sample */
} trace_t;

/* ------------------------------------------------------------------------------------------
 * L0 math kernel
 * ---------------------------------------------------------------------------------------- */

/* src/lib.rs:36-55 */
float grail_oracle_random_f32(uint32_t *state)
{
    *state = *state * 16807u + 1u;                 /* wrapping_mul / wrapping_add */
    uint32_t res = (*state >> 9) | 0x3F800000u;
    float f;
    memcpy(&f, &res, 4);
    return (f - 1.5f) * 2.0f;
}

/* src/lib.rs:63-70 */
float grail_oracle_tan_approx(float x)
{
    return ((1.0f - x) * x * (5.0f - 4.0f * (x + 0.5f) * (0.5f - x))) /
           ((x + 0.5f) * (5.0f - 4.0f * (1.0f - x) * x) * (0.5f - x));
}

/* src/lib.rs:75-82 */
float grail_oracle_exp_approx(float x)
{
    float o = 1.0f - x;
    float o2 = o * o;
    return o2 * o2 * o;
}

/* Array helpers, src/lib.rs:88-211.  Elementwise, one rounding per operation. */
static arr_t a_splat(float x) { arr_t r; for (int i = 0; i < NF; i++) r.v[i] = x; return r; }
static arr_t a_add(arr_t a, arr_t b) { arr_t r; for (int i = 0; i < NF; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
static arr_t a_sub(arr_t a, arr_t b) { arr_t r; for (int i = 0; i < NF; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
static arr_t a_mul(arr_t a, arr_t b) { arr_t r; for (int i = 0; i < NF; i++) r.v[i] = a.v[i] * b.v[i]; return r; }
static arr_t a_div(arr_t a, arr_t b) { arr_t r; for (int i = 0; i < NF; i++) r.v[i] = a.v[i] / b.v[i]; return r; }
/* src/lib.rs:123-125: iter().sum::<f32>() is a sequential left fold */
static float a_sum(arr_t a) { float s = 0.0f; for (int i = 0; i < NF; i++) s = s + a.v[i]; return s; }
/* src/lib.rs:135-137: 1.0 - alpha is recomputed per element (same value each time) */
static arr_t a_blend(arr_t a, arr_t b, float alpha)
{
    arr_t r;
    for (int i = 0; i < NF; i++) r.v[i] = a.v[i] * (1.0f - alpha) + b.v[i] * alpha;
    return r;
}
/* src/lib.rs:141-143 */
static arr_t a_blend_multiple(arr_t a, arr_t b, arr_t alpha)
{
    return a_add(a_mul(a, a_sub(a_splat(1.0f), alpha)), a_mul(b, alpha));
}

/* ------------------------------------------------------------------------------------------
 * SynthesisElem helpers, src/lib.rs:341-460
 * ---------------------------------------------------------------------------------------- */

/* src/lib.rs:367-377 */
static elem_t elem_silent(void)
{
    elem_t e;
    e.frequency = 0.25f;
    e.formant_freq = a_splat(0.25f);
    e.formant_bw = a_splat(0.25f);
    e.formant_smooth = a_splat(0.25f);
    e.formant_breath = a_splat(0.0f);
    e.formant_turb = a_splat(0.0f);
    e.formant_amp = a_splat(0.0f);
    return e;
}

/* src/lib.rs:404-414 */
static elem_t elem_blend(elem_t s, elem_t o, float alpha)
{
    elem_t r;
    r.frequency = s.frequency * (1.0f - alpha) + o.frequency * alpha;
    r.formant_freq = a_blend(s.formant_freq, o.formant_freq, alpha);
    r.formant_smooth = a_blend(s.formant_smooth, o.formant_smooth, alpha);
    r.formant_bw = a_blend(s.formant_bw, o.formant_bw, alpha);
    r.formant_turb = a_blend(s.formant_turb, o.formant_turb, alpha);
    r.formant_breath = a_blend(s.formant_breath, o.formant_breath, alpha);
    r.formant_amp = a_blend(s.formant_amp, o.formant_amp, alpha);
    return r;
}

/* src/lib.rs:454-459 */
static elem_t elem_copy_silent(elem_t s) { s.formant_amp = a_splat(0.0f); return s; }

/* src/lib.rs:418-440 */
void grail_oracle_elem_resample(const elem_t *in, float old_rate, float new_rate, elem_t *out)
{
    float scale = old_rate / new_rate;
    arr_t ff = a_mul(in->formant_freq, a_splat(scale));
    elem_t r = *in;
    r.frequency = fminf(in->frequency * scale, 0.5f);
    for (int i = 0; i < NF; i++) r.formant_freq.v[i] = fminf(ff.v[i], 0.5f);
    r.formant_bw = a_mul(in->formant_bw, a_splat(scale));
    r.formant_smooth = a_mul(in->formant_smooth, a_splat(scale));
    for (int i = 0; i < NF; i++) r.formant_amp.v[i] = (ff.v[i] > 0.5f) ? 0.0f : in->formant_amp.v[i];
    *out = r;
}

/* src/lib.rs:381-401; argument order freq, bw, smooth, turb, breath, amp */
void grail_oracle_new_phoneme(const float *freq, const float *bw, const float *smooth,
                              const float *turb, const float *breath, const float *amp, elem_t *out)
{
    elem_t e;
    arr_t a;
    e.frequency = 0.0f;
    memcpy(e.formant_freq.v, freq, sizeof(arr_t));
    memcpy(e.formant_bw.v, bw, sizeof(arr_t));
    memcpy(e.formant_smooth.v, smooth, sizeof(arr_t));
    memcpy(e.formant_breath.v, breath, sizeof(arr_t));
    memcpy(e.formant_turb.v, turb, sizeof(arr_t));
    memcpy(a.v, amp, sizeof(arr_t));
    e.formant_amp = a_div(a, a_splat(a_sum(a)));   /* :398 */
    grail_oracle_elem_resample(&e, 1.0f, 44100.0f, out); /* :400, DEFAULT_SAMPLE_RATE :21 */
}

/* src/lib.rs:445-450 */
void grail_oracle_copy_with_frequency(const elem_t *in, float frequency, elem_t *out)
{
    *out = *in;
    out->frequency = fminf(frequency, 0.5f);
}

/* ------------------------------------------------------------------------------------------
 * Value noise, src/lib.rs:218-307
 * ---------------------------------------------------------------------------------------- */
typedef struct { float current, next, phase; uint32_t state; } value_noise_t;
typedef struct { arr_t current, next; float phase; uint32_t state; } array_value_noise_t;

/* src/lib.rs:227-237 */
static value_noise_t vn_new(uint32_t *state)
{
    value_noise_t n;
    n.current = grail_oracle_random_f32(state);
    n.next = grail_oracle_random_f32(state);
    n.phase = 0.0f;
    n.state = *state;          /* private copy of the shared seed */
    return n;
}

/* src/lib.rs:240-255 */
static float vn_next(value_noise_t *n, float increment)
{
    n->phase += increment;
    if (n->phase > 1.0f) {
        n->phase -= 1.0f;
        n->current = n->next;
        n->next = grail_oracle_random_f32(&n->state);
    }
    return n->current * (1.0f - n->phase) + n->next * n->phase;
}

/* src/lib.rs:270-286: current[i], next[i] drawn interleaved */
static array_value_noise_t avn_new(uint32_t *state)
{
    array_value_noise_t n;
    for (int i = 0; i < NF; i++) {
        n.current.v[i] = grail_oracle_random_f32(state);
        n.next.v[i] = grail_oracle_random_f32(state);
    }
    n.phase = 0.0f;
    n.state = *state;
    return n;
}

/* src/lib.rs:289-306 */
static arr_t avn_next(array_value_noise_t *n, float increment)
{
    n->phase += increment;
    if (n->phase > 1.0f) {
        n->phase -= 1.0f;
        n->current = n->next;
        for (int i = 0; i < NF; i++) n->next.v[i] = grail_oracle_random_f32(&n->state);
    }
    return a_add(a_mul(n->current, a_splat(1.0f - n->phase)), a_mul(n->next, a_splat(n->phase)));
}

/* ------------------------------------------------------------------------------------------
 * Sequencer, src/lib.rs:839-953.  The upstream iterator is a cursor over an array.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const seq_elem_t *src;
    uint32_t n_src, pos;          /* upstream iterator */
    int has_cur, has_next;
    seq_elem_t cur, next;
    float time, delta_time;
    uint32_t cur_index;           /* index of cur in src (trace only) */
} sequencer_t;

static int upstream_next(sequencer_t *s, seq_elem_t *out)
{
    if (s->pos >= s->n_src) return 0;
    *out = s->src[s->pos++];
    return 1;
}

/* src/lib.rs:941-949 */
static void sequencer_new(sequencer_t *s, const seq_elem_t *src, uint32_t n, float sample_rate)
{
    memset(s, 0, sizeof *s);
    s->src = src;
    s->n_src = n;
    s->delta_time = 1.0f / sample_rate;
    s->time = 0.0f;
}

/* src/lib.rs:859-932; returns 0 for None */
static int sequencer_next(sequencer_t *s, elem_t *out, float *alpha_out)
{
    s->time -= s->delta_time;                                  /* :861 */
    if (s->time < 0.0f) {                                      /* :864 */
        if (s->has_cur && s->has_next) {                       /* :868-874 */
            float len = s->next.length;
            s->cur = s->next;
            s->cur_index = s->pos - 1;
            s->has_next = upstream_next(s, &s->next);
            s->time += len;
        } else if (!s->has_cur && !s->has_next) {              /* :876-884 */
            s->has_cur = upstream_next(s, &s->cur);
            s->cur_index = 0;
            s->has_next = upstream_next(s, &s->next);
            if (s->has_cur) s->time += s->cur.length;
        } else {
            return 0;                                          /* :886 */
        }
    }
    if (!s->has_cur) return 0;                                 /* :930 */
    int b_on = s->cur.has_elem;
    int c_on = s->has_next && s->next.has_elem;
    float alpha = 1.0f;
    if (b_on || c_on) alpha = fminf(s->time / s->cur.blend_length, 1.0f); /* :899,908,917 */
    if (alpha_out) *alpha_out = alpha;
    if (b_on && c_on) {
        *out = elem_blend(s->next.elem, s->cur.elem, alpha);   /* :902 */
    } else if (b_on) {
        *out = elem_blend(elem_copy_silent(s->cur.elem), s->cur.elem, alpha);   /* :911 */
    } else if (c_on) {
        *out = elem_blend(s->next.elem, elem_copy_silent(s->next.elem), alpha); /* :920 */
    } else {
        *out = elem_silent();                                  /* :926 */
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Jitter, src/lib.rs:724-801
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    value_noise_t freq_noise;
    array_value_noise_t formant_freq_noise, formant_amp_noise;
    float frequency, delta_frequency, delta_formant_freq, delta_amplitude;
} jitter_t;

/* src/lib.rs:786-797: the three generators are seeded from one running `mut seed` */
static void jitter_new(jitter_t *j, const voice_params_t *v)
{
    uint32_t seed = v->jitter_seed;
    j->freq_noise = vn_new(&seed);
    j->formant_freq_noise = avn_new(&seed);
    j->formant_amp_noise = avn_new(&seed);
    j->frequency = v->jitter_frequency;
    j->delta_frequency = v->jitter_delta_frequency;
    j->delta_formant_freq = v->jitter_delta_formant_frequency;
    j->delta_amplitude = v->jitter_delta_amplitude;
}

/* src/lib.rs:753-777 */
static void jitter_apply(jitter_t *j, elem_t *elem)
{
    float freq = vn_next(&j->freq_noise, j->frequency);
    arr_t formant_freq = avn_next(&j->formant_freq_noise, j->frequency);
    arr_t formant_amp = avn_next(&j->formant_amp_noise, j->frequency);
    elem->frequency += freq * j->delta_frequency;                                       /* :763 */
    elem->formant_freq = a_add(elem->formant_freq, a_mul(formant_freq, a_splat(j->delta_formant_freq)));
    arr_t delta = a_mul(a_add(formant_amp, a_splat(1.0f)), a_splat(0.5f * j->delta_amplitude)); /* :768-769 */
    arr_t mul = a_sub(a_splat(1.0f), delta);                                            /* :772 */
    elem->formant_amp = a_mul(elem->formant_amp, mul);                                  /* :773 */
}

/* ------------------------------------------------------------------------------------------
 * Synthesize, src/lib.rs:470-600
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float phase;
    arr_t a, b, c;
    uint32_t seed;
} synth_t;

/* src/lib.rs:587-596 */
static void synth_new(synth_t *s, uint32_t seed)
{
    s->phase = 0.0f;
    s->a = a_splat(0.0f);
    s->b = a_splat(0.0f);
    s->c = a_splat(0.0f);
    s->seed = seed;
}

/* src/lib.rs:497-578 */
static float synth_next(synth_t *s, const elem_t *elem)
{
    float polyblep;
    if (s->phase < elem->frequency) {                               /* :503-506 */
        float t = s->phase / elem->frequency;
        polyblep = 2.0f * t - (t * t) - 1.0f;
    } else if (s->phase > (1.0f - elem->frequency)) {               /* :507-510 */
        float t = (s->phase - 1.0f) / elem->frequency;
        polyblep = (t * t) + 2.0f * t + 1.0f;
    } else {
        polyblep = 0.0f;
    }
    arr_t saw_wave = a_splat((2.0f * s->phase - 1.0f) - polyblep);  /* :517 */
    s->phase += elem->frequency;                                    /* :520 */
    if (s->phase >= 1.0f) s->phase -= 1.0f;                         /* :523-525 */
    arr_t noise = a_splat(grail_oracle_random_f32(&s->seed));       /* :528 */
    arr_t noise_wave = a_blend_multiple(saw_wave, noise, elem->formant_breath); /* :531 */
    arr_t alpha;
    for (int i = 0; i < NF; i++) alpha.v[i] = grail_oracle_exp_approx(elem->formant_smooth.v[i]); /* :535 */
    s->a = a_add(s->a, a_mul(a_sub(a_splat(1.0f), alpha), a_sub(noise_wave, s->a))); /* :538 */
    arr_t glottal_wave = s->a;
    arr_t turbulence_wave = a_mul(glottal_wave, a_blend_multiple(a_splat(1.0f), noise, elem->formant_turb)); /* :544-545 */
    arr_t v0 = a_mul(turbulence_wave, elem->formant_amp);           /* :550 */
    arr_t g;
    for (int i = 0; i < NF; i++) g.v[i] = grail_oracle_tan_approx(elem->formant_freq.v[i]); /* :555 */
    arr_t k = a_div(elem->formant_bw, elem->formant_freq);          /* :558 */
    arr_t a1 = a_div(a_splat(1.0f), a_add(a_splat(1.0f), a_mul(g, a_add(g, k)))); /* :560 */
    arr_t a2 = a_mul(g, a1);
    arr_t a3 = a_mul(g, a2);
    arr_t v3 = a_sub(v0, s->c);                                     /* :565 */
    arr_t v1 = a_add(a_mul(a1, s->b), a_mul(a2, v3));               /* :566 */
    arr_t v2 = a_add(a_add(s->c, a_mul(a2, s->b)), a_mul(a3, v3));  /* :567 */
    s->b = a_sub(a_mul(a_splat(2.0f), v1), s->b);                   /* :570 */
    s->c = a_sub(a_mul(a_splat(2.0f), v2), s->c);                   /* :571 */
    return a_sum(v1) * 0.5f;                                        /* :574 */
}

/* ------------------------------------------------------------------------------------------
 * The whole chain: elems.sequence(voice).jitter(seed, voice).synthesize() drained into out.
 * Returns the number of samples the iterator yields; writes min(n, cap) of them.
 * If final_states is non-NULL it receives {carrier phase, a[8], b[8], c[8], jitter phase,
 * sequencer time} (27 floats) followed by the three LCG states as raw bits (3 words) = 30 words.
 * ---------------------------------------------------------------------------------------- */
uint64_t grail_oracle_synthesize(const seq_elem_t *elems, uint32_t n_elems, const voice_params_t *voice,
                                 float *out, uint64_t cap, const trace_t *trace, uint32_t *final_states)
{
    sequencer_t sq;
    jitter_t jt;
    synth_t sy;
    sequencer_new(&sq, elems, n_elems, voice->sample_rate);
    jitter_new(&jt, voice);
    synth_new(&sy, voice->synth_seed);
    uint64_t n = 0;
    elem_t e;
    float alpha;
    while (sequencer_next(&sq, &e, &alpha)) {
        jitter_apply(&jt, &e);
        if (trace && n < cap) {
            if (trace->time) trace->time[n] = sq.time;
            if (trace->alpha) trace->alpha[n] = alpha;
            if (trace->phoneme_index) trace->phoneme_index[n] = sq.cur_index;
            if (trace->jitter_phase) trace->jitter_phase[n] = jt.freq_noise.phase;
            if (trace->frequency) trace->frequency[n] = e.frequency;
            if (trace->carrier_phase) trace->carrier_phase[n] = sy.phase; /* phase BEFORE this sample */
        }
        float y = synth_next(&sy, &e);
        if (n < cap) {
            if (out) out[n] = y;
            if (trace) {
                if (trace->state_a) memcpy(trace->state_a + n * NF, sy.a.v, sizeof(arr_t));
                if (trace->state_b) memcpy(trace->state_b + n * NF, sy.b.v, sizeof(arr_t));
                if (trace->state_c) memcpy(trace->state_c + n * NF, sy.c.v, sizeof(arr_t));
            }
        }
        n++;
    }
    if (final_states) {
        float f[27];
        f[0] = sy.phase;
        memcpy(f + 1, sy.a.v, 32);
        memcpy(f + 9, sy.b.v, 32);
        memcpy(f + 17, sy.c.v, 32);
        f[25] = jt.freq_noise.phase;
        f[26] = sq.time;
        memcpy(final_states, f, sizeof f);
        final_states[27] = sy.seed;
        final_states[28] = jt.freq_noise.state;
        final_states[29] = jt.formant_amp_noise.state;
    }
    return n;
}

/* sample count only (no DSP): drains the Sequencer alone, src/lib.rs:859-932 */
uint64_t grail_oracle_count_samples(const seq_elem_t *elems, uint32_t n_elems, float sample_rate)
{
    sequencer_t sq;
    sequencer_new(&sq, elems, n_elems, sample_rate);
    uint64_t n = 0;
    for (;;) {
        /* only the clock and the hand-over matter for the count */
        sq.time -= sq.delta_time;
        if (sq.time < 0.0f) {
            if (sq.has_cur && sq.has_next) {
                float len = sq.next.length;
                sq.cur = sq.next;
                sq.has_next = upstream_next(&sq, &sq.next);
                sq.time += len;
            } else if (!sq.has_cur && !sq.has_next) {
                sq.has_cur = upstream_next(&sq, &sq.cur);
                sq.has_next = upstream_next(&sq, &sq.next);
                if (sq.has_cur) sq.time += sq.cur.length;
            } else {
                break;
            }
        }
        if (!sq.has_cur) break;
        n++;
    }
    return n;
}

/* jitter construction facts used by the known-answer tests */
void grail_oracle_jitter_init_states(uint32_t seed, uint32_t *states3, float *freq_cur_next2)
{
    voice_params_t v;
    memset(&v, 0, sizeof v);
    v.jitter_seed = seed;
    jitter_t j;
    jitter_new(&j, &v);
    states3[0] = j.freq_noise.state;
    states3[1] = j.formant_freq_noise.state;
    states3[2] = j.formant_amp_noise.state;
    freq_cur_next2[0] = j.freq_noise.current;
    freq_cur_next2[1] = j.freq_noise.next;
}

/* ------------------------------------------------------------------------------------------
 * Batch driver: one utterance per task over n_threads host threads.  This is the CPU baseline
 * ("the reference's CPU path on the box's host cores"); utterances are independent so this is
 * exactly what running the reference once per utterance on a thread pool does.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const seq_elem_t *elems;
    const uint32_t *utt_offsets;
    const voice_params_t *voices;
    uint32_t n_utts;
    float *out;
    const uint64_t *out_offsets;
    uint64_t *counts;
    volatile uint32_t *next_utt;
} batch_job_t;

static void *batch_worker(void *arg)
{
    batch_job_t *job = (batch_job_t *)arg;
    for (;;) {
        uint32_t u = __atomic_fetch_add(job->next_utt, 1u, __ATOMIC_RELAXED);
        if (u >= job->n_utts) break;
        uint64_t cap = job->out_offsets[u + 1] - job->out_offsets[u];
        uint64_t n = grail_oracle_synthesize(job->elems + job->utt_offsets[u],
                                             job->utt_offsets[u + 1] - job->utt_offsets[u],
                                             job->voices + u, job->out + job->out_offsets[u], cap, NULL, NULL);
        if (job->counts) job->counts[u] = n;
    }
    return NULL;
}

int grail_oracle_synthesize_batch(const seq_elem_t *elems, const uint32_t *utt_offsets,
                                  const voice_params_t *voices, uint32_t n_utts, float *out,
                                  const uint64_t *out_offsets, uint64_t *counts, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    volatile uint32_t next = 0;
    batch_job_t job = { elems, utt_offsets, voices, n_utts, out, out_offsets, counts, &next };
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    if (!th) return -1;
    int started = 0;
    for (int i = 0; i < n_threads - 1; i++) {
        if (pthread_create(&th[started], NULL, batch_worker, &job) == 0) started++;
    }
    batch_worker(&job);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
    return 0;
}

/* word-wise FNV-1a over f32 bit patterns */
uint32_t grail_oracle_fnv(const float *x, uint64_t n)
{
    uint32_t h = 2166136261u;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t b;
        memcpy(&b, x + i, 4);
        h = (h ^ b) * 16777619u;
    }
    return h;
}

uint32_t grail_oracle_sizeof_seq_elem(void) { return (uint32_t)sizeof(seq_elem_t); }
uint32_t grail_oracle_sizeof_voice_params(void) { return (uint32_t)sizeof(voice_params_t); }
