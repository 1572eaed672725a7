/*
 * aecm_oracle.c -- TEST INFRASTRUCTURE (see aecm_oracle.h).  Plain-C restatement of the AECM hot
 * path of cpuimage/WebRTC_AECM: 10 ms API wrapper -> 80-sample frames -> 64-sample blocks -> block
 * DSP.  Compile with -fwrapv (oracle/Makefile does) so the few intentional modular operations are
 * defined.  Sequential, one instance at a time, no attempt at speed.
 *
 * Citations are reference file:line under /root/reference/aecm.
 */
#include "aecm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "aecm_oracle_tables.h"

/* algorithm constants, aecm_defines.h:28-85 */
#define CONV_LEN 512
#define FAR_ENERGY_MIN 1025
#define FAR_ENERGY_DIFF 929
#define ENERGY_DEV_TOL 400
#define FAR_ENERGY_VAD_REGION 230
#define MU_MIN 10
#define MU_MAX 1
#define MU_DIFF 9
#define MIN_MSE_COUNT 20
#define MIN_MSE_DIFF 29
#define MSE_RESOLUTION 5
#define RES_CH16 12
#define RES_CH32 28
#define CHANNEL_VAD 16
#define RES_SUPGAIN 8
#define SUPGAIN_DEFAULT 256
#define SUPGAIN_A 3072
#define SUPGAIN_B 1536
#define SUPGAIN_D 256
#define SUPGAIN_EPC_DT 200
#define ONE_Q14 16384
#define NLP_LOW 3277
#define INIT_MAGIC 42 /* echo_control_mobile.cc:40 */

/* ------------------------------------------------------------------------------------------ */
/* scalar primitives                                                                          */
/* ------------------------------------------------------------------------------------------ */

static int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }

int oracle_norm_w32(int32_t a) { /* spl_inl.h:96-98: 0 for a==0 */
    if (a == 0) return 0;
    return clz32((uint32_t)(a < 0 ? ~a : a)) - 1;
}
int oracle_norm_u32(uint32_t a) { /* spl_inl.h:101-103 */
    return a == 0 ? 0 : clz32(a);
}
int oracle_norm_w16(int16_t a) { /* spl_inl.h:107-111 */
    int32_t w = a;
    if (a == 0) return 0;
    return clz32((uint32_t)(w < 0 ? ~w : w)) - 17;
}

static int32_t add_sat32(int32_t a, int32_t b) { /* spl_inl.h:70-82 */
    int64_t s = (int64_t)a + b;
    if (s > INT32_MAX) return INT32_MAX;
    if (s < INT32_MIN) return INT32_MIN;
    return (int32_t)s;
}
static int16_t sat16(int32_t v) { /* spl_inl.h:59-68 */
    return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
}
static int16_t neg16(int16_t v) { return (int16_t)(-(int32_t)v); } /* -(-32768) wraps back */
/* WEBRTC_SPL_ABS_W16 stored into an int16_t (aecm_core_c.cc:350-351): |-32768| stays -32768 */
static int16_t abs16_wrap(int16_t v) { return v >= 0 ? v : neg16(v); }

/* WEBRTC_SPL_SHIFT_W32 on signed / unsigned operands, signal_processing_library.h:130.  One call site
 * (aecm_core.cc:911, the NLMS rescaling) reaches shift counts of 32 and more on loud-far / near-silent input --
 * undefined in C; the reference library as compiled for x86 executes SAR/SHL, which use the count modulo 32
 * (verified with a UBSan build of the reference: reports at aecm_core.cc:911, and this restatement stays
 * bit-identical to both the -O1 UBSan and the -O3 builds).  Counts are therefore masked explicitly. */
static unsigned g_wide_shift_events; /* how often a count outside [-31, 31] was seen (test hook) */
unsigned oracle_debug_wide_shift_events(void) { return g_wide_shift_events; }
static int32_t shift_s32(int32_t x, int c) {
    if (c > 31 || c < -31) g_wide_shift_events++;
    return c >= 0 ? (int32_t)((uint32_t)x << (c & 31)) : x >> ((-c) & 31);
}
static uint32_t shift_u32(uint32_t x, int c) { return c >= 0 ? x << (c & 31) : x >> ((-c) & 31); }

static int32_t div_w32_w16(int32_t num, int16_t den) { /* signal_processing_library.cc:116-123 */
    return den != 0 ? num / den : (int32_t)0x7FFFFFFF;
}
static uint32_t div_u32_u16(uint32_t num, uint16_t den) { /* signal_processing_library.cc:107-114 */
    return den != 0 ? num / den : 0xFFFFFFFFu;
}

int32_t oracle_sqrt_floor(int32_t v) {
    /* signal_processing_library.cc:84-105 computes floor(sqrt(v)) bit by bit for v in [0, 2^31),
     * and 0 for negative v (verified exhaustively).  Restated through the
     * double-precision root with an integer correction. */
    int64_t r;
    if (v <= 0) return 0;
    r = (int64_t)sqrt((double)v);
    while (r * r > v) r--;
    while ((r + 1) * (r + 1) <= v) r++;
    return (int32_t)r;
}

void oracle_rand_u_array(int16_t *v, int n, uint32_t *seed) { /* signal_processing_library.cc:129-147 */
    int i;
    for (i = 0; i < n; i++) {
        *seed = (*seed * 69069u + 1u) & 0x7FFFFFFFu;
        v[i] = (int16_t)(*seed >> 16);
    }
}

static int16_t max_abs16(const int16_t *v, int n) { /* signal_processing_library.cc:154-174 */
    int i, m = 0;
    for (i = 0; i < n; i++) {
        int a = v[i] < 0 ? -(int)v[i] : (int)v[i];
        if (a > m) m = a;
    }
    return (int16_t)(m > 32767 ? 32767 : m);
}

/* ------------------------------------------------------------------------------------------ */
/* FIFO with movable read pointer (ring_buffer.c)                                             */
/* ------------------------------------------------------------------------------------------ */

static void fifo_init(OrFifo *f, int16_t *store, int cap) { /* ring_buffer.c:76-83 */
    f->cap = cap;
    f->rd = 0;
    f->cnt = 0;
    memset(store, 0, sizeof(int16_t) * (size_t)cap);
}
static int fifo_write(OrFifo *f, int16_t *store, const int16_t *src, int n) { /* ring_buffer.c:140-175 */
    int room = f->cap - f->cnt, k;
    if (n > room) n = room; /* a full FIFO silently drops the excess */
    for (k = 0; k < n; k++) store[(f->rd + f->cnt + k) % f->cap] = src[k];
    f->cnt += n;
    return n;
}
static int fifo_move(OrFifo *f, int k) { /* ring_buffer.c:177-212 */
    int room = f->cap - f->cnt;
    if (k > f->cnt) k = f->cnt;
    if (k < -room) k = -room;
    f->rd = ((f->rd + k) % f->cap + f->cap) % f->cap;
    f->cnt -= k;
    return k;
}
static int fifo_read(OrFifo *f, const int16_t *store, int16_t *dst, int n) { /* ring_buffer.c:97-138 */
    int k;
    if (n > f->cnt) n = f->cnt;
    for (k = 0; k < n; k++) dst[k] = store[(f->rd + k) % f->cap];
    fifo_move(f, n);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* 128-point fixed-point complex FFT / IFFT (complex_fft.c)                                   */
/* ------------------------------------------------------------------------------------------ */

static int bitrev7(int i) { /* complex_fft.c:181-209: a true 7-bit reversal for stages==7 */
    int r = 0, b;
    for (b = 0; b < 7; b++) r |= ((i >> b) & 1) << (6 - b);
    return r;
}

/* In-place radix-2 decimation-in-time transform on 128 complex int16 values that are already in
 * bit-reversed order.  Forward: complex_fft.c:293-357 (mode 1), every stage halves.  Inverse:
 * complex_fft.c:381-488 (mode 1) with the data-dependent per-stage shift; returns the scale. */
static int cfft128(int16_t *re, int16_t *im, int inverse) {
    int scale = 0, s;
    for (s = 0; s < 7; s++) {
        const int half = 1 << s;
        int shift = 1, m;
        int32_t rnd = 16384;
        if (inverse) { /* complex_fft.c:383-396 */
            int16_t a = max_abs16(re, 128), b = max_abs16(im, 128);
            int32_t peak = a > b ? a : b;
            shift = 0;
            rnd = 8192;
            if (peak > 13573) { shift++; scale++; rnd <<= 1; }
            if (peak > 27146) { shift++; scale++; rnd <<= 1; }
        }
        for (m = 0; m < half; m++) {
            const int t = m << (6 - s); /* table index j = m << (9-s) = 8*t */
            const int32_t wr = kOracleSin8[t + 32];
            const int32_t wi = inverse ? kOracleSin8[t] : -kOracleSin8[t];
            int i;
            for (i = m; i < 128; i += 2 * half) {
                const int j = i + half;
                /* complex_fft.c:332-350 / :465-482 (CFFTSFT = CIFFTSFT = 14, rounding constant 1) */
                const int32_t tr = (wr * re[j] - wi * im[j] + 1) >> 1;
                const int32_t ti = (wr * im[j] + wi * re[j] + 1) >> 1;
                const int32_t qr = (int32_t)re[i] * 16384;
                const int32_t qi = (int32_t)im[i] * 16384;
                const int sh = inverse ? shift + 14 : 15;
                re[j] = (int16_t)((qr - tr + rnd) >> sh);
                im[j] = (int16_t)((qi - ti + rnd) >> sh);
                re[i] = (int16_t)((qr + tr + rnd) >> sh);
                im[i] = (int16_t)((qi + ti + rnd) >> sh);
            }
        }
    }
    return scale;
}

void oracle_real_fft128_forward(const int16_t in[128], int16_t out[130]) { /* real_fft.c:47-72 */
    int16_t re[128], im[128];
    int i;
    for (i = 0; i < 128; i++) {
        re[bitrev7(i)] = in[i];
        im[i] = 0;
    }
    cfft128(re, im, 0);
    for (i = 0; i <= 64; i++) {
        out[2 * i] = re[i];
        out[2 * i + 1] = im[i];
    }
}

int oracle_real_fft128_inverse(const int16_t in[130], int16_t out[128]) { /* real_fft.c:74-102 */
    int16_t re[128], im[128];
    int k, scale;
    for (k = 0; k <= 64; k++) {
        re[bitrev7(k)] = in[2 * k];
        im[bitrev7(k)] = in[2 * k + 1];
    }
    for (k = 65; k < 128; k++) { /* conjugate-symmetric upper half, real_fft.c:87-91 */
        re[bitrev7(k)] = in[2 * (128 - k)];
        im[bitrev7(k)] = neg16(in[2 * (128 - k) + 1]);
    }
    scale = cfft128(re, im, 1);
    memcpy(out, re, sizeof(re));
    return scale;
}

/* ------------------------------------------------------------------------------------------ */
/* block core helpers                                                                         */
/* ------------------------------------------------------------------------------------------ */

/* aecm_core_c.cc:261-365 with WindowAndFFT :166-191. Fills re/im (65 bins, sign convention after
 * the conjugation at :188-190), magnitudes and their sum; returns the dynamic-Q shift. */
static int time_to_frequency(const int16_t t[128], int16_t re[OR_BINS], int16_t im[OR_BINS],
                             uint16_t mag[OR_BINS], uint32_t *mag_sum) {
    int16_t win[128], spec[130];
    const int q = oracle_norm_w16(max_abs16(t, 128)); /* :288-289 */
    uint32_t sum;
    int i;
    for (i = 0; i < 64; i++) { /* :174-182: two int16 truncations per sample */
        int16_t s = (int16_t)(t[i] * (1 << q));
        win[i] = (int16_t)((s * kOracleSqrtHann[i]) >> 14);
        s = (int16_t)(t[64 + i] * (1 << q));
        win[64 + i] = (int16_t)((s * kOracleSqrtHann[64 - i]) >> 14);
    }
    oracle_real_fft128_forward(win, spec);
    for (i = 0; i <= 64; i++) {
        re[i] = spec[2 * i];
        im[i] = i < 64 ? neg16(spec[2 * i + 1]) : spec[2 * i + 1]; /* :188-190 touches bins 0..63 */
    }
    im[0] = 0; /* :296-297 */
    im[64] = 0;
    mag[0] = (uint16_t)(re[0] >= 0 ? re[0] : -(int32_t)re[0]); /* :298-300 */
    mag[64] = (uint16_t)(re[64] >= 0 ? re[64] : -(int32_t)re[64]);
    sum = (uint32_t)mag[0] + mag[64];
    for (i = 1; i < 64; i++) { /* :304-362 */
        if (re[i] == 0) {
            mag[i] = (uint16_t)(im[i] >= 0 ? im[i] : -(int32_t)im[i]);
        } else if (im[i] == 0) {
            mag[i] = (uint16_t)(re[i] >= 0 ? re[i] : -(int32_t)re[i]);
        } else {
            const int16_t a = abs16_wrap(re[i]), b = abs16_wrap(im[i]);
            mag[i] = (uint16_t)oracle_sqrt_floor(add_sat32(a * a, b * b));
        }
        sum += mag[i];
    }
    *mag_sum = sum;
    return q;
}

static void mean_estimator_fix(int32_t v, int factor, int32_t *mean) { /* delay_estimator.cc:690-702 */
    int32_t d = v - *mean;
    d = d < 0 ? -((-d) >> factor) : d >> factor;
    *mean += d;
}

/* delay_estimator_wrapper.cc:92-125, bands 12..43 -> bits 0..31 */
static uint32_t binary_spectrum(const uint16_t *spec, int32_t *thr, int q, int32_t *initialized) {
    uint32_t bits = 0;
    int b;
    if (!*initialized) {
        for (b = 12; b <= 43; b++) {
            if (spec[b] > 0) {
                thr[b] = (int32_t)(((uint32_t)spec[b] << (15 - q)) >> 1);
                *initialized = 1;
            }
        }
    }
    for (b = 12; b <= 43; b++) {
        const int32_t s15 = (int32_t)((uint32_t)spec[b] << (15 - q));
        mean_estimator_fix(s15, 6, &thr[b]);
        if (s15 > thr[b]) bits |= 1u << (b - 12);
    }
    return bits;
}

static uint32_t far_word(const OracleAecm *a, int delay) { return a->bin_far[(a->bin_far_head + delay) % OR_HIST]; }

/* UpdateRobustValidationStatistics, delay_estimator.cc:96-154.  Single-precision arithmetic, one rounding per
 * operation, in the reference's order (the library is built without FMA contraction). */
static void update_robust_statistics(OracleAecm *a, int cand, int32_t depth_q14, int32_t level_q14) {
    const float q14 = 1.f / (1 << 14);
    const float depth = depth_q14 * q14;
    float decrease_in_last_set = depth;
    const int max_hits_for_slow_change = cand < a->last_delay ? 10 : 1000; /* :101-103 */
    int i;
    if (cand != a->last_candidate_delay) { /* :108-111 */
        a->candidate_hits = 0;
        a->last_candidate_delay = cand;
    }
    a->candidate_hits++;
    a->histogram[cand] += depth; /* :119-122 */
    if (a->histogram[cand] > 3000.f) a->histogram[cand] = 3000.f;
    if (a->candidate_hits < max_hits_for_slow_change) /* :132-136 */
        decrease_in_last_set = (a->mean_bit_counts[a->compare_delay] - level_q14) * q14;
    for (i = 0; i < OR_HIST; i++) { /* :140-153 */
        const int in_last = i >= a->last_delay - 2 && i <= a->last_delay + 1 && i != cand;
        const int in_cand = i >= cand - 2 && i <= cand + 1;
        a->histogram[i] -= decrease_in_last_set * in_last + depth * (!in_last && !in_cand);
        if (a->histogram[i] < 0) a->histogram[i] = 0;
    }
}

static int histogram_based_validation(const OracleAecm *a, int cand) { /* delay_estimator.cc:178-225 */
    float fraction = 1.f;
    float threshold = a->histogram[a->compare_delay];
    const int diff = cand - a->last_delay;
    if (diff > a->allowed_offset) {
        fraction = 1.f - 0.05f * (diff - a->allowed_offset);
        fraction = fraction > 0.5f ? fraction : 0.5f;
    } else if (diff < 0) {
        fraction = 0.25f - 0.05f * diff;
        fraction = fraction > 1.f ? 1.f : fraction;
    }
    threshold *= fraction;
    threshold = threshold > 1.5f ? threshold : 1.5f;
    return a->histogram[cand] >= threshold && a->candidate_hits > 10;
}

static int robust_validation(const OracleAecm *a, int cand, int instantaneous, int by_histogram) { /* :245-266 */
    int robust = a->last_delay < 0 && (instantaneous || by_histogram);
    robust |= instantaneous && by_histogram;
    robust |= by_histogram && a->histogram[cand] > a->last_delay_histogram;
    return robust;
}

/* WebRtc_ProcessBinarySpectrum, delay_estimator.cc:521-664.  AECM as shipped runs it with robust validation off and
 * lookahead 0 (aecm_core.cc:218,225); the robust statistics are collected all the same (:628-633). */
static int process_binary_spectrum(OracleAecm *a, uint32_t near_bits) {
    int32_t best = 32 << 9, worst = 0, depth;
    int cand = -1, valid, any_far = 0, d;
    if (a->near_history_size > 1) { /* lookahead: delay the near-end binary spectrum, :536-543 */
        memmove(&a->bin_near_hist[1], &a->bin_near_hist[0], (size_t)(a->near_history_size - 1) * sizeof(uint32_t));
        a->bin_near_hist[0] = near_bits;
        near_bits = a->bin_near_hist[a->lookahead];
    }
    for (d = 0; d < OR_HIST; d++) { /* :545-564 */
        const uint32_t fw = far_word(a, d);
        const int fbits = __builtin_popcount(fw);
        const int32_t bc = (int32_t)__builtin_popcount(near_bits ^ fw) << 9;
        if (fbits > 0) {
            mean_estimator_fix(bc, 13 - ((3 * fbits) >> 4), &a->mean_bit_counts[d]);
            any_far = 1; /* :623-626 */
        }
    }
    for (d = 0; d < OR_HIST; d++) { /* :568-577, first index wins ties */
        if (a->mean_bit_counts[d] < best) { best = a->mean_bit_counts[d]; cand = d; }
        if (a->mean_bit_counts[d] > worst) worst = a->mean_bit_counts[d];
    }
    depth = worst - best;
    if (a->minimum_probability > 8704 && depth > 2816) { /* :593-605 */
        int32_t thr = best + 1024;
        if (thr < 8704) thr = 8704;
        if (a->minimum_probability > thr) a->minimum_probability = thr;
    }
    a->last_delay_probability++; /* :608 */
    valid = depth > 1024 && (best < a->minimum_probability || best < a->last_delay_probability); /* :617-619 */
    if (any_far && cand >= 0) update_robust_statistics(a, cand, depth, best); /* :628-633 */
    if (a->robust_enabled && cand >= 0) /* :635-639 */
        valid = robust_validation(a, cand, valid, histogram_based_validation(a, cand));
    if (any_far && valid) { /* :643-661 */
        if (cand != a->last_delay) {
            a->last_delay_histogram = a->histogram[cand] > 250.f ? 250.f : a->histogram[cand];
            if (a->histogram[cand] < a->histogram[a->compare_delay]) a->histogram[a->compare_delay] = a->histogram[cand];
        }
        a->last_delay = cand;
        if (best < a->last_delay_probability) a->last_delay_probability = best;
        a->compare_delay = a->last_delay;
    }
    return a->last_delay;
}

/* WebRtc_InitBinaryDelayEstimator, delay_estimator.cc:483-504 (robust_validation_enabled, allowed_offset, lookahead
 * and near_history_size belong to the estimator OBJECT and survive Init) */
static void init_binary_delay_estimator(OracleAecm *a) {
    int i;
    if (a->near_history_size < 1) a->near_history_size = 1; /* a calloc'ed instance: WebRtc_CreateDelayEstimator(farend, 0) */
    memset(a->bin_near_hist, 0, sizeof(a->bin_near_hist));
    for (i = 0; i <= OR_HIST; i++) {
        a->mean_bit_counts[i] = 20 << 9;
        a->histogram[i] = 0.f;
    }
    a->minimum_probability = 32 << 9;
    a->last_delay_probability = 32 << 9;
    a->last_delay = -2;
    a->last_candidate_delay = -2;
    a->compare_delay = OR_HIST;
    a->candidate_hits = 0;
    a->last_delay_histogram = 0.f;
}

int32_t oracle_aecm_set_delay_estimator_options(OracleAecm *a, int robust, int allowed_offset, int max_lookahead, int lookahead) {
    if (!a) return -1;
    if (robust < 0 || robust > 1 || allowed_offset < 0 || max_lookahead > OR_MAX_LOOKAHEAD) return -1;
    if (max_lookahead >= 0) { /* WebRtc_CreateDelayEstimator(farend, max_lookahead) + WebRtc_InitDelayEstimator */
        a->near_history_size = max_lookahead + 1;
        a->lookahead = max_lookahead;
        a->robust_enabled = 0;
        a->allowed_offset = 0;
        memset(a->thr_near, 0, sizeof(a->thr_near));
        a->thr_near_init = 0;
        init_binary_delay_estimator(a);
    }
    if (a->near_history_size < 1) a->near_history_size = 1;
    if (lookahead >= 0) { /* WebRtc_set_lookahead */
        if (lookahead > a->near_history_size - 1) return -1;
        a->lookahead = lookahead;
    }
    a->robust_enabled = robust;
    a->allowed_offset = allowed_offset;
    return 0;
}

int32_t oracle_aecm_soft_reset_delay_estimator(OracleAecm *a, int delay_shift) {
    uint32_t shifted[OR_HIST];
    const int abs_shift = delay_shift < 0 ? -delay_shift : delay_shift;
    int i, old;
    if (!a || abs_shift >= OR_HIST) return -1;
    if (delay_shift != 0) { /* far end, delay_estimator.cc:336-367: shift the binary history, zero-fill the gap */
        for (i = 0; i < OR_HIST; i++) {
            const int src = i - delay_shift;
            shifted[i] = (src >= 0 && src < OR_HIST) ? far_word(a, src) : 0u;
        }
        memcpy(a->bin_far, shifted, sizeof(shifted));
        a->bin_far_head = 0;
    }
    if (a->near_history_size < 1) a->near_history_size = 1;
    old = a->lookahead; /* near end, delay_estimator.cc:506-519 */
    a->lookahead -= delay_shift;
    if (a->lookahead < 0) a->lookahead = 0;
    if (a->lookahead > a->near_history_size - 1) a->lookahead = a->near_history_size - 1;
    return old - a->lookahead;
}

int32_t oracle_aecm_last_delay(const OracleAecm *a) { return a ? a->last_delay : -1; }

float oracle_aecm_last_delay_quality(const OracleAecm *a) { /* delay_estimator.cc:671-688 */
    float q;
    if (a->robust_enabled) return a->histogram[a->compare_delay] / 3000.f;
    q = (float)((32 << 9) - a->last_delay_probability) / (32 << 9);
    return q < 0 ? 0 : q;
}

static int16_t log_energy_q8(uint32_t e, int q) { /* aecm_core.cc:612-628 */
    int16_t r = 7 << 7;
    if (e > 0) {
        const int z = oracle_norm_u32(e);
        const int16_t frac = (int16_t)(((e << z) & 0x7FFFFFFFu) >> 23);
        r = (int16_t)(r + ((31 - z) << 8) + frac - (q << 8));
    }
    return r;
}

static int16_t asym_filt(int16_t old, int16_t in, int16_t step_pos, int16_t step_neg) { /* aecm_core.cc:588-605 */
    int16_t r = old;
    if (old == 32767 || old == -32768) return in;
    if (old > in) r = (int16_t)(r - ((old - in) >> step_neg));
    else r = (int16_t)(r + ((in - old) >> step_pos));
    return r;
}

/* aecm_core.cc:644-755 */
static void calc_energies(OracleAecm *a, const uint16_t *far_spec, int16_t far_q, uint32_t near_energy,
                          int32_t *echo_est) {
    uint32_t e_far = 0, e_adapt = 0, e_stored = 0;
    int16_t inc_max = 4, dec_max = 11, inc_min = 11, dec_min = 3, t16;
    int i;
    memmove(a->near_log + 1, a->near_log, sizeof(int16_t) * (OR_ELOG - 1));
    a->near_log[0] = log_energy_q8(near_energy, a->q_noisy);
    for (i = 0; i < OR_BINS; i++) { /* :267-284 */
        echo_est[i] = (int32_t)a->ch_stored[i] * far_spec[i];
        e_far += far_spec[i];
        e_adapt += (uint32_t)((int32_t)a->ch_adapt16[i] * far_spec[i]);
        e_stored += (uint32_t)echo_est[i];
    }
    memmove(a->echo_adapt_log + 1, a->echo_adapt_log, sizeof(int16_t) * (OR_ELOG - 1));
    memmove(a->echo_stored_log + 1, a->echo_stored_log, sizeof(int16_t) * (OR_ELOG - 1));
    a->far_log = log_energy_q8(e_far, far_q);
    a->echo_adapt_log[0] = log_energy_q8(e_adapt, RES_CH16 + far_q);
    a->echo_stored_log[0] = log_energy_q8(e_stored, RES_CH16 + far_q);

    if (a->far_log > FAR_ENERGY_MIN) { /* :692-730 */
        if (a->startup_state == 0) { inc_max = 2; dec_min = 2; inc_min = 8; }
        a->far_e_min = asym_filt(a->far_e_min, a->far_log, inc_min, dec_min);
        a->far_e_max = asym_filt(a->far_e_max, a->far_log, inc_max, dec_max);
        a->far_e_maxmin = (int16_t)(a->far_e_max - a->far_e_min);
        t16 = (int16_t)(2560 - a->far_e_min);
        t16 = t16 > 0 ? (int16_t)((t16 * FAR_ENERGY_VAD_REGION) >> 9) : 0;
        t16 = (int16_t)(t16 + FAR_ENERGY_VAD_REGION);
        if ((a->startup_state == 0) | (a->vad_update_count > 1024)) {
            a->far_e_vad = (int16_t)(a->far_e_min + t16);
        } else if (a->far_e_vad > a->far_log) {
            a->far_e_vad = (int16_t)(a->far_e_vad + ((a->far_log + t16 - a->far_e_vad) >> 6));
            a->vad_update_count = 0;
        } else {
            a->vad_update_count++;
        }
        a->far_e_mse = (int16_t)(a->far_e_vad + (1 << 8));
    }
    if (a->far_log > a->far_e_vad) { /* :733-740 */
        if ((a->startup_state == 0) | (a->far_e_maxmin > FAR_ENERGY_DIFF)) a->current_vad = 1;
    } else {
        a->current_vad = 0;
    }
    if (a->current_vad && a->first_vad) { /* :741-754 */
        a->first_vad = 0;
        if (a->echo_adapt_log[0] > a->near_log[0]) {
            for (i = 0; i < OR_BINS; i++) a->ch_adapt16[i] >>= 3;
            a->echo_adapt_log[0] = (int16_t)(a->echo_adapt_log[0] - (3 << 8));
            a->first_vad = 1;
        }
    }
}

static int16_t calc_step_size(const OracleAecm *a) { /* aecm_core.cc:767-794 */
    int16_t mu = MU_MAX;
    if (!a->current_vad) {
        mu = 0;
    } else if (a->startup_state > 0) {
        if (a->far_e_min >= a->far_e_max) {
            mu = MU_MIN;
        } else {
            const int16_t t16 = (int16_t)(a->far_log - a->far_e_min);
            const int32_t t32 = div_w32_w16(t16 * MU_DIFF, a->far_e_maxmin);
            mu = (int16_t)(MU_MIN - 1 - (int16_t)t32);
        }
        if (mu < MU_MAX) mu = MU_MAX;
    }
    return mu;
}

static void store_adaptive_channel(OracleAecm *a, const uint16_t *far_spec, int32_t *echo_est) { /* aecm_core.cc:286-306 */
    int i;
    memcpy(a->ch_stored, a->ch_adapt16, sizeof(a->ch_stored));
    for (i = 0; i < OR_BINS; i++) echo_est[i] = (int32_t)a->ch_stored[i] * far_spec[i];
}

/* aecm_core.cc:810-986 */
static void update_channel(OracleAecm *a, const uint16_t *far_spec, int16_t far_q, const uint16_t *dfa,
                           int16_t mu, int32_t *echo_est) {
    int i;
    if (mu) {
        for (i = 0; i < OR_BINS; i++) { /* :831-921 */
            uint32_t prod, dfa_al;
            int32_t err, upd;
            int16_t z_ch, z_far, z_num, z_dfa, sh_chfar, sh_num, sh2, t16, xq, dq;
            z_ch = (int16_t)oracle_norm_u32((uint32_t)a->ch_adapt32[i]);
            z_far = (int16_t)oracle_norm_u32(far_spec[i]);
            if (z_ch + z_far > 31) {
                prod = (uint32_t)a->ch_adapt32[i] * far_spec[i];
                sh_chfar = 0;
            } else {
                sh_chfar = (int16_t)(32 - z_ch - z_far);
                /* :847-849: int32 >> shift (arithmetic), converted to uint32, times the uint16 */
                prod = (uint32_t)(sh_chfar >= 32 ? 0 : a->ch_adapt32[i] >> sh_chfar) * far_spec[i];
            }
            z_num = (int16_t)oracle_norm_u32(prod);
            z_dfa = dfa[i] ? (int16_t)oracle_norm_u32(dfa[i]) : 32;
            t16 = (int16_t)(z_dfa - 2 + a->q_noisy - RES_CH32 - far_q + sh_chfar);
            if (z_num > t16 + 1) {
                xq = t16;
                dq = (int16_t)(z_dfa - 2);
            } else {
                xq = (int16_t)(z_num - 2);
                dq = (int16_t)(RES_CH32 + far_q - a->q_noisy - sh_chfar + xq);
            }
            prod = shift_u32(prod, xq);              /* :869 */
            dfa_al = shift_u32((uint32_t)dfa[i], dq); /* :870 */
            err = (int32_t)dfa_al - (int32_t)prod;
            z_num = (int16_t)oracle_norm_w32(err);
            if (err && (far_spec[i] > (CHANNEL_VAD << far_q))) { /* :873 */
                if (z_num + z_far > 31) { /* :886-902 */
                    upd = err > 0 ? (int32_t)((uint32_t)err * far_spec[i])
                                  : -(int32_t)((uint32_t)(-err) * far_spec[i]);
                    sh_num = 0;
                } else {
                    sh_num = (int16_t)(32 - (z_num + z_far));
                    upd = err > 0 ? (err >> sh_num) * far_spec[i] : -((-err >> sh_num) * far_spec[i]);
                }
                upd = div_w32_w16(upd, (int16_t)(i + 1)); /* :904 */
                sh2 = (int16_t)(sh_num + sh_chfar - xq - mu - ((30 - z_far) << 1));
                if (oracle_norm_w32(upd) < sh2) upd = INT32_MAX; /* :908-912 */
                else upd = shift_s32(upd, sh2);
                a->ch_adapt32[i] = add_sat32(a->ch_adapt32[i], upd);
                if (a->ch_adapt32[i] < 0) a->ch_adapt32[i] = 0;
                a->ch_adapt16[i] = (int16_t)(a->ch_adapt32[i] >> 16);
            }
        }
    }
    if ((a->startup_state == 0) & (a->current_vad != 0)) { /* :926-929 */
        store_adaptive_channel(a, far_spec, echo_est);
    } else {
        if (a->far_log < a->far_e_mse) a->mse_channel_count = 0;
        else a->mse_channel_count++;
        if (a->mse_channel_count >= MIN_MSE_COUNT + 10) { /* :937-983 */
            int32_t mse_stored = 0, mse_adapt = 0;
            for (i = 0; i < MIN_MSE_COUNT; i++) {
                int32_t d = (int32_t)a->echo_stored_log[i] - a->near_log[i];
                mse_stored += d >= 0 ? d : -d;
                d = (int32_t)a->echo_adapt_log[i] - a->near_log[i];
                mse_adapt += d >= 0 ? d : -d;
            }
            if (((mse_stored << MSE_RESOLUTION) < (MIN_MSE_DIFF * mse_adapt)) &
                ((a->mse_stored_old << MSE_RESOLUTION) < (MIN_MSE_DIFF * a->mse_adapt_old))) {
                /* reset the adaptive channel, aecm_core.cc:308-323 */
                memcpy(a->ch_adapt16, a->ch_stored, sizeof(a->ch_adapt16));
                for (i = 0; i < OR_BINS; i++) a->ch_adapt32[i] = (int32_t)((uint32_t)(int32_t)a->ch_stored[i] << 16);
            } else if (((MIN_MSE_DIFF * mse_stored) > (mse_adapt << MSE_RESOLUTION)) &
                       (mse_adapt < a->mse_threshold) & (a->mse_adapt_old < a->mse_threshold)) {
                store_adaptive_channel(a, far_spec, echo_est);
                if (a->mse_threshold == INT32_MAX) {
                    a->mse_threshold = mse_adapt + a->mse_adapt_old;
                } else {
                    const int scaled = a->mse_threshold * 5 / 8;
                    a->mse_threshold += ((mse_adapt - scaled) * 205) >> 8;
                }
            }
            a->mse_channel_count = 0;
            a->mse_stored_old = mse_stored;
            a->mse_adapt_old = mse_adapt;
        }
    }
}

static int16_t calc_suppression_gain(OracleAecm *a) { /* aecm_core.cc:1000-1052 */
    int16_t g = SUPGAIN_DEFAULT, t16;
    if (!a->current_vad) {
        g = 0;
    } else {
        int16_t de;
        t16 = (int16_t)(a->near_log[0] - a->echo_stored_log[0]);
        de = abs16_wrap(t16);
        if (de < ENERGY_DEV_TOL) {
            if (de < SUPGAIN_EPC_DT) {
                int32_t t32 = a->sup_diff_ab * de + (SUPGAIN_EPC_DT >> 1);
                g = (int16_t)(a->sup_a - (int16_t)div_w32_w16(t32, SUPGAIN_EPC_DT));
            } else {
                int32_t t32 = a->sup_diff_bd * (ENERGY_DEV_TOL - de) + ((ENERGY_DEV_TOL - SUPGAIN_EPC_DT) >> 1);
                g = (int16_t)(a->sup_d + (int16_t)div_w32_w16(t32, ENERGY_DEV_TOL - SUPGAIN_EPC_DT));
            }
        } else {
            g = a->sup_d;
        }
    }
    t16 = g > a->sup_gain_old ? g : a->sup_gain_old;
    a->sup_gain_old = g;
    a->sup_gain = (int16_t)(a->sup_gain + (int16_t)((t16 - a->sup_gain) >> 4));
    return a->sup_gain;
}

/* aecm_core_c.cc:52-164 */
static void comfort_noise(OracleAecm *a, const uint16_t *dfa, int16_t *ere, int16_t *eim, const int16_t *hnl) {
    int16_t rnd[OR_PART], noise16[OR_BINS];
    const int sh = 15 - a->q_clean; /* kNoiseEstQDomain - dfaCleanQDomain */
    int track, i;
    if (a->noise_est_ctr < 100) { a->noise_est_ctr++; track = 6; } else { track = 9; }
    for (i = 0; i < OR_BINS; i++) { /* :81-127 */
        const int32_t lvl = (int32_t)((uint32_t)dfa[i] << sh);
        if (lvl < a->noise_est[i]) {
            a->noise_low_ctr[i] = 0;
            if (a->noise_est[i] < (1 << track)) {
                a->noise_high_ctr[i]++;
                if (a->noise_high_ctr[i] >= 5) { a->noise_est[i]--; a->noise_high_ctr[i] = 0; }
            } else {
                a->noise_est[i] -= (a->noise_est[i] - lvl) >> track;
            }
        } else {
            a->noise_high_ctr[i] = 0;
            if ((a->noise_est[i] >> 19) > 0) {
                a->noise_est[i] >>= 11;
                a->noise_est[i] *= 2049;
            } else if ((a->noise_est[i] >> 11) > 0) {
                a->noise_est[i] *= 2049;
                a->noise_est[i] >>= 11;
            } else {
                a->noise_low_ctr[i]++;
                if (a->noise_low_ctr[i] >= 5) { a->noise_est[i] += (a->noise_est[i] >> 9) + 1; a->noise_low_ctr[i] = 0; }
            }
        }
    }
    for (i = 0; i < OR_BINS; i++) { /* :129-140 */
        int32_t t32 = a->noise_est[i] >> sh;
        int16_t w;
        if (t32 > 32767) { t32 = 32767; a->noise_est[i] = t32 << sh; }
        w = (int16_t)(ONE_Q14 - hnl[i]);
        noise16[i] = (int16_t)((w * (int16_t)t32) >> 14);
    }
    oracle_rand_u_array(rnd, OR_PART, &a->seed); /* :143 */
    for (i = 1; i < OR_BINS; i++) { /* :146-163; bin 0 receives no noise, bin 64 real part only */
        const int16_t idx = (int16_t)((359 * rnd[i - 1]) >> 15);
        const int16_t ur = (int16_t)((noise16[i] * kOracleCos360[idx]) >> 13);
        const int16_t ui = i < OR_PART ? (int16_t)((-noise16[i] * kOracleSin360[idx]) >> 13) : 0;
        ere[i] = sat16((int32_t)ere[i] + ur);
        eim[i] = sat16((int32_t)eim[i] + ui);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* one 64-sample block: aecm_core_c.cc:368-711                                                */
/* ------------------------------------------------------------------------------------------ */
int oracle_process_block(OracleAecm *a, const int16_t *far64, const int16_t *near64, const int16_t *clean64,
                         int16_t *out64) {
    int16_t xre[OR_BINS], xim[OR_BINS], dre[OR_BINS], dim[OR_BINS], ere[OR_BINS], eim[OR_BINS];
    uint16_t xfa[OR_BINS], dfa_noisy[OR_BINS], dfa_clean_buf[OR_BINS];
    const uint16_t *dfa_clean = dfa_clean_buf, *far_spec;
    uint32_t xfa_sum, dfa_noisy_sum, dfa_clean_sum;
    int32_t echo_est[OR_BINS];
    int16_t hnl[OR_BINS], mu, sup_gain, z_xbuf, num_pos = 0;
    int far_q, q_noisy, delay, pos, i;

    if (a->startup_state < 2) /* :420-424 */
        a->startup_state = (int16_t)((a->tot_count >= CONV_LEN) + (a->tot_count >= 2 * CONV_LEN));

    memcpy(a->x_buf + OR_PART, far64, sizeof(int16_t) * OR_PART); /* :428-436 */
    memcpy(a->d_noisy + OR_PART, near64, sizeof(int16_t) * OR_PART);
    if (clean64) memcpy(a->d_clean + OR_PART, clean64, sizeof(int16_t) * OR_PART);

    far_q = time_to_frequency(a->x_buf, xre, xim, xfa, &xfa_sum);               /* :439 */
    q_noisy = time_to_frequency(a->d_noisy, dre, dim, dfa_noisy, &dfa_noisy_sum); /* :442-447 */
    a->q_noisy_old = a->q_noisy;
    a->q_noisy = (int16_t)q_noisy;
    if (!clean64) { /* :449-464 */
        dfa_clean = dfa_noisy;
        a->q_clean_old = a->q_noisy_old;
        a->q_clean = a->q_noisy;
        dfa_clean_sum = dfa_noisy_sum;
    } else {
        /* the clean transform overwrites dfw: the Wiener gain is applied to the clean spectrum */
        const int qc = time_to_frequency(a->d_clean, dre, dim, dfa_clean_buf, &dfa_clean_sum);
        a->q_clean_old = a->q_clean;
        a->q_clean = (int16_t)qc;
    }
    (void)dfa_clean_sum;
    (void)xre; (void)xim;

    /* far magnitude history, aecm_core.cc:125-138 */
    a->far_hist_pos++;
    if (a->far_hist_pos >= OR_HIST) a->far_hist_pos = 0;
    a->far_hist_q[a->far_hist_pos] = far_q;
    memcpy(&a->far_hist[a->far_hist_pos * OR_BINS], xfa, sizeof(xfa));

    /* delay estimator: far word into the history (delay_estimator_wrapper.cc:233-262,
     * delay_estimator.cc:369-382), then the near word against it (:447-476) */
    {
        const uint32_t fbits = binary_spectrum(xfa, a->thr_far, far_q, &a->thr_far_init);
        uint32_t nbits;
        a->bin_far_head = (a->bin_far_head + OR_HIST - 1) % OR_HIST;
        a->bin_far[a->bin_far_head] = fbits;
        nbits = binary_spectrum(dfa_noisy, a->thr_near, q_noisy, &a->thr_near_init);
        delay = process_binary_spectrum(a, nbits);
    }
    if (delay == -2) delay = 0; /* aecm_core_c.cc:479-483; fixedDelay is -1 forever (aecm_core.cc:400) */

    pos = a->far_hist_pos - delay; /* aecm_core.cc:157-172 */
    if (pos < 0) pos += OR_HIST;
    z_xbuf = (int16_t)a->far_hist_q[pos];
    far_spec = &a->far_hist[pos * OR_BINS];

    calc_energies(a, far_spec, z_xbuf, dfa_noisy_sum, echo_est); /* :498 */
    mu = calc_step_size(a);                                      /* :503 */
    a->tot_count++;                                              /* :506 */
    update_channel(a, far_spec, z_xbuf, dfa_noisy, mu, echo_est); /* :511 */
    sup_gain = calc_suppression_gain(a);                         /* :514 */

    for (i = 0; i < OR_BINS; i++) { /* Wiener gain, :517-615 */
        uint32_t gained;
        int32_t t32;
        int16_t z32, z16, res_diff, q_diff, dq, n1, n2;
        t32 = echo_est[i] - a->echo_filt[i];
        a->echo_filt[i] += (int32_t)(((int64_t)t32 * 50) >> 8);
        z32 = (int16_t)(oracle_norm_w32(a->echo_filt[i]) + 1);
        z16 = (int16_t)(oracle_norm_w16(sup_gain) + 1);
        if (z32 + z16 > 16) {
            gained = (uint32_t)a->echo_filt[i] * (uint16_t)sup_gain;
            res_diff = (int16_t)(14 - RES_CH16 - RES_SUPGAIN + (a->q_clean - z_xbuf));
        } else {
            const int16_t t = (int16_t)(17 - z32 - z16);
            res_diff = (int16_t)(14 + t - RES_CH16 - RES_SUPGAIN + (a->q_clean - z_xbuf));
            if (z32 > t) gained = (uint32_t)a->echo_filt[i] * (uint16_t)(sup_gain >> t);
            else gained = (uint32_t)((a->echo_filt[i] >> t) * sup_gain);
        }
        z16 = (int16_t)oracle_norm_w16(a->near_filt[i]); /* :552-579 */
        dq = (int16_t)(a->q_clean - a->q_clean_old);
        if (z16 < dq && a->near_filt[i]) {
            n1 = (int16_t)(a->near_filt[i] * (1 << z16));
            q_diff = (int16_t)(z16 - dq);
            n2 = (int16_t)(dfa_clean[i] >> -q_diff);
        } else {
            n1 = dq < 0 ? (int16_t)(a->near_filt[i] >> -dq) : (int16_t)(a->near_filt[i] * (1 << dq));
            q_diff = 0;
            n2 = (int16_t)dfa_clean[i];
        }
        t32 = (int32_t)(n2 - n1);
        n2 = (int16_t)(t32 >> 4);
        n2 = (int16_t)(n2 + n1);
        z16 = (int16_t)oracle_norm_w16(n2);
        if ((n2) & (-q_diff > z16)) a->near_filt[i] = 32767; /* bitwise & of value and flag, :572 */
        else a->near_filt[i] = q_diff < 0 ? (int16_t)(n2 * (1 << -q_diff)) : (int16_t)(n2 >> q_diff);

        if (gained == 0) { /* :582-611 */
            hnl[i] = ONE_Q14;
        } else if (a->near_filt[i] == 0) {
            hnl[i] = 0;
        } else {
            uint32_t ratio;
            gained += (uint32_t)(a->near_filt[i] >> 1);
            ratio = div_u32_u16(gained, (uint16_t)a->near_filt[i]);
            t32 = (int32_t)shift_u32(ratio, res_diff);
            if (t32 > ONE_Q14) hnl[i] = 0;
            else if (t32 < 0) hnl[i] = ONE_Q14;
            else { hnl[i] = (int16_t)(ONE_Q14 - (int16_t)t32); if (hnl[i] < 0) hnl[i] = 0; }
        }
        if (hnl[i]) num_pos++;
    }
    if (a->mult == 2) { /* wideband only, :618-648 */
        int32_t avg = 0;
        for (i = 0; i < OR_BINS; i++) hnl[i] = (int16_t)((hnl[i] * hnl[i]) >> 14);
        for (i = 4; i <= 24; i++) avg += hnl[i];
        avg /= 21;
        for (i = 24; i < OR_BINS; i++) if (hnl[i] > (int16_t)avg) hnl[i] = (int16_t)avg;
    }
    for (i = 0; i < OR_BINS; i++) { /* NLP (nlpFlag is 1 forever, aecm_core.cc:399), :651-686 */
        const int16_t nlp_gain = num_pos < 3 ? 0 : ONE_Q14;
        if (hnl[i] > ONE_Q14) hnl[i] = ONE_Q14;
        else if (hnl[i] < NLP_LOW) hnl[i] = 0;
        if (!(hnl[i] == ONE_Q14 && nlp_gain == ONE_Q14)) hnl[i] = (int16_t)((hnl[i] * nlp_gain) >> 14);
        ere[i] = (int16_t)((dre[i] * hnl[i] + 8192) >> 14);
        eim[i] = (int16_t)((dim[i] * hnl[i] + 8192) >> 14);
    }
    if (a->cng_mode == 1) comfort_noise(a, dfa_clean, ere, eim, hnl); /* :702-705 */

    { /* synthesis, :193-246 */
        int16_t packed[130], td[128];
        int scale, c;
        for (i = 0; i <= 64; i++) {
            packed[2 * i] = ere[i];
            packed[2 * i + 1] = neg16(eim[i]);
        }
        scale = oracle_real_fft128_inverse(packed, td);
        c = scale - a->q_clean;
        for (i = 0; i < OR_PART; i++) {
            int32_t t32 = (int16_t)((td[i] * kOracleSqrtHann[i] + 8192) >> 14);
            t32 = shift_s32(t32, c);
            out64[i] = sat16(t32 + a->out_ola[i]);
            t32 = (td[OR_PART + i] * kOracleSqrtHann[OR_PART - i]) >> 14;
            t32 = shift_s32(t32, c);
            a->out_ola[i] = sat16(t32);
        }
    }
    memcpy(a->x_buf, a->x_buf + OR_PART, sizeof(int16_t) * OR_PART); /* :239-245 */
    memcpy(a->d_noisy, a->d_noisy + OR_PART, sizeof(int16_t) * OR_PART);
    if (clean64) memcpy(a->d_clean, a->d_clean + OR_PART, sizeof(int16_t) * OR_PART);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* L2: one 80-sample frame, aecm_core.cc:501-572.  The farBuf stage at :515-516 is an identity  */
/* FIFO because the core's knownDelay is 0 forever (aecm_core.cc:372), so it is not modelled.   */
/* ------------------------------------------------------------------------------------------ */
static int process_frame(OracleAecm *a, const int16_t *far80, const int16_t *near80, const int16_t *clean80,
                         int16_t *out80) {
    fifo_write(&a->frm_far_fifo, a->frm_far, far80, OR_FRAME);
    fifo_write(&a->frm_near_fifo, a->frm_near, near80, OR_FRAME);
    if (clean80) fifo_write(&a->frm_clean_fifo, a->frm_clean, clean80, OR_FRAME);
    while (a->frm_far_fifo.cnt >= OR_PART) {
        int16_t fb[OR_PART], nb[OR_PART], cb[OR_PART], ob[OR_PART];
        fifo_read(&a->frm_far_fifo, a->frm_far, fb, OR_PART);
        fifo_read(&a->frm_near_fifo, a->frm_near, nb, OR_PART);
        if (clean80) fifo_read(&a->frm_clean_fifo, a->frm_clean, cb, OR_PART);
        if (oracle_process_block(a, fb, nb, clean80 ? cb : NULL, ob) == -1) return -1;
        fifo_write(&a->frm_out_fifo, a->frm_out, ob, OR_PART);
    }
    if (a->frm_out_fifo.cnt < OR_FRAME) /* :559-562: rewind, re-emitting old samples */
        fifo_move(&a->frm_out_fifo, a->frm_out_fifo.cnt - OR_FRAME);
    fifo_read(&a->frm_out_fifo, a->frm_out, out80, OR_FRAME);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* L3: public API                                                                             */
/* ------------------------------------------------------------------------------------------ */
size_t oracle_aecm_sizeof(void) { return sizeof(OracleAecm); }

OracleAecm *oracle_aecm_create(void) { /* echo_control_mobile.cc:89-117 */
    OracleAecm *a = (OracleAecm *)calloc(1, sizeof(OracleAecm));
    if (!a) return NULL;
    fifo_init(&a->far_fifo, a->far_ring, OR_FARRING);
    fifo_init(&a->frm_far_fifo, a->frm_far, OR_FRMRING);
    fifo_init(&a->frm_near_fifo, a->frm_near, OR_FRMRING);
    fifo_init(&a->frm_clean_fifo, a->frm_clean, OR_FRMRING);
    fifo_init(&a->frm_out_fifo, a->frm_out, OR_FRMRING);
    return a;
}

void oracle_aecm_free(OracleAecm *a) { free(a); } /* echo_control_mobile.cc:119-140 */

static void init_echo_path_core(OracleAecm *a, const int16_t *path) { /* aecm_core.cc:249-265 */
    int i;
    memcpy(a->ch_stored, path, sizeof(a->ch_stored));
    memcpy(a->ch_adapt16, path, sizeof(a->ch_adapt16));
    for (i = 0; i < OR_BINS; i++) a->ch_adapt32[i] = (int32_t)((uint32_t)(int32_t)path[i] << 16);
    a->mse_adapt_old = 1000;
    a->mse_stored_old = 1000;
    a->mse_threshold = INT32_MAX;
    a->mse_channel_count = 0;
}

static void init_core(OracleAecm *a, int fs) { /* aecm_core.cc:358-473 */
    int32_t t32 = OR_BINS * OR_BINS;
    int16_t t16 = OR_BINS;
    int i;
    a->mult = (int16_t)(fs / 8000);
    fifo_init(&a->frm_far_fifo, a->frm_far, OR_FRMRING);
    fifo_init(&a->frm_near_fifo, a->frm_near, OR_FRMRING);
    fifo_init(&a->frm_clean_fifo, a->frm_clean, OR_FRMRING);
    fifo_init(&a->frm_out_fifo, a->frm_out, OR_FRMRING);
    memset(a->x_buf, 0, sizeof(a->x_buf));
    memset(a->d_clean, 0, sizeof(a->d_clean));
    memset(a->d_noisy, 0, sizeof(a->d_noisy));
    memset(a->out_ola, 0, sizeof(a->out_ola));
    a->seed = 666;
    a->tot_count = 0;
    /* delay estimator init: delay_estimator_wrapper.cc:208-225, 338-355; delay_estimator.cc:330-334, 483-504 */
    memset(a->bin_far, 0, sizeof(a->bin_far));
    a->bin_far_head = 0;
    memset(a->thr_far, 0, sizeof(a->thr_far));
    memset(a->thr_near, 0, sizeof(a->thr_near));
    a->thr_far_init = 0;
    a->thr_near_init = 0;
    init_binary_delay_estimator(a);
    memset(a->far_hist, 0, sizeof(a->far_hist));
    memset(a->far_hist_q, 0, sizeof(a->far_hist_q));
    a->far_hist_pos = OR_HIST;
    a->q_clean = a->q_clean_old = a->q_noisy = a->q_noisy_old = 0;
    memset(a->near_log, 0, sizeof(a->near_log));
    a->far_log = 0;
    memset(a->echo_adapt_log, 0, sizeof(a->echo_adapt_log));
    memset(a->echo_stored_log, 0, sizeof(a->echo_stored_log));
    init_echo_path_core(a, fs == 8000 ? kOracleChannel8k : kOracleChannel16k);
    memset(a->echo_filt, 0, sizeof(a->echo_filt));
    memset(a->near_filt, 0, sizeof(a->near_filt));
    a->noise_est_ctr = 0;
    a->cng_mode = 1;
    memset(a->noise_low_ctr, 0, sizeof(a->noise_low_ctr));
    memset(a->noise_high_ctr, 0, sizeof(a->noise_high_ctr));
    for (i = 0; i < (OR_BINS >> 1) - 1; i++) { /* pink-ish initial noise floor, :428-435 */
        a->noise_est[i] = t32 << 8;
        t16--;
        t32 -= (int32_t)((t16 << 1) + 1);
    }
    for (; i < OR_BINS; i++) a->noise_est[i] = t32 << 8;
    a->far_e_min = 32767;
    a->far_e_max = -32768;
    a->far_e_maxmin = 0;
    a->far_e_vad = FAR_ENERGY_MIN;
    a->far_e_mse = 0;
    a->current_vad = 0;
    a->vad_update_count = 0;
    a->first_vad = 1;
    a->startup_state = 0;
    a->sup_gain = SUPGAIN_DEFAULT;
    a->sup_gain_old = SUPGAIN_DEFAULT;
    a->sup_a = SUPGAIN_A;
    a->sup_d = SUPGAIN_D;
    a->sup_diff_ab = SUPGAIN_A - SUPGAIN_B;
    a->sup_diff_bd = SUPGAIN_B - SUPGAIN_D;
}

int32_t oracle_aecm_set_config(OracleAecm *a, int16_t cng_mode, int16_t echo_mode) { /* echo_control_mobile.cc:410-479 */
    if (!a) return -1;
    if (a->init_flag != INIT_MAGIC) return OR_ERR_UNINITIALIZED;
    if (cng_mode != 0 && cng_mode != 1) return OR_ERR_BAD_PARAMETER;
    a->cng_mode = cng_mode; /* assigned before echoMode is validated, as in the reference */
    if (echo_mode < 0 || echo_mode > 4) return OR_ERR_BAD_PARAMETER;
    a->echo_mode = echo_mode;
    {
        /* modes 0..2 shift each constant right by 3,2,1 BEFORE subtracting; mode 4 shifts left */
        const int rs = echo_mode < 3 ? 3 - echo_mode : 0, ls = echo_mode == 4 ? 1 : 0;
        const int16_t def = (int16_t)((SUPGAIN_DEFAULT >> rs) << ls);
        const int16_t pa = (int16_t)((SUPGAIN_A >> rs) << ls), pb = (int16_t)((SUPGAIN_B >> rs) << ls);
        const int16_t pd = (int16_t)((SUPGAIN_D >> rs) << ls);
        a->sup_gain = def;
        a->sup_gain_old = def;
        a->sup_a = pa;
        a->sup_d = pd;
        a->sup_diff_ab = (int16_t)(pa - pb);
        a->sup_diff_bd = (int16_t)(pb - pd);
    }
    return 0;
}

int32_t oracle_aecm_init(OracleAecm *a, int32_t fs) { /* echo_control_mobile.cc:142-191 */
    if (!a) return -1;
    if (fs != 8000 && fs != 16000) return OR_ERR_BAD_PARAMETER;
    a->fs = fs;
    init_core(a, fs);
    fifo_init(&a->far_fifo, a->far_ring, OR_FARRING);
    a->init_flag = INIT_MAGIC;
    a->delay_change = 1;
    a->sum = 0;
    a->counter = 0;
    a->check_buff_size = 1;
    a->first_val = 0;
    a->ec_startup = 1;
    a->buf_size_start = 0;
    a->check_buf_size_ctr = 0;
    a->filt_delay = 0;
    a->time_for_delay_change = 0;
    a->wrapper_known_delay = 0;
    a->last_delay_diff = 0;
    memset(a->farend_old, 0, sizeof(a->farend_old));
    if (oracle_aecm_set_config(a, 1, 3) == -1) return OR_ERR_UNSPECIFIED;
    return 0;
}

int32_t oracle_aecm_get_buffer_farend_error(OracleAecm *a, const int16_t *farend, size_t n) { /* :195-213 */
    if (!a) return -1;
    if (!farend) return OR_ERR_NULL_POINTER;
    if (a->init_flag != INIT_MAGIC) return OR_ERR_UNINITIALIZED;
    if (n != 80 && n != 160) return OR_ERR_BAD_PARAMETER;
    return 0;
}

static void delay_comp(OracleAecm *a) { /* echo_control_mobile.cc:575-594 */
    const int n_far = a->far_fifo.cnt;
    const int n_snd = a->ms_in_snd_card_buf * 8 * a->mult;
    if (n_snd - n_far > 4 * OR_PART - OR_FRAME * a->mult) {
        int add = (n_snd >> 1) - n_far;
        if (add < OR_FRAME) add = OR_FRAME;
        if (add > 10 * OR_FRAME) add = 10 * OR_FRAME;
        fifo_move(&a->far_fifo, -add);
        a->delay_change = 1;
    }
}

int32_t oracle_aecm_buffer_farend(OracleAecm *a, const int16_t *farend, size_t n) { /* :215-234 */
    const int32_t err = oracle_aecm_get_buffer_farend_error(a, farend, n);
    if (err != 0) return err;
    if (!a->ec_startup) delay_comp(a);
    fifo_write(&a->far_fifo, a->far_ring, farend, (int)n);
    return 0;
}

static void est_buf_delay(OracleAecm *a, int16_t ms) { /* echo_control_mobile.cc:534-573 */
    const int16_t n_far = (int16_t)a->far_fifo.cnt;
    const int16_t n_snd = (int16_t)(ms * 8 * a->mult);
    int16_t delay_new = (int16_t)(n_snd - n_far), diff;
    if (delay_new < OR_FRAME) {
        fifo_move(&a->far_fifo, OR_FRAME);
        delay_new = (int16_t)(delay_new + OR_FRAME);
    }
    {
        const int f = (8 * a->filt_delay + 2 * delay_new) / 10;
        a->filt_delay = (int16_t)(f > 0 ? f : 0);
    }
    diff = (int16_t)(a->filt_delay - a->wrapper_known_delay);
    if (diff > 224) {
        if (a->last_delay_diff < 96) a->time_for_delay_change = 0; else a->time_for_delay_change++;
    } else if (diff < 96 && a->wrapper_known_delay > 0) {
        if (a->last_delay_diff > 224) a->time_for_delay_change = 0; else a->time_for_delay_change++;
    } else {
        a->time_for_delay_change = 0;
    }
    a->last_delay_diff = diff;
    if (a->time_for_delay_change > 25) {
        const int k = (int)a->filt_delay - 160;
        a->wrapper_known_delay = k > 0 ? k : 0; /* never forwarded to the core (:390-391 commented out) */
    }
}

int32_t oracle_aecm_process(OracleAecm *a, const int16_t *near_noisy, const int16_t *near_clean, int16_t *out,
                            size_t n, int16_t ms) { /* echo_control_mobile.cc:236-408 */
    int32_t ret = 0;
    size_t n_frames, n_blocks10ms, i;
    if (!a) return -1;
    if (!near_noisy) return OR_ERR_NULL_POINTER;
    if (!out) return OR_ERR_NULL_POINTER;
    if (a->init_flag != INIT_MAGIC) return OR_ERR_UNINITIALIZED;
    if (n != 80 && n != 160) return OR_ERR_BAD_PARAMETER;
    if (ms < 0) { ms = 0; ret = OR_WARN_BAD_PARAMETER; }
    else if (ms > 500) { ms = 500; ret = OR_WARN_BAD_PARAMETER; }
    ms = (int16_t)(ms + 10);
    a->ms_in_snd_card_buf = ms;
    n_frames = n / OR_FRAME;
    n_blocks10ms = n_frames / (size_t)a->mult;

    if (a->ec_startup) { /* :285-356 */
        const int16_t *src = near_clean ? near_clean : near_noisy;
        int16_t filled;
        if (out != src) memmove(out, src, sizeof(int16_t) * n);
        filled = (int16_t)((int16_t)a->far_fifo.cnt / OR_FRAME);
        if (a->check_buff_size) {
            int dev;
            double tol;
            a->check_buf_size_ctr++;
            if (a->counter == 0) { a->first_val = a->ms_in_snd_card_buf; a->sum = 0; }
            dev = abs(a->first_val - a->ms_in_snd_card_buf);
            tol = 0.2 * a->ms_in_snd_card_buf > 8 ? 0.2 * a->ms_in_snd_card_buf : 8; /* :312-313 */
            if (dev < tol) { a->sum = (int16_t)(a->sum + a->ms_in_snd_card_buf); a->counter++; }
            else a->counter = 0;
            if ((size_t)a->counter * n_blocks10ms >= 6) {
                const int v = (3 * a->sum * a->mult) / (a->counter * 40);
                a->buf_size_start = (int16_t)(v < 50 ? v : 50);
                a->check_buff_size = 0;
            }
            if ((size_t)a->check_buf_size_ctr * n_blocks10ms > 50) {
                const int v = (3 * a->ms_in_snd_card_buf * a->mult) / 40;
                a->buf_size_start = (int16_t)(v < 50 ? v : 50);
                a->check_buff_size = 0;
            }
        }
        if (!a->check_buff_size) {
            if (filled == a->buf_size_start) {
                a->ec_startup = 0;
            } else if (filled > a->buf_size_start) {
                fifo_move(&a->far_fifo, a->far_fifo.cnt - (int)a->buf_size_start * OR_FRAME);
                a->ec_startup = 0;
            }
        }
    } else { /* :358-397 */
        for (i = 0; i < n_frames; i++) {
            int16_t far80[OR_FRAME];
            const int16_t filled = (int16_t)((int16_t)a->far_fifo.cnt / OR_FRAME);
            if (filled > 0) {
                fifo_read(&a->far_fifo, a->far_ring, far80, OR_FRAME);
                memcpy(a->farend_old[i], far80, sizeof(far80));
            } else {
                memcpy(far80, a->farend_old[i], sizeof(far80));
            }
            if ((i == 0 && a->fs == 8000) || (i == 1 && a->fs == 16000)) est_buf_delay(a, a->ms_in_snd_card_buf);
            if (process_frame(a, far80, near_noisy + OR_FRAME * i, near_clean ? near_clean + OR_FRAME * i : NULL,
                              out + OR_FRAME * i) == -1)
                return -1;
        }
    }
    return ret;
}

size_t oracle_aecm_echo_path_size_bytes(void) { return OR_BINS * sizeof(int16_t); } /* :530-532 */

int32_t oracle_aecm_init_echo_path(OracleAecm *a, const void *echo_path, size_t size_bytes) { /* :481-503 */
    if (!a) return -1;
    if (!echo_path) return OR_ERR_NULL_POINTER;
    if (size_bytes != oracle_aecm_echo_path_size_bytes()) return OR_ERR_BAD_PARAMETER;
    if (a->init_flag != INIT_MAGIC) return OR_ERR_UNINITIALIZED;
    init_echo_path_core(a, (const int16_t *)echo_path);
    return 0;
}

int32_t oracle_aecm_get_echo_path(OracleAecm *a, void *echo_path, size_t size_bytes) { /* :505-528 */
    if (!a) return -1;
    if (!echo_path) return OR_ERR_NULL_POINTER;
    if (size_bytes != oracle_aecm_echo_path_size_bytes()) return OR_ERR_BAD_PARAMETER;
    if (a->init_flag != INIT_MAGIC) return OR_ERR_UNINITIALIZED;
    memcpy(echo_path, a->ch_stored, size_bytes);
    return 0;
}

#include "aecm_runner.inc"
