/*
 * aecm_state.h -- per-instance device state of the batched AECM engine (shared by the CUDA kernels
 * and the C host layer; plain C, no CUDA types).
 *
 * HBM layout ("instance-record SoA"): the batch owns three planes, each indexed by instance:
 *
 *   state    [n]            AecmInstanceState   the record a warp stages into shared memory for every
 *                                               launch (16-byte aligned, size multiple of 16 so one
 *                                               warp moves it with 128-bit coalesced accesses)
 *   far_hist [n][100][66]   uint16              magnitude spectra of the last 100 far-end blocks plus
 *                                               their Q (entry 65); a block writes one row and reads
 *                                               one row (reference aecm_core.h:65-69, aecm_core.cc:125-172)
 *   far_ring [n][4000]      int16               the API-level far-end jitter FIFO
 *                                               (reference echo_control_mobile.cc:29,35,99)
 *
 * One warp owns one instance, so "coalesced" means: everything a warp touches for its instance is
 * contiguous.  Fields mirror the reference's AecmCore (aecm_core.h:41-141), BinaryDelayEstimator /
 * Farend (delay_estimator.h:22-63), DelayEstimator / Farend (delay_estimator_wrapper.cc:25-47) and
 * AecMobile (echo_control_mobile.cc:42-79); state that can never influence the output in this
 * reference (farBuf identity FIFO, float histogram of the disabled robust
 * validation, lookahead history) is not stored.
 */
#ifndef AECM_B200_STATE_H_
#define AECM_B200_STATE_H_

#include <stdint.h>

#define AECM_PART_LEN 64       /* aecm_defines.h:19 */
#define AECM_BINS 65           /* aecm_defines.h:22 PART_LEN1 */
#define AECM_FRAME_LEN 80      /* aecm_defines.h:17 */
#define AECM_MAX_DELAY 100     /* aecm_defines.h:26 */
#define AECM_FAR_RING 4000     /* echo_control_mobile.cc:29,35 */
#define AECM_FRM_RING 144      /* aecm_core.cc:184 FRAME_LEN + PART_LEN */
#define AECM_HIST_ROW 66       /* 65 magnitudes + Q */
#define AECM_LOG_RING 32       /* >= MIN_MSE_COUNT (20) newest log-energies are ever read (aecm_core.cc:943-953) */
#define AECM_INIT_MAGIC 42     /* echo_control_mobile.cc:40 kInitCheck */

/* scalar state, all widened to 32 bit; the narrower reference types are re-imposed by casts in the
 * kernel wherever the reference relies on truncation */
typedef struct {
    /* L1 core */
    int32_t mult, cng_mode;
    uint32_t seed, tot_count;
    int32_t first_vad;
    int32_t q_clean, q_clean_old, q_noisy, q_noisy_old;
    int32_t noise_est_ctr;
    int32_t far_log;
    int32_t mse_adapt_old, mse_stored_old, mse_threshold;
    int32_t far_e_min, far_e_max, far_e_maxmin, far_e_vad, far_e_mse;
    int32_t current_vad, vad_update_count;
    int32_t startup_state, mse_channel_count;
    int32_t sup_gain, sup_gain_old, sup_a, sup_d, sup_diff_ab, sup_diff_bd;
    int32_t far_hist_pos;
    int32_t log_head;               /* newest entry of the three log-energy rings */
    /* delay estimator */
    int32_t thr_far_init, thr_near_init;
    int32_t bin_far_head;           /* newest entry of bin_far[] */
    int32_t minimum_probability, last_delay_probability, last_delay;
    /* L2 frame FIFOs: read index and fill count */
    int32_t frm_in_rd, frm_in_cnt;  /* the far and near FIFOs always move together */
    int32_t frm_clean_rd, frm_clean_cnt; /* the clean FIFO only moves on calls that pass a clean plane */
    int32_t frm_out_rd, frm_out_cnt;
    /* L3 API wrapper */
    int32_t fs, init_flag, echo_mode;
    int32_t buf_size_start, counter, sum, first_val, check_buf_size_ctr;
    int32_t ms_in_snd_card_buf;
    int32_t ec_startup, check_buff_size;
    int32_t far_rd, far_cnt;        /* far_ring FIFO */
    int32_t filt_delay, last_delay_diff, time_for_delay_change, delay_change, wrapper_known_delay;
    int32_t pad_[3];
} AecmScalars;

typedef struct {
    /* ---- part A: what a call needs first (scalars, FIFOs, previous block halves); loaded by the first bulk copy */
    AecmScalars s;
    /* L3: last far frames for replay on underrun */
    int16_t farend_old[2][AECM_FRAME_LEN];
    /* L2 frame->block FIFOs */
    int16_t frm_far[AECM_FRM_RING];
    int16_t frm_near[AECM_FRM_RING];
    int16_t frm_out[AECM_FRM_RING];
    /* time domain: previous block halves */
    int16_t x_old[64];
    int16_t d_noisy_old[64];
    /* ---- part B: needed only after the analysis FFTs; its bulk copy is waited for later ---- */
    int16_t out_ola[64];             /* overlap-add tail */
    /* delay estimator: band thresholds (bands 12..43), binary far history ring, smoothed bit counts */
    int32_t thr_far[32];
    int32_t thr_near[32];
    uint32_t bin_far[104];
    int32_t mean_bit_counts[104];
    /* log-energy rings (newest at log_head, age k at (log_head + k) & 31) */
    int16_t near_log[AECM_LOG_RING];
    int16_t echo_adapt_log[AECM_LOG_RING];
    int16_t echo_stored_log[AECM_LOG_RING];
    /* 65-bin adaptive state (padded so every array starts 16-byte aligned) */
    int32_t ch_adapt32[68];
    int32_t echo_filt[68];
    int32_t noise_est[68];
    int16_t ch_stored[72];
    int16_t ch_adapt16[72];
    int16_t near_filt[72];
    uint8_t noise_low_ctr[80];
    uint8_t noise_high_ctr[80];
    /* ---- clean-path extension: only moved when a clean plane is supplied ---- */
    int16_t d_clean_old[64];
    int16_t frm_clean[AECM_FRM_RING];
} AecmInstanceState;

/* the delay estimator's optional machinery -- robust validation (float histogram of candidate delays,
 * delay_estimator.cc:96-266), near-end lookahead (:536-543) and soft reset (:336-367, 506-519) -- which AECM leaves at
 * its Create-time defaults (aecm_core.cc:218, 225).  Kept in a plane of its own, touched only by the launches of a
 * batch on which AecmB200_SetDelayEstimatorOptions / AecmB200_SoftResetDelayEstimator has been called (they run the
 * kExt instantiation of the call kernel, straight on this HBM record); the default path never reads it. */
#define AECM_MAX_LOOKAHEAD 15
typedef struct {
    float histogram[AECM_MAX_DELAY + 4];   /* [100] is the dummy bin used while last_delay == -2 (delay_estimator.cc:459-461) */
    uint32_t near_hist[AECM_MAX_LOOKAHEAD + 1]; /* binary_near_history, newest first */
    /* properties of the estimator OBJECT: survive Init (delay_estimator.cc:428-433, delay_estimator_wrapper.cc:384-434) */
    int32_t robust, allowed_offset, lookahead, near_size; /* near_size 0 (a fresh batch) means 1 = max_lookahead 0 */
    /* robust-validation statistics, reset by Init (delay_estimator.cc:499-503) */
    int32_t candidate_hits, last_candidate, compare_delay;
    float last_delay_histogram;
} AecmDelayExt;

/* bytes of part A (multiple of 16) */
#define AECM_STATE_PART_A ((int)(sizeof(AecmScalars) + 2 * AECM_FRAME_LEN * 2 + 3 * AECM_FRM_RING * 2 + 2 * 64 * 2))
#define AECM_STATE_BYTES ((int)sizeof(AecmInstanceState))
/* bytes a warp has to move when no clean plane is in use */
#define AECM_STATE_BYTES_NOCLEAN ((int)(sizeof(AecmInstanceState) - 64 * 2 - AECM_FRM_RING * 2))

/* error codes, reference echo_control_mobile.h:22-30 */
#define AECM_UNSPECIFIED_ERROR 12000
#define AECM_UNSUPPORTED_FUNCTION_ERROR 12001
#define AECM_UNINITIALIZED_ERROR 12002
#define AECM_NULL_POINTER_ERROR 12003
#define AECM_BAD_PARAMETER_ERROR 12004
#define AECM_BAD_PARAMETER_WARNING 12100

/* ---- launch interface between the C host layer and the CUDA translation unit ---- */
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    AecmInstanceState *state;
    uint16_t *far_hist;
    int16_t *far_ring;
    AecmDelayExt *dly_ext;
    int32_t n_instances;
} AecmBatchPlanes;

typedef struct {
    const int16_t *far;     /* [count][row_stride] or NULL (no BufferFarend in this launch) */
    const int16_t *near_noisy; /* NULL: no Process in this launch */
    const int16_t *near_clean; /* optional */
    int16_t *out;
    const int16_t *ms;      /* [count] per-instance msInSndCardBuf, or NULL -> ms_uniform */
    const uint8_t *far_mask;/* optional [n_calls][mask_stride]: 0 = skip BufferFarend for that call (underrun) */
    int32_t *status;        /* [count] first non-zero return code of the fused calls, or NULL */
    int64_t row_stride;     /* samples between consecutive instances in the planes */
    int32_t first, count;   /* instance range; plane row r belongs to instance first + r */
    int32_t samples_per_call; /* 80 or 160 */
    int32_t n_calls;        /* fused consecutive 10 ms calls */
    int32_t ms_uniform;
    int32_t mask_stride;    /* entries between consecutive calls in far_mask (>= count) */
    const int32_t *index;   /* optional [count]: plane row r belongs to instance index[r] (distinct); `first` unused */
    int32_t ext;            /* != 0: run the instantiation that drives the delay estimator's optional machinery (AecmDelayExt) */
} AecmCallArgs;

/* all return a cudaError_t value (0 = success); launches are asynchronous on `stream` */
int aecm_cuda_launch_init(const AecmBatchPlanes *p, int first, int count, int samp_freq, void *stream);
int aecm_cuda_launch_set_config(const AecmBatchPlanes *p, int first, int count, int cng_mode, int echo_mode,
                                int32_t *status, void *stream);
int aecm_cuda_launch_init_echo_path(const AecmBatchPlanes *p, int first, int count, const int16_t *d_path,
                                    int path_stride, int32_t *status, void *stream);
int aecm_cuda_launch_call(const AecmBatchPlanes *p, const AecmCallArgs *a, void *stream);
/* SURVEY 8-f4 control kernels; robust/allowed_offset/max_lookahead/lookahead as in AecmB200_SetDelayEstimatorOptions */
int aecm_cuda_launch_dly_options(const AecmBatchPlanes *p, int first, int count, int robust, int allowed_offset, int max_lookahead,
                                 int lookahead, int32_t *status, void *stream);
int aecm_cuda_launch_dly_soft_reset(const AecmBatchPlanes *p, int first, int count, int delay_shift, int32_t *applied, void *stream);
int aecm_cuda_launch_dly_query(const AecmBatchPlanes *p, int first, int count, int32_t *last_delay, float *quality, void *stream);
long long aecm_cuda_selftest_sqrt(uint32_t first, uint32_t count);
int aecm_cuda_selftest_fft(int mode, const int16_t *in, int16_t *out, int32_t *scale, int n); /* host arrays */
int aecm_cuda_kernel_launch_count(void); /* kernels launched so far by this translation unit */

#ifdef __cplusplus
}
#endif
#endif /* AECM_B200_STATE_H_ */
