/*
 * aecm_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement (plain C99) of the AECM hot path of
 * cpuimage/WebRTC_AECM, used only as the checker by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py (plus the developer utilities under tools/ that generate the
 * golden fixtures and coverage reports).  Nothing in the product (webrtc_aecm_b200/, include/) may include, link or
 * call this file (tests/test_abi.py checks that).
 *
 * Parity status: PINNED.  The reference ships no golden vectors or tests, so the
 * oracle is pinned against the reference itself: oracle/Makefile compiles the unmodified reference
 * sources into oracle/_ref/libaecm_ref.so, and tests/test_oracle.py requires bit-identical
 * int16 output (and identical return codes) between this restatement and that library on every
 * stream family of SURVEY.md section 8(c); tests/golden/ holds vectors generated from the reference
 * library (script committed) for the GPU box, where /root/reference does not exist.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference/aecm).
 */
#ifndef AECM_ORACLE_H_
#define AECM_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    OR_PART = 64,      /* aecm_defines.h:19  PART_LEN   */
    OR_BINS = 65,      /* aecm_defines.h:22  PART_LEN1  */
    OR_FRAME = 80,     /* aecm_defines.h:17  FRAME_LEN  */
    OR_HIST = 100,     /* aecm_defines.h:26  MAX_DELAY  */
    OR_ELOG = 64,      /* aecm_defines.h:33  MAX_BUF_LEN */
    OR_FARRING = 4000, /* echo_control_mobile.cc:29,35  kBufSizeSamp */
    OR_FRMRING = 144,  /* aecm_core.cc:184  FRAME_LEN + PART_LEN */
    OR_MAX_LOOKAHEAD = 15 /* largest max_lookahead the checkers (and the product) accept; AECM itself uses 0 (aecm_core.cc:218) */
};

/* error codes, echo_control_mobile.h:22-30 */
enum {
    OR_ERR_UNSPECIFIED = 12000,
    OR_ERR_UNSUPPORTED = 12001,
    OR_ERR_UNINITIALIZED = 12002,
    OR_ERR_NULL_POINTER = 12003,
    OR_ERR_BAD_PARAMETER = 12004,
    OR_WARN_BAD_PARAMETER = 12100
};

/* A FIFO with a movable read pointer; models ring_buffer.c:53-231 as (read index, fill count). */
typedef struct {
    int32_t cap, rd, cnt;
} OrFifo;

typedef struct {
    /* ------- L3: API wrapper state (echo_control_mobile.cc:42-79) ------- */
    int32_t fs;
    int16_t init_flag;
    int16_t buf_size_start;
    int16_t counter, sum, first_val, check_buf_size_ctr;
    int16_t ms_in_snd_card_buf;
    int16_t filt_delay;
    int16_t last_delay_diff;
    int16_t echo_mode;
    int32_t time_for_delay_change;
    int32_t ec_startup;
    int32_t check_buff_size;
    int32_t delay_change;
    int32_t wrapper_known_delay;
    int16_t farend_old[2][OR_FRAME];
    OrFifo far_fifo;
    int16_t far_ring[OR_FARRING];

    /* ------- L2: frame -> block re-chunking (aecm_core.cc:183-209, 501-572) ------- */
    OrFifo frm_far_fifo, frm_near_fifo, frm_clean_fifo, frm_out_fifo;
    int16_t frm_far[OR_FRMRING], frm_near[OR_FRMRING], frm_clean[OR_FRMRING], frm_out[OR_FRMRING];

    /* ------- L1: block core (aecm_core.h:41-141) ------- */
    int16_t mult;
    int16_t cng_mode;
    uint32_t seed;
    uint32_t tot_count;
    int32_t first_vad;
    int16_t q_clean, q_clean_old, q_noisy, q_noisy_old;
    int16_t x_buf[2 * OR_PART], d_noisy[2 * OR_PART], d_clean[2 * OR_PART], out_ola[OR_PART];
    int16_t ch_stored[OR_BINS], ch_adapt16[OR_BINS];
    int32_t ch_adapt32[OR_BINS];
    int32_t echo_filt[OR_BINS];
    int16_t near_filt[OR_BINS];
    int32_t noise_est[OR_BINS];
    int32_t noise_low_ctr[OR_BINS], noise_high_ctr[OR_BINS];
    int16_t noise_est_ctr;
    int16_t near_log[OR_ELOG], echo_adapt_log[OR_ELOG], echo_stored_log[OR_ELOG];
    int16_t far_log;
    int32_t mse_adapt_old, mse_stored_old, mse_threshold;
    int16_t far_e_min, far_e_max, far_e_maxmin, far_e_vad, far_e_mse;
    int32_t current_vad;
    int16_t vad_update_count;
    int16_t startup_state;
    int16_t mse_channel_count;
    int16_t sup_gain, sup_gain_old;
    int16_t sup_a, sup_d, sup_diff_ab, sup_diff_bd;
    /* far magnitude history, aecm_core.h:65-69 */
    uint16_t far_hist[OR_HIST * OR_BINS];
    int32_t far_hist_pos;
    int32_t far_hist_q[OR_HIST];
    /* binary-spectrum delay estimator (delay_estimator.h:22-63, delay_estimator_wrapper.cc:25-47).
     * The far binary history is kept as a ring (newest entry at bin_far_head) instead of the
     * reference's memmove shift register; bit counts are derived by popcount. */
    int32_t thr_far[OR_BINS], thr_near[OR_BINS];
    int32_t thr_far_init, thr_near_init;
    uint32_t bin_far[OR_HIST];
    int32_t bin_far_head;
    int32_t mean_bit_counts[OR_HIST + 1]; /* + the dummy element used while last_delay == -2 (delay_estimator.cc:459-461) */
    int32_t minimum_probability, last_delay_probability, last_delay;
    /* robust validation, lookahead, soft reset (delay_estimator.h:34-62).  AECM leaves all of it at the
     * Create-time defaults (robust validation off, aecm_core.cc:225; max_lookahead 0, :218); the statistics are updated
     * regardless, as in the reference (delay_estimator.cc:628-633). */
    int32_t robust_enabled, allowed_offset;
    int32_t near_history_size, lookahead;
    uint32_t bin_near_hist[OR_MAX_LOOKAHEAD + 1];
    float histogram[OR_HIST + 1];
    float last_delay_histogram;
    int32_t candidate_hits, last_candidate_delay, compare_delay;
} OracleAecm;

/* ---- the ten API entry points (echo_control_mobile.h:46-202), one instance per call ---- */
OracleAecm *oracle_aecm_create(void);
void oracle_aecm_free(OracleAecm *a);
int32_t oracle_aecm_init(OracleAecm *a, int32_t samp_freq);
int32_t oracle_aecm_set_config(OracleAecm *a, int16_t cng_mode, int16_t echo_mode);
int32_t oracle_aecm_get_buffer_farend_error(OracleAecm *a, const int16_t *farend, size_t n);
int32_t oracle_aecm_buffer_farend(OracleAecm *a, const int16_t *farend, size_t n);
int32_t oracle_aecm_process(OracleAecm *a, const int16_t *near_noisy, const int16_t *near_clean,
                            int16_t *out, size_t n, int16_t ms_in_snd_card_buf);
int32_t oracle_aecm_init_echo_path(OracleAecm *a, const void *echo_path, size_t size_bytes);
int32_t oracle_aecm_get_echo_path(OracleAecm *a, void *echo_path, size_t size_bytes);
size_t oracle_aecm_echo_path_size_bytes(void);
/* SURVEY 8-f4, the delay estimator's knobs as AECM's estimator object would take them:
 *   robust      WebRtc_enable_robust_validation (delay_estimator_wrapper.cc:421-434), 0/1
 *   allowed_offset  WebRtc_set_allowed_offset (:401-409), >= 0
 *   max_lookahead   >= 0: the estimator is re-created with this max_lookahead (WebRtc_CreateDelayEstimator, :289-336;
 *                   lookahead starts at max_lookahead, delay_estimator.cc:433) and initialised; < 0: left as it is
 *   lookahead       WebRtc_set_lookahead (:384-393) when >= 0
 * Returns 0, or -1 where the reference call returns -1. */
int32_t oracle_aecm_set_delay_estimator_options(OracleAecm *a, int robust, int allowed_offset, int max_lookahead, int lookahead);
/* WebRtc_SoftResetDelayEstimatorFarend + WebRtc_SoftResetDelayEstimator (delay_estimator_wrapper.cc:227-231, 357-361;
 * delay_estimator.cc:336-367, 506-519); returns what the latter returns (the lookahead shift actually applied) */
int32_t oracle_aecm_soft_reset_delay_estimator(OracleAecm *a, int delay_shift);
int32_t oracle_aecm_last_delay(const OracleAecm *a);             /* WebRtc_last_delay (:498-506) */
float oracle_aecm_last_delay_quality(const OracleAecm *a);       /* WebRtc_last_delay_quality (:508-512) */
size_t oracle_aecm_sizeof(void);

/* ---- stage-level entry points for known-answer tests ---- */
/* 128 real samples -> bins 0..64 of the reference's forward transform, interleaved re,im (130 values);
 * real_fft.c:47-72 (before the sign flip of aecm_core_c.cc:188-190). */
void oracle_real_fft128_forward(const int16_t in[128], int16_t out[130]);
/* 65 complex bins (interleaved, 130 values) -> 128 real samples; returns the block scale;
 * real_fft.c:74-102. */
int oracle_real_fft128_inverse(const int16_t in[130], int16_t out[128]);
int32_t oracle_sqrt_floor(int32_t v);             /* signal_processing_library.cc:84-105 */
void oracle_rand_u_array(int16_t *v, int n, uint32_t *seed); /* signal_processing_library.cc:126-147 */
int oracle_norm_w16(int16_t a);                   /* spl_inl.h:107-111 */
int oracle_norm_w32(int32_t a);                   /* spl_inl.h:96-98 */
int oracle_norm_u32(uint32_t a);                  /* spl_inl.h:101-103 */
unsigned oracle_debug_wide_shift_events(void);    /* times aecm_core.cc:911 saw a shift count beyond +-31 (not thread safe) */
/* one block of the core, aecm_core_c.cc:368-711; instance must be initialised */
int oracle_process_block(OracleAecm *a, const int16_t *far64, const int16_t *near64,
                         const int16_t *clean64, int16_t *out64);

/* ---- multi-threaded stream runners (CPU baseline + fixtures); see aecm_runner.c ---- */
typedef struct {
    int32_t n_instances;
    int32_t fs;              /* 8000 / 16000 */
    int32_t samples_per_call;/* 80 / 160 */
    int32_t n_calls;
    int32_t cng_mode, echo_mode;
    int32_t use_clean;       /* pass the clean plane */
    int32_t n_threads;
    int32_t skip_far_every;  /* >0: skip BufferFarend on calls where (call % skip)==skip-1 (underrun) */
    int32_t ms_per_call;     /* !=0: ms is [n_instances][n_calls], one msInSndCardBuf per call */
    int32_t extra_far_every; /* >0: BufferFarend is called TWICE (same frame) where (call % k)==k-1 (overrun) */
    /* SURVEY 8-f4 (all zero = the reference's own wiring).  dly_options != 0: apply the four values below after
     * Init/set_config through oracle_aecm_set_delay_estimator_options / the reference shim (max_lookahead and lookahead
     * are passed as given, -1 = leave alone). */
    int32_t dly_options;
    int32_t dly_robust, dly_allowed_offset, dly_max_lookahead, dly_lookahead;
    int32_t soft_reset_every; /* >0: soft-reset the delay estimator by soft_reset_shift before the calls with (call % k)==k-1,
                                 the sign of the shift alternating from one reset to the next */
    int32_t soft_reset_shift;
} OracleRunConfig;
/* planes are [n_instances][n_calls*samples_per_call] int16; ms is [n_instances] (or [n_instances][n_calls] with
 * ms_per_call) or NULL (=40);
 * ret_or is [n_instances] OR of all return codes. impl: 0 = this oracle, 1 = reference library at
 * ref_lib_path (dlopen, RTLD_LOCAL). Returns 0, or <0 on setup failure. */
int oracle_run_streams(const OracleRunConfig *cfg, int impl, const char *ref_lib_path,
                       const int16_t *far, const int16_t *near_noisy, const int16_t *near_clean,
                       const int16_t *ms, int16_t *out, int32_t *ret_or, double *seconds);

/* the same, additionally recording the delay estimate after every call (WebRtc_last_delay) in delay_trace
 * [n_instances][n_calls] when it is not NULL.  With impl 1 the delay-estimator options, soft resets and the trace need
 * the reference SHIM library (oracle/_ref/libaecm_ref_shim.so) as ref_lib_path; -4 if its helpers are missing. */
int oracle_run_streams_ex(const OracleRunConfig *cfg, int impl, const char *ref_lib_path,
                          const int16_t *far, const int16_t *near_noisy, const int16_t *near_clean,
                          const int16_t *ms, int16_t *out, int32_t *ret_or, int32_t *delay_trace, double *seconds);

/* persistent pool (state survives across runs): warm up once, then time steady-state calls */
typedef struct OraclePool OraclePool;
OraclePool *oracle_pool_create(int impl, const char *ref_lib_path, int n, int fs, int cng_mode, int echo_mode,
                               int threads);
/* runs n_calls x (BufferFarend + Process) on every instance; returns the OR of all return codes */
int oracle_pool_run(OraclePool *p, const int16_t *far, const int16_t *near_noisy, const int16_t *near_clean,
                    int16_t *out, int64_t row_stride, int spc, int n_calls, int ms, double *seconds);
/* throughput variant: tile_rows distinct streams in the planes, instance i is fed row i % tile_rows, output dropped */
int oracle_pool_run_tiled(OraclePool *p, const int16_t *far, const int16_t *near_noisy, int tile_rows, int64_t row_stride,
                          int spc, int n_calls, int ms, double *seconds);
void oracle_pool_free(OraclePool *p);

#ifdef __cplusplus
}
#endif
#endif /* AECM_ORACLE_H_ */
