// aecm_init.cuh -- WebRtcAecm_Init / set_config / echo-path import on one instance record (warp-cooperative)
// (part of the single device translation unit aecm_kernels.cu; also compiled by tests/host_emu on the CPU)
#ifndef AECM_INIT_CUH_
#define AECM_INIT_CUH_

#include "aecm_prims.cuh"

namespace {

// WebRtcAecm_Init (echo_control_mobile.cc:142-191) + WebRtcAecm_InitCore (aecm_core.cc:358-473) +
// delay-estimator init (delay_estimator.cc:330-334, 483-504; delay_estimator_wrapper.cc:208-225, 338-355).
// One warp per instance, working directly on the HBM record (far_hist / far_ring are zeroed by memset).
__device__ void init_echo_path_core(AecmInstanceState *g, const int16_t *path, int lane) {  // aecm_core.cc:249-265
    for (int i = lane; i < AECM_BINS; i += 32) {
        const int v = path[i];
        g->ch_stored[i] = (int16_t)v;
        g->ch_adapt16[i] = (int16_t)v;
        g->ch_adapt32[i] = (int)((uint32_t)v << 16);
    }
    if (lane == 0) {
        g->s.mse_adapt_old = 1000;
        g->s.mse_stored_old = 1000;
        g->s.mse_threshold = 0x7fffffff;
        g->s.mse_channel_count = 0;
    }
}

__device__ void apply_config(AecmScalars &S, int cng_mode, int echo_mode) {  // echo_control_mobile.cc:431-476
    const int rs = echo_mode < 3 ? 3 - echo_mode : 0, ls = echo_mode == 4 ? 1 : 0;
    const int def = (kSupGainDefault >> rs) << ls, pa = (kSupGainA >> rs) << ls, pb = (kSupGainB >> rs) << ls,
              pd = (kSupGainD >> rs) << ls;
    S.cng_mode = cng_mode;
    S.echo_mode = echo_mode;
    S.sup_gain = def;
    S.sup_gain_old = def;
    S.sup_a = pa;
    S.sup_d = pd;
    S.sup_diff_ab = pa - pb;
    S.sup_diff_bd = pb - pd;
}

// WebRtcAecm_set_config with its validation (echo_control_mobile.cc:410-430) on an initialised or uninitialised record
__device__ __forceinline__ int set_config_checked(AecmScalars &S, int cng_mode, int echo_mode) {
    if (S.init_flag != AECM_INIT_MAGIC) return AECM_UNINITIALIZED_ERROR;
    if (cng_mode != 0 && cng_mode != 1) return AECM_BAD_PARAMETER_ERROR;
    if (echo_mode < 0 || echo_mode > 4) {
        S.cng_mode = cng_mode;  // the reference assigns cngMode before validating echoMode (:424-428)
        return AECM_BAD_PARAMETER_ERROR;
    }
    apply_config(S, cng_mode, echo_mode);
    return 0;
}

// ---- the delay estimator's optional machinery on the AecmDelayExt record (one warp per instance) ----
// WebRtc_InitBinaryDelayEstimator's share of it (delay_estimator.cc:483-504): statistics and near history are cleared, the
// properties of the estimator object (robust_validation_enabled, allowed_offset, lookahead, near_history_size) stay.
__device__ __forceinline__ void init_delay_ext(AecmDelayExt *e, int lane) {
    for (int k = lane; k < AECM_MAX_DELAY + 4; k += 32) e->histogram[k] = 0.f;
    if (lane <= AECM_MAX_LOOKAHEAD) e->near_hist[lane] = 0;
    if (lane == 0) {
        e->candidate_hits = 0;
        e->last_candidate = -2;
        e->compare_delay = AECM_MAX_DELAY;
        e->last_delay_histogram = 0.f;
    }
}

// AecmB200_SetDelayEstimatorOptions for one instance.  max_lookahead >= 0 is "the estimator is re-created with this
// max_lookahead and initialised" -- WebRtc_CreateDelayEstimator(farend, max_lookahead) (delay_estimator_wrapper.cc:289-336,
// delay_estimator.cc:410-446: lookahead starts at max_lookahead, robust validation off, allowed offset 0) followed by
// WebRtc_InitDelayEstimator (:338-355); then WebRtc_set_lookahead (:384-393), WebRtc_enable_robust_validation (:421-434)
// and WebRtc_set_allowed_offset (:401-409).  Returns 0, or -1 where the reference's setter refuses (lookahead out of range).
__device__ __forceinline__ int set_delay_options(AecmInstanceState *g, AecmDelayExt *e, int robust, int allowed_offset,
                                                 int max_lookahead, int lookahead, int lane) {
    if (max_lookahead >= 0) {
        for (int k = lane; k < 32; k += 32) g->thr_near[k] = 0;  // mean_near_spectrum, near_spectrum_initialized
        for (int k = lane; k <= AECM_MAX_DELAY; k += 32) g->mean_bit_counts[k] = 20 << 9;
        init_delay_ext(e, lane);
        if (lane == 0) {
            g->s.thr_near_init = 0;
            g->s.minimum_probability = 32 << 9;
            g->s.last_delay_probability = 32 << 9;
            g->s.last_delay = -2;
            e->near_size = max_lookahead + 1;
            e->lookahead = max_lookahead;
            e->robust = 0;
            e->allowed_offset = 0;
        }
    }
    __syncwarp();
    const int near_size = max(e->near_size, 1);
    if (lookahead >= 0 && lookahead > near_size - 1) return -1;
    __syncwarp();
    if (lane == 0) {
        if (lookahead >= 0) e->lookahead = lookahead;
        e->robust = robust;
        e->allowed_offset = allowed_offset;
    }
    return 0;
}

// WebRtc_SoftResetDelayEstimatorFarend (delay_estimator.cc:336-367: the binary far-end history moves by delay_shift
// entries, the gap is zero-filled; the bit counts follow, being popcounts of the words) + WebRtc_SoftResetDelayEstimator
// (:506-519: the lookahead moves the other way, clamped to the near history).  Returns the lookahead shift applied.
// The history is a ring here (newest entry at bin_far_head): logical entry i lives at (head + i) mod 100.
__device__ __forceinline__ int soft_reset_delay_estimator(AecmInstanceState *g, AecmDelayExt *e, int delay_shift, int lane) {
    if (delay_shift != 0) {
        const int head = g->s.bin_far_head;
        uint32_t w[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int i = lane + 32 * p, src = i - delay_shift;
            w[p] = (i < AECM_MAX_DELAY && src >= 0 && src < AECM_MAX_DELAY) ? g->bin_far[(head + src) % AECM_MAX_DELAY] : 0u;
        }
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int i = lane + 32 * p;
            if (i < AECM_MAX_DELAY) g->bin_far[(head + i) % AECM_MAX_DELAY] = w[p];
        }
    }
    const int near_size = max(e->near_size, 1);
    const int old = e->lookahead;
    const int now = min(max(old - delay_shift, 0), near_size - 1);
    __syncwarp();
    if (lane == 0) e->lookahead = now;
    return old - now;
}

// WebRtc_last_delay (delay_estimator_wrapper.cc:498-506) and WebRtc_last_delay_quality (delay_estimator.cc:671-688)
__device__ __forceinline__ float last_delay_quality(const AecmInstanceState *g, const AecmDelayExt *e) {
    if (e->robust) return __fdiv_rn(e->histogram[e->compare_delay], 3000.f);
    const float q = __fdiv_rn(__int2float_rn((32 << 9) - g->s.last_delay_probability), __int2float_rn(32 << 9));
    return q < 0.f ? 0.f : q;
}

// the record `g` becomes a freshly initialised instance at sampling rate fs (all 32 lanes of the warp call this)
__device__ __forceinline__ void init_instance(AecmInstanceState *g, int fs, int lane) {
    uint32_t *w = reinterpret_cast<uint32_t *>(g);
    for (int k = lane; k < AECM_STATE_BYTES / 4; k += 32) w[k] = 0;
    __syncwarp();
    // delay_estimator.cc:490-493, the dummy element [100] (used while last_delay == -2 by the robust validation) included
    for (int k = lane; k <= AECM_MAX_DELAY; k += 32) g->mean_bit_counts[k] = 20 << 9;
    // pink-ish initial noise floor, aecm_core.cc:428-435: noiseEst[i] = (65 - i)^2 << 8 for i < 31, then flat
    for (int i = lane; i < AECM_BINS; i += 32) {
        const int k = i < 31 ? i : 31;
        g->noise_est[i] = ((AECM_BINS - k) * (AECM_BINS - k)) << 8;
    }
    init_echo_path_core(g, fs == 8000 ? kAecmChannelInit8k : kAecmChannelInit16k, lane);
    __syncwarp();
    if (lane == 0) {
        AecmScalars &S = g->s;
        S.mult = fs / 8000;
        S.seed = 666;
        S.first_vad = 1;
        S.far_e_min = 32767;
        S.far_e_max = -32768;
        S.far_e_vad = kFarEnergyMin;
        S.far_hist_pos = AECM_MAX_DELAY;
        S.minimum_probability = 32 << 9;
        S.last_delay_probability = 32 << 9;
        S.last_delay = -2;
        S.fs = fs;
        S.init_flag = AECM_INIT_MAGIC;
        S.delay_change = 1;
        S.check_buff_size = 1;
        S.ec_startup = 1;
        apply_config(S, 1, 3);  // default config, echo_control_mobile.cc:183-188
    }
}

}  // namespace

#endif  // AECM_INIT_CUH_
