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

// the record `g` becomes a freshly initialised instance at sampling rate fs (all 32 lanes of the warp call this)
__device__ __forceinline__ void init_instance(AecmInstanceState *g, int fs, int lane) {
    uint32_t *w = reinterpret_cast<uint32_t *>(g);
    for (int k = lane; k < AECM_STATE_BYTES / 4; k += 32) w[k] = 0;
    __syncwarp();
    for (int k = lane; k < AECM_MAX_DELAY; k += 32) g->mean_bit_counts[k] = 20 << 9;  // delay_estimator.cc:490-493
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
