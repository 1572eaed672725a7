// aecm_block.cuh -- L1: one 64-sample block (aecm_core_c.cc:368-711 and everything it calls)
// (part of the single device translation unit aecm_kernels.cu; see that file's header for the mapping)
#ifndef AECM_BLOCK_CUH_
#define AECM_BLOCK_CUH_

#include "aecm_prims.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// TimeToFrequencyDomain (aecm_core_c.cc:261-365) incl. WindowAndFFT (:166-191) for NS signals at once
// (far + near share twiddles, addresses and instruction-level parallelism).  The transform itself is the
// register-resident warp FFT of aecm_fft.cuh.  old64[n]: previous block half in shared memory; new_at(n, k):
// k-th new sample of signal n.  Writes 65 magnitudes per signal, optionally the packed complex bins (sign
// convention after the conjugation of :188-190), the magnitude sums and the dynamic-Q shifts.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ aecm_fft::LaneWords lane_words_of(const Ctx &c) {
    const int4 v = __ldg(&kLaneTables.words[c.lane]);
    return aecm_fft::LaneWords{v.x, v.y, v.z, v.w};
}

template <int NS, class NewAt>
__device__ __forceinline__ void time_to_frequency(const Ctx &c, const int16_t *const (&old64)[NS], NewAt new_at,
                                                  uint16_t *const (&mag_out)[NS], uint32_t *const (&cplx_out)[NS],
                                                  uint32_t (&sum_out)[NS], int (&q_out)[NS]) {
    const int lane = c.lane;
    const int n0 = (int)(__brev((unsigned)lane) >> 27);  // the lane owns time indices n0 + 32k (bit-reversed order)
    const int w0 = hann_q14(n0), w1 = hann_q14(n0 + 32), w2 = hann_q14(64 - n0), w3 = hann_q14(32 - n0);
    int x[NS][4];
#pragma unroll
    for (int n = 0; n < NS; n++) {
        x[n][0] = old64[n][n0];
        x[n][1] = old64[n][n0 + 32];
        x[n][2] = new_at(n, n0);
        x[n][3] = new_at(n, n0 + 32);
        int peak = max(max(abs(x[n][0]), abs(x[n][1])), max(abs(x[n][2]), abs(x[n][3])));
        peak = min(__reduce_max_sync(kFull, peak), 32767);  // signal_processing_library.cc:154-174
        const int q = norm_w16(peak);                       // :288-289
        q_out[n] = q;
        // :177-181; q = NormW16(max|x|) guarantees x << q stays int16, and the Q14 window keeps the product int16
        x[n][0] = ((x[n][0] << q) * w0) >> 14;
        x[n][1] = ((x[n][1] << q) * w1) >> 14;
        x[n][2] = ((x[n][2] << q) * w2) >> 14;
        x[n][3] = ((x[n][3] << q) * w3) >> 14;
    }
    int bre[NS][3], bim[NS][3];
    aecm_fft::forward128<NS>(x, c.sc->fft, twiddle_table(), lane_words_of(c), lane, bre, bim);
#pragma unroll
    for (int n = 0; n < NS; n++) {
        uint32_t sum = 0;
#pragma unroll
        for (int p = 0; p < kBinPasses; p++) {
            const int i = p < 2 ? lane + 32 * p : 64;
            const int re = bre[n][p];
            const int im = (i == 0 || i == 64) ? 0 : s16(-bim[n][p]);  // :188-190, :296-297
            // :305-358.  The reference short-cuts re == 0 / im == 0 to |im| / |re| (as uint16, so |-32768| = 32768);
            // floor(sqrt(v*v)) = |v| exactly, so the general formula (with the AddSatW32 clamp) covers all three cases.
            // Bin 64 (p == 2, im == 0) IS the short cut: |re|, no root needed (only lane 0's value is kept).
            const uint32_t e = (uint32_t)(re * re) + (uint32_t)(im * im);
            const int mag = p < 2 ? sqrt_floor((int)min(e, 0x7fffffffu)) : abs(re);
            if (p < 2 || lane == 0) {
                mag_out[n][i] = (uint16_t)mag;
                if (cplx_out[n]) cplx_out[n][i] = pack16(re, im);
                sum += (uint32_t)mag;
            }
        }
        sum_out[n] = __reduce_add_sync(kFull, sum);
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// binary spectrum of bands 12..43 (delay_estimator_wrapper.cc:92-125): lane l <-> band 12+l <-> bit l
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t binary_spectrum(const uint16_t *spec, int32_t *thr, int q, int &initialized, int lane) {
    const int sp = spec[12 + lane];
    const int s15 = (int)((uint32_t)sp << (15 - q));
    int t = thr[lane];
    if (!initialized) {  // warp-uniform
        const unsigned any = __ballot_sync(kFull, sp > 0);
        if (sp > 0) t = s15 >> 1;
        if (any) initialized = 1;
    }
    mean_estimator_fix(s15, 6, t);
    thr[lane] = t;
    return __ballot_sync(kFull, s15 > t);
}

// WebRtc_ProcessBinarySpectrum (delay_estimator.cc:521-664).  The three scalars of the estimator are passed in registers
// (warp-uniform) and updated in place.  kExt = false is AECM as shipped: robust validation off, lookahead 0
// (aecm_core.cc:218,225), the float statistics that cannot reach the output are not kept.  kExt = true adds
// the near-end lookahead (:536-543), UpdateRobustValidationStatistics (:96-154), HistogramBasedValidation (:178-225) and
// RobustValidation (:245-266) on the instance's AecmDelayExt record in HBM.  The float arithmetic is written with the
// round-to-nearest intrinsics, one rounding per operation in the reference's order, so that no multiply-add is fused:
// the reference library is built for baseline x86-64 (SSE2, no FMA), where every float operation rounds on its own.
template <bool kExt>
__device__ __forceinline__ void process_binary_spectrum(const Ctx &c, uint32_t near_bits, int head, int &minimum_probability,
                                                        int &last_delay_probability, int &last_delay) {
    AecmInstanceState *st = c.st;
    AecmDelayExt *e = c.ext;
    const int lane = c.lane;
    if (kExt) {
        const int near_size = max(e->near_size, 1);
        if (near_size > 1) {  // delay the near-end binary spectrum by `lookahead` blocks, :536-543
            uint32_t prev = 0;
            const bool mine = lane >= 1 && lane < near_size;
            if (mine) prev = e->near_hist[lane - 1];
            __syncwarp();
            if (mine) e->near_hist[lane] = prev;
            if (lane == 0) e->near_hist[0] = near_bits;
            __syncwarp();
            near_bits = e->near_hist[e->lookahead];
        }
    }
    int key = 0x7fffffff, worst = 0, any_far = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int d = lane + 32 * p;
        if (d < AECM_MAX_DELAY) {
            const uint32_t fw = st->bin_far[wrap_add(head, d, AECM_MAX_DELAY)];
            const int fbits = __popc(fw);
            int mean = st->mean_bit_counts[d];
            if (fbits > 0) {  // :550-564
                mean_estimator_fix(__popc(near_bits ^ fw) << 9, 13 - ((3 * fbits) >> 4), mean);
                st->mean_bit_counts[d] = mean;
                any_far = 1;
            }
            key = min(key, (mean << 7) | d);  // value-major, index-minor: first index wins ties (:568-572)
            worst = max(worst, mean);
        }
    }
    key = __reduce_min_sync(kFull, key);
    worst = __reduce_max_sync(kFull, worst);
    any_far = __any_sync(kFull, any_far);  // :623-626
    const int best = min(key >> 7, 32 << 9);
    const int cand = key & 127;
    const int depth = worst - best;
    if (minimum_probability > 8704 && depth > 2816) {  // :593-605
        const int thr = max(best + 1024, 8704);
        if (minimum_probability > thr) minimum_probability = thr;
    }
    last_delay_probability++;  // :608
    bool valid = depth > 1024 && (best < minimum_probability || best < last_delay_probability);
    if (kExt) {
        constexpr float kQ14 = 1.f / 16384.f;
        int hits = e->candidate_hits, last_cand = e->last_candidate, compare = e->compare_delay;
        float ldh = e->last_delay_histogram;
        const int robust = e->robust, allowed = e->allowed_offset;
        __syncwarp();  // the smoothed bit counts written above are read across lanes below; nobody writes the scalars before all have read them
        if (any_far) {  // UpdateRobustValidationStatistics, :96-154
            const float depth_f = __fmul_rn(__int2float_rn(depth), kQ14);
            float decrease = depth_f;
            const int max_hits = cand < last_delay ? 10 : 1000;  // :101-103
            if (cand != last_cand) {  // :108-111
                hits = 0;
                last_cand = cand;
            }
            hits++;
            if (hits < max_hits) decrease = __fmul_rn(__int2float_rn(st->mean_bit_counts[compare] - best), kQ14);  // :132-136
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const int i = lane + 32 * p;
                if (i < AECM_MAX_DELAY) {
                    float h = e->histogram[i];
                    if (i == cand) {  // :119-122
                        h = __fadd_rn(h, depth_f);
                        if (h > 3000.f) h = 3000.f;
                    }
                    const int in_last = i >= last_delay - 2 && i <= last_delay + 1 && i != cand;  // :141-144
                    const int in_cand = i >= cand - 2 && i <= cand + 1;
                    const float dec = __fadd_rn(__fmul_rn(decrease, __int2float_rn(in_last)),
                                                __fmul_rn(depth_f, __int2float_rn(!in_last && !in_cand)));
                    h = __fsub_rn(h, dec);
                    if (h < 0.f) h = 0.f;
                    e->histogram[i] = h;
                }
            }
            __syncwarp();
        }
        const float h_cand = e->histogram[cand], h_cmp = e->histogram[compare];
        if (robust) {  // HistogramBasedValidation :178-225, RobustValidation :245-266
            float fraction = 1.f;
            const int diff = cand - last_delay;
            if (diff > allowed) {
                fraction = __fsub_rn(1.f, __fmul_rn(0.05f, __int2float_rn(diff - allowed)));
                fraction = fraction > 0.5f ? fraction : 0.5f;
            } else if (diff < 0) {
                fraction = __fsub_rn(0.25f, __fmul_rn(0.05f, __int2float_rn(diff)));
                fraction = fraction > 1.f ? 1.f : fraction;
            }
            float thr = __fmul_rn(h_cmp, fraction);
            thr = thr > 1.5f ? thr : 1.5f;
            const bool by_hist = h_cand >= thr && hits > 10;
            bool ok = last_delay < 0 && (valid || by_hist);
            ok |= valid && by_hist;
            ok |= by_hist && h_cand > ldh;
            valid = ok;
        }
        __syncwarp();  // every lane has read the histogram bins it needs
        if (any_far && valid) {  // :643-661
            if (cand != last_delay) {
                ldh = h_cand > 250.f ? 250.f : h_cand;
                if (h_cand < h_cmp && lane == 0) e->histogram[compare] = h_cand;
            }
            compare = cand;
        }
        if (lane == 0) {
            e->candidate_hits = hits;
            e->last_candidate = last_cand;
            e->compare_delay = compare;
            e->last_delay_histogram = ldh;
        }
    }
    if (any_far && valid) {  // :643-661
        last_delay = cand;
        if (best < last_delay_probability) last_delay_probability = best;
    }
}

// ------------------------------------------------------------------------------------------------
// one 64-sample block (aecm_core_c.cc:368-711).  Input samples are read from the frame FIFOs in shared
// memory starting at in_rd / clean_rd; the first out_n samples of the output block are appended to the output
// FIFO at index out_wr (WebRtc_WriteBuffer of aecm_core.cc:554).
// ------------------------------------------------------------------------------------------------
template <bool kExt>
__device__ __forceinline__ void process_block(const Ctx &c, int in_rd, int clean_rd, bool use_clean, int out_wr, int out_n) {
    AecmInstanceState *st = c.st;
    WarpScratch *sc = c.sc;
    AecmScalars &S = st->s;
    const int lane = c.lane;

    // Scalar state discipline: the record's scalars live in shared memory; every lane reads what it needs into
    // registers, all lanes compute the (warp-uniform) updates redundantly, and lane 0 alone writes them back between
    // two __syncwarp()s ("commit").  No lane ever reads a scalar that another lane may be writing.
    int startup_state = S.startup_state;
    const uint32_t tot_count = S.tot_count;
    if (startup_state < 2) startup_state = (tot_count >= kConvLen) + (tot_count >= 2 * kConvLen);  // :420-424

    // ---- analysis: far, near (noisy) and optionally near (clean), :428-464 ----
    uint32_t dfa_noisy_sum;
    int far_q, q_noisy;
    {
        const int16_t *const olds[2] = {st->x_old, st->d_noisy_old};
        uint16_t *const mags[2] = {sc->xfa, sc->dfa_noisy};
        uint32_t *const cplx[2] = {nullptr, sc->dfw};
        uint32_t sums[2];
        int qs[2];
        time_to_frequency<2>(
            c, olds,
            [&](int n, int k) {
                const int idx = wrap_add(in_rd, k, AECM_FRM_RING);
                return (int)(n == 0 ? st->frm_far[idx] : st->frm_near[idx]);
            },
            mags, cplx, sums, qs);
        far_q = qs[0];
        q_noisy = qs[1];
        dfa_noisy_sum = sums[1];
    }
    const int q_noisy_old = S.q_noisy;  // :444-447
    int q_clean, q_clean_old;
    const uint16_t *dfa_clean = sc->dfa_noisy;
    if (!use_clean) {
        q_clean_old = q_noisy_old;
        q_clean = q_noisy;
    } else {
        const int16_t *const olds[1] = {st->d_clean_old};
        uint16_t *const mags[1] = {sc->dfa_clean};
        uint32_t *const cplx[1] = {sc->dfw};  // the Wiener gain is applied to the clean spectrum (:458)
        uint32_t sums[1];
        int qs[1];
        time_to_frequency<1>(
            c, olds, [&](int, int k) { return (int)st->frm_clean[wrap_add(clean_rd, k, AECM_FRM_RING)]; }, mags, cplx,
            sums, qs);
        q_clean_old = S.q_clean;
        q_clean = qs[0];
        dfa_clean = sc->dfa_clean;
    }

    // ---- far magnitude history (aecm_core.cc:125-138): one 132-byte row to HBM ----
    int pos = S.far_hist_pos + 1;
    if (pos >= AECM_MAX_DELAY) pos = 0;
    {
        uint32_t *row = reinterpret_cast<uint32_t *>(c.far_hist() + pos * AECM_HIST_ROW);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(sc->xfa);
        row[lane] = src[lane];
        if (lane == 0) row[32] = (uint32_t)sc->xfa[64] | ((uint32_t)far_q << 16);
    }

    // ---- delay estimation (delay_estimator_wrapper.cc:233-262, 447-476; delay_estimator.cc:369-382) ----
    mbar_wait(&sc->mbar_b, sc->phase);  // everything from here on lives in part B of the record (its copy ran behind the FFTs)
    int thr_far_init = S.thr_far_init, thr_near_init = S.thr_near_init;
    const int old_head = S.bin_far_head;
    const int head = old_head == 0 ? AECM_MAX_DELAY - 1 : old_head - 1;
    const uint32_t far_bits = binary_spectrum(sc->xfa, st->thr_far, far_q, thr_far_init, lane);
    const uint32_t near_bits = binary_spectrum(sc->dfa_noisy, st->thr_near, q_noisy, thr_near_init, lane);
    if (lane == 0) {
        st->bin_far[head] = far_bits;
        c.g()->bin_far[head] = far_bits;  // write-through: one word changes per block, so the ring is not bulk-stored back
    }
    int minimum_probability = S.minimum_probability, last_delay_probability = S.last_delay_probability;
    int last_delay = S.last_delay;
    __syncwarp();
    process_binary_spectrum<kExt>(c, near_bits, head, minimum_probability, last_delay_probability, last_delay);
    const int delay = last_delay == -2 ? 0 : last_delay;  // aecm_core_c.cc:479-483
    // the energy/VAD section below needs these old values; read them before the first commit
    const int lh = (S.log_head + AECM_LOG_RING - 1) & (AECM_LOG_RING - 1);
    int far_e_min = S.far_e_min, far_e_max = S.far_e_max, far_e_maxmin = S.far_e_maxmin, far_e_vad = S.far_e_vad;
    int far_e_mse = S.far_e_mse, vad_update_count = S.vad_update_count, current_vad = S.current_vad;
    int first_vad = S.first_vad;
    __syncwarp();
    if (lane == 0) {  // commit 1
        S.startup_state = startup_state;
        S.q_noisy_old = q_noisy_old;
        S.q_noisy = q_noisy;
        S.q_clean_old = q_clean_old;
        S.q_clean = q_clean;
        S.far_hist_pos = pos;
        S.thr_far_init = thr_far_init;
        S.thr_near_init = thr_near_init;
        S.bin_far_head = head;
        S.minimum_probability = minimum_probability;
        S.last_delay_probability = last_delay_probability;
        S.last_delay = last_delay;
        S.log_head = lh;
    }

    // ---- aligned far spectrum (aecm_core.cc:157-172) ----
    const uint16_t *far_spec = sc->xfa;
    int zx = far_q;
    if (delay != 0) {
        int rp = pos - delay;
        if (rp < 0) rp += AECM_MAX_DELAY;
        const uint32_t *row = reinterpret_cast<const uint32_t *>(c.far_hist() + rp * AECM_HIST_ROW);
        uint32_t *dst = reinterpret_cast<uint32_t *>(sc->far_row);
        dst[lane] = row[lane];
        const uint32_t tail = row[32];
        if (lane == 0) dst[32] = tail;
        zx = (int)(tail >> 16);
        far_spec = sc->far_row;
    }
    __syncwarp();

    // ---- CalcEnergies (aecm_core.cc:644-755) ----
    const int near_log0 = log_energy_q8(dfa_noisy_sum, q_noisy);
    int far_log, echo_adapt_log0, echo_stored_log0;
    {
        uint32_t e_far = 0, e_adapt = 0, e_stored = 0;
#pragma unroll
        for (int p = 0; p < kBinPasses; p++) {
            const int i = lane + 32 * p;
            if (i < AECM_BINS) {
                const int f = far_spec[i];
                const int est = (int)st->ch_stored[i] * f;  // aecm_core.cc:278-279
                sc->echo_est[i] = est;
                e_far += (uint32_t)f;
                e_adapt += (uint32_t)((int)st->ch_adapt16[i] * f);
                e_stored += (uint32_t)est;
            }
        }
        e_far = __reduce_add_sync(kFull, e_far);
        e_adapt = __reduce_add_sync(kFull, e_adapt);
        e_stored = __reduce_add_sync(kFull, e_stored);
        far_log = log_energy_q8(e_far, zx);
        echo_adapt_log0 = log_energy_q8(e_adapt, kResCh16 + zx);
        echo_stored_log0 = log_energy_q8(e_stored, kResCh16 + zx);
    }
    if (far_log > kFarEnergyMin) {  // :692-730
        int inc_max = 4, dec_max = 11, inc_min = 11, dec_min = 3;
        if (startup_state == 0) { inc_max = 2; dec_min = 2; inc_min = 8; }
        far_e_min = asym_filt(far_e_min, far_log, inc_min, dec_min);
        far_e_max = asym_filt(far_e_max, far_log, inc_max, dec_max);
        far_e_maxmin = s16(far_e_max - far_e_min);
        int t16 = s16(2560 - far_e_min);
        t16 = t16 > 0 ? s16((t16 * kFarEnergyVadRegion) >> 9) : 0;
        t16 = s16(t16 + kFarEnergyVadRegion);
        if ((startup_state == 0) | (vad_update_count > 1024)) {
            far_e_vad = s16(far_e_min + t16);
        } else if (far_e_vad > far_log) {
            far_e_vad = s16(far_e_vad + ((far_log + t16 - far_e_vad) >> 6));
            vad_update_count = 0;
        } else {
            vad_update_count = s16(vad_update_count + 1);
        }
        far_e_mse = s16(far_e_vad + (1 << 8));
    }
    if (far_log > far_e_vad) {  // :733-740
        if ((startup_state == 0) | (far_e_maxmin > kFarEnergyDiff)) current_vad = 1;
    } else {
        current_vad = 0;
    }
    if (current_vad && first_vad) {  // :741-754
        first_vad = 0;
        if (echo_adapt_log0 > near_log0) {
#pragma unroll
            for (int p = 0; p < kBinPasses; p++) {
                const int i = lane + 32 * p;
                if (i < AECM_BINS) st->ch_adapt16[i] = (int16_t)(st->ch_adapt16[i] >> 3);
            }
            echo_adapt_log0 = s16(echo_adapt_log0 - (3 << 8));
            first_vad = 1;
        }
    }

    // ---- CalcStepSize (aecm_core.cc:767-794) ----
    int mu = kMuMax;
    if (!current_vad) {
        mu = 0;
    } else if (startup_state > 0) {
        if (far_e_min >= far_e_max) {
            mu = kMuMin;
        } else {
            const int t16 = s16(far_log - far_e_min);
            const int den = far_e_maxmin;
            const int t32 = den != 0 ? (t16 * kMuDiff) / den : 0x7fffffff;
            mu = s16(kMuMin - 1 - s16(t32));
        }
        if (mu < kMuMax) mu = kMuMax;
    }
    __syncwarp();
    if (lane == 0) {  // commit 2
        S.far_log = far_log;
        st->near_log[lh] = (int16_t)near_log0;
        st->echo_adapt_log[lh] = (int16_t)echo_adapt_log0;
        st->echo_stored_log[lh] = (int16_t)echo_stored_log0;
        S.far_e_min = far_e_min;
        S.far_e_max = far_e_max;
        S.far_e_maxmin = far_e_maxmin;
        S.far_e_vad = far_e_vad;
        S.far_e_mse = far_e_mse;
        S.vad_update_count = vad_update_count;
        S.current_vad = current_vad;
        S.first_vad = first_vad;
        S.tot_count = tot_count + 1;  // aecm_core_c.cc:506
    }
    __syncwarp();

    // ---- UpdateChannel (aecm_core.cc:810-986) ----
    const int q_n = q_noisy;
    if (mu) {
#pragma unroll
        for (int p = 0; p < kBinPasses; p++) {
            const int i = lane + 32 * p;
            if (i < AECM_BINS) {
                const int ch32 = st->ch_adapt32[i];
                const uint32_t far = far_spec[i];
                const uint32_t dfa = sc->dfa_noisy[i];
                const int z_ch = norm_u32((uint32_t)ch32);
                const int z_far = norm_u32(far);
                uint32_t prod;
                int sh_chfar;
                if (z_ch + z_far > 31) {  // :836-850
                    prod = (uint32_t)ch32 * far;
                    sh_chfar = 0;
                } else {
                    sh_chfar = 32 - z_ch - z_far;
                    prod = (uint32_t)(sh_chfar >= 32 ? 0 : ch32 >> sh_chfar) * far;
                }
                int z_num = norm_u32(prod);
                const int z_dfa = dfa ? norm_u32(dfa) : 32;
                const int t16 = z_dfa - 2 + q_n - kResCh32 - zx + sh_chfar;
                int xq, dq;
                if (z_num > t16 + 1) {  // :860-867
                    xq = t16;
                    dq = z_dfa - 2;
                } else {
                    xq = z_num - 2;
                    dq = kResCh32 + zx - q_n - sh_chfar + xq;
                }
                prod = shift_u32(prod, xq);
                const uint32_t dfa_al = shift_u32(dfa, dq);
                const int err = (int)dfa_al - (int)prod;
                z_num = norm_w32(err);
                if (err && (far > (uint32_t)(kChannelVad << zx))) {  // :873
                    int upd, sh_num;
                    if (z_num + z_far > 31) {  // :886-902
                        upd = err > 0 ? (int)((uint32_t)err * far) : -(int)((uint32_t)(-err) * far);
                        sh_num = 0;
                    } else {
                        sh_num = 32 - (z_num + z_far);
                        upd = err > 0 ? (err >> sh_num) * (int)far : -((-err >> sh_num) * (int)far);
                    }
                    upd = div_by_bin(upd, i + 1);  // :904 WebRtcSpl_DivW32W16(tmp32no2, i + 1)
                    const int sh2 = sh_num + sh_chfar - xq - mu - ((30 - z_far) << 1);
                    if (norm_w32(upd) < sh2) upd = 0x7fffffff;  // :908-912
                    else upd = shift_s32(upd, sh2);
                    int nch = add_sat32(ch32, upd);
                    if (nch < 0) nch = 0;
                    st->ch_adapt32[i] = nch;
                    st->ch_adapt16[i] = (int16_t)(nch >> 16);
                }
            }
        }
        __syncwarp();
    }
    // scalars of the channel-store decision and of the suppression gain (last written by the previous block's commit 3)
    int mse_channel_count = S.mse_channel_count, mse_threshold = S.mse_threshold;
    int mse_stored_old = S.mse_stored_old, mse_adapt_old = S.mse_adapt_old;
    bool do_store = false;
    if ((startup_state == 0) & (current_vad != 0)) {  // :926-929
        do_store = true;
    } else {
        if (far_log < far_e_mse) mse_channel_count = 0;
        else mse_channel_count = s16(mse_channel_count + 1);
        if (mse_channel_count >= kMinMseCount + 10) {  // :937-983
            int ds = 0, da = 0;
            if (lane < kMinMseCount) {  // the 20 newest log energies (entry lh was committed above)
                const int k = (lh + lane) & (AECM_LOG_RING - 1);
                const int nl = st->near_log[k];
                ds = abs((int)st->echo_stored_log[k] - nl);
                da = abs((int)st->echo_adapt_log[k] - nl);
            }
            const int mse_stored = __reduce_add_sync(kFull, ds);
            const int mse_adapt = __reduce_add_sync(kFull, da);
            if (((mse_stored << kMseResolution) < (kMinMseDiff * mse_adapt)) &
                ((mse_stored_old << kMseResolution) < (kMinMseDiff * mse_adapt_old))) {
                // ResetAdaptiveChannel, aecm_core.cc:308-323
#pragma unroll
                for (int p = 0; p < kBinPasses; p++) {
                    const int i = lane + 32 * p;
                    if (i < AECM_BINS) {
                        const int v = st->ch_stored[i];
                        st->ch_adapt16[i] = (int16_t)v;
                        st->ch_adapt32[i] = (int)((uint32_t)v << 16);
                    }
                }
            } else if (((kMinMseDiff * mse_stored) > (mse_adapt << kMseResolution)) & (mse_adapt < mse_threshold) &
                       (mse_adapt_old < mse_threshold)) {
                do_store = true;
                if (mse_threshold == 0x7fffffff) {
                    mse_threshold = mse_adapt + mse_adapt_old;
                } else {
                    const int scaled = mse_threshold * 5 / 8;
                    mse_threshold += ((mse_adapt - scaled) * 205) >> 8;
                }
            }
            mse_channel_count = 0;
            mse_stored_old = mse_stored;
            mse_adapt_old = mse_adapt;
        }
    }
    if (do_store) {  // StoreAdaptiveChannel, aecm_core.cc:286-306
        if (lane == 0) sc->stored_dirty = 1;
#pragma unroll
        for (int p = 0; p < kBinPasses; p++) {
            const int i = lane + 32 * p;
            if (i < AECM_BINS) {
                const int v = st->ch_adapt16[i];
                st->ch_stored[i] = (int16_t)v;
                sc->echo_est[i] = v * (int)far_spec[i];
            }
        }
    }

    // ---- CalcSuppressionGain (aecm_core.cc:1000-1052) ----
    int sup_gain_old = S.sup_gain_old, sup_gain = S.sup_gain;
    {
        const int sup_a = S.sup_a, sup_d = S.sup_d, sup_diff_ab = S.sup_diff_ab, sup_diff_bd = S.sup_diff_bd;
        int g = kSupGainDefault;
        if (!current_vad) {
            g = 0;
        } else {
            const int t16 = s16(near_log0 - echo_stored_log0);
            const int de = t16 >= 0 ? t16 : s16(-t16);
            if (de < kEnergyDevTol) {
                if (de < kSupGainEpcDt) {
                    const int t32 = sup_diff_ab * de + (kSupGainEpcDt >> 1);
                    g = s16(sup_a - s16(t32 / kSupGainEpcDt));
                } else {
                    const int t32 = sup_diff_bd * (kEnergyDevTol - de) + ((kEnergyDevTol - kSupGainEpcDt) >> 1);
                    g = s16(sup_d + s16(t32 / (kEnergyDevTol - kSupGainEpcDt)));
                }
            } else {
                g = sup_d;
            }
        }
        const int t = g > sup_gain_old ? g : sup_gain_old;
        sup_gain_old = g;
        sup_gain = s16(sup_gain + s16((t - sup_gain) >> 4));
    }
    __syncwarp();
    if (lane == 0) {  // commit 3
        S.mse_channel_count = mse_channel_count;
        S.mse_threshold = mse_threshold;
        S.mse_stored_old = mse_stored_old;
        S.mse_adapt_old = mse_adapt_old;
        S.sup_gain_old = sup_gain_old;
        S.sup_gain = sup_gain;
    }

    // ---- Wiener gain per bin (aecm_core_c.cc:517-615) ----
    int hnl[3];
    int num_pos = 0;
    const int dq_clean = s16(q_clean - q_clean_old);
    const int z16_sup = norm_w16(sup_gain) + 1;
#pragma unroll
    for (int p = 0; p < kBinPasses; p++) {
        const int i = lane + 32 * p;
        hnl[p] = 0;
        if (i < AECM_BINS) {
            int ef = st->echo_filt[i];
            const int diff = sc->echo_est[i] - ef;
            ef += (int)(((long long)diff * 50) >> 8);  // :523-525
            st->echo_filt[i] = ef;
            const int z32 = norm_w32(ef) + 1;
            uint32_t gained;
            int res_diff;
            if (z32 + z16_sup > 16) {  // :529-550
                gained = (uint32_t)ef * (uint32_t)(uint16_t)sup_gain;
                res_diff = 14 - kResCh16 - kResSupGain + (q_clean - zx);
            } else {
                const int t = 17 - z32 - z16_sup;
                res_diff = 14 + t - kResCh16 - kResSupGain + (q_clean - zx);
                if (z32 > t) gained = (uint32_t)ef * (uint32_t)(uint16_t)(sup_gain >> t);
                else gained = (uint32_t)((ef >> t) * sup_gain);
            }
            const int nf = st->near_filt[i];
            const int dfc = dfa_clean[i];
            int z16 = norm_w16(nf);
            int n1, n2, q_diff;
            if (z16 < dq_clean && nf) {  // :556-566
                n1 = nf << z16;  // NormW16 shift: stays int16
                q_diff = z16 - dq_clean;
                n2 = dfc >> -q_diff;  // uint16 >> (>=1): fits int16
            } else {
                n1 = dq_clean < 0 ? nf >> -dq_clean : nf << dq_clean;  // here dq_clean <= NormW16(nf) (or nf == 0)
                q_diff = 0;
                n2 = s16(dfc);
            }
            n2 = s16(((n2 - n1) >> 4) + n1);  // :567-570 (the inner int16 store cannot truncate)
            z16 = norm_w16(n2);
            int nfn;
            if (n2 & (int)(-q_diff > z16)) nfn = 32767;  // :572 (bitwise and of the value with the flag)
            else nfn = q_diff < 0 ? s16(n2 << -q_diff) : n2 >> q_diff;
            st->near_filt[i] = (int16_t)nfn;
            int h;
            if (gained == 0) {  // :582-611
                h = kOneQ14;
            } else if (nfn == 0) {
                h = 0;
            } else {
                gained += (uint32_t)(nfn >> 1);
                const uint32_t ratio = gained / (uint32_t)(uint16_t)nfn;
                const int t32 = (int)shift_u32(ratio, res_diff);
                if (t32 > kOneQ14) h = 0;
                else if (t32 < 0) h = kOneQ14;
                else h = kOneQ14 - t32;  // t32 in [0, 2^14]: neither int16 store truncates, never negative
            }
            hnl[p] = h;
            num_pos += (h != 0);
        }
    }
    num_pos = __reduce_add_sync(kFull, num_pos);
    if (S.mult == 2) {  // wideband only, :618-648 (mult is only written by Init)
        int part = 0;
#pragma unroll
        for (int p = 0; p < kBinPasses; p++) hnl[p] = (hnl[p] * hnl[p]) >> 14;
        if (lane >= 4 && lane <= 24) part = hnl[0];
        const int avg = __reduce_add_sync(kFull, part) / 21;
#pragma unroll
        for (int p = 0; p < kBinPasses; p++) {
            const int i = lane + 32 * p;
            if (i >= 24 && hnl[p] > avg) hnl[p] = avg;
        }
    }
    // ---- NLP and gain application (aecm_core_c.cc:651-686; nlpFlag is 1 forever, aecm_core.cc:399) ----
    int ere[3], eim[3];
    const int nlp_gain = num_pos < 3 ? 0 : kOneQ14;
#pragma unroll
    for (int p = 0; p < kBinPasses; p++) {
        const int i = lane + 32 * p;
        ere[p] = eim[p] = 0;
        if (i < AECM_BINS) {
            int h = hnl[p];
            if (h > kOneQ14) h = kOneQ14;
            else if (h < kNlpLow) h = 0;
            if (!(h == kOneQ14 && nlp_gain == kOneQ14)) h = (h * nlp_gain) >> 14;
            hnl[p] = h;
            const uint32_t d = sc->dfw[i];
            ere[p] = (lo16(d) * h + 8192) >> 14;  // |d| <= 2^15, h <= 2^14: always int16
            eim[p] = (hi16(d) * h + 8192) >> 14;
        }
    }

    // ---- ComfortNoise (aecm_core_c.cc:52-164) ----
    if (S.cng_mode == 1) {  // cng_mode is only written by set_config
        // noise_est_ctr / seed were last written by the previous block's commit 4
        int noise_est_ctr = S.noise_est_ctr;
        const uint32_t seed = S.seed;
        const int sh = 15 - q_clean;
        int track;
        if (noise_est_ctr < 100) { noise_est_ctr++; track = 6; } else { track = 9; }
#pragma unroll
        for (int p = 0; p < kBinPasses; p++) {
            const int i = lane + 32 * p;
            if (i < AECM_BINS) {
                int ne = st->noise_est[i];
                int lo_c = st->noise_low_ctr[i], hi_c = st->noise_high_ctr[i];
                const int lvl = (int)((uint32_t)dfa_clean[i] << sh);
                if (lvl < ne) {  // :86-102
                    lo_c = 0;
                    if (ne < (1 << track)) {
                        hi_c++;
                        if (hi_c >= 5) { ne--; hi_c = 0; }
                    } else {
                        ne -= (ne - lvl) >> track;
                    }
                } else {  // :103-126
                    hi_c = 0;
                    if ((ne >> 19) > 0) { ne >>= 11; ne *= 2049; }
                    else if ((ne >> 11) > 0) { ne *= 2049; ne >>= 11; }
                    else {
                        lo_c++;
                        if (lo_c >= 5) { ne += (ne >> 9) + 1; lo_c = 0; }
                    }
                }
                int t32 = ne >> sh;  // :130-139
                if (t32 > 32767) { t32 = 32767; ne = t32 << sh; }
                st->noise_est[i] = ne;
                st->noise_low_ctr[i] = (uint8_t)lo_c;
                st->noise_high_ctr[i] = (uint8_t)hi_c;
                const int noise = ((kOneQ14 - hnl[p]) * t32) >> 14;  // in [0, 32767]
                if (i > 0) {  // :146-163: bin 0 gets no noise, bin 64 only a real part
                    const uint32_t r = (__ldg(&kLcgJump.mul[i - 1]) * seed + __ldg(&kLcgJump.add[i - 1])) & 0x7fffffffu;
                    const int rnd = (int)(r >> 16);
                    const int idx = (359 * rnd) >> 15;
                    const uint32_t cs = __ldg(&kAecmCosSinQ13[idx]);
                    const int ur = (noise * lo16(cs)) >> 13;  // |cos|, |sin| <= 2^13: always int16
                    const int ui = i < 64 ? (-noise * hi16(cs)) >> 13 : 0;
                    ere[p] = sat16(ere[p] + ur);
                    eim[p] = sat16(eim[p] + ui);
                }
            }
        }
        __syncwarp();
        if (lane == 0) {  // commit 4 (nobody reads these again in this block)
            S.noise_est_ctr = noise_est_ctr;
            S.seed = (kLcgJump.mul[63] * seed + kLcgJump.add[63]) & 0x7fffffffu;  // 64 draws per block, :143
        }
    }

    // ---- InverseFFTAndWindow (aecm_core_c.cc:193-246, real_fft.c:74-102) ----
    int peak = 0;
    {
        const int4 lb = __ldg(&kLaneTables.bins[lane]);  // tile words of bins lane, lane+32 and of their mirrors
        const int words[2] = {lb.x, lb.y}, mirrors[2] = {lb.z, lb.w};
#pragma unroll
        for (int p = 0; p < kBinPasses; p++) {
            const int i = p < 2 ? lane + 32 * p : 64;
            if (p < 2 || lane == 0) {
                const int nim = s16(-eim[p]);  // aecm_core_c.cc:206-214
                sc->fft[p < 2 ? words[p] : aecm_fft::phys(1)] = pack16(ere[p], nim);  // bin 64 sits at position brev7(64) = 1
                peak = max(peak, max(abs(ere[p]), abs(nim)));
                if (p < 2 && i > 0) sc->fft[mirrors[p]] = pack16(ere[p], s16(-nim));  // real_fft.c:87-91
            }
        }
    }
    peak = min(__reduce_max_sync(kFull, peak), 32767);
    __syncwarp();
    int td[4];
    const int scale = aecm_fft::inverse128(sc->fft, twiddle_table(), lane_words_of(c), lane, peak, td);
    const int cshift = scale - q_clean;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int i = lane + 32 * h;
        const int w0 = hann_q14(i), w1 = hann_q14(64 - i);
        int t32 = (td[h] * w0 + 8192) >> 14;  // :219-221 (int16 by the Q14 window)
        t32 = shift_s32(t32, cshift);
        if (i < out_n) st->frm_out[wrap_add(out_wr, i, AECM_FRM_RING)] = (int16_t)sat16(t32 + (int)st->out_ola[i]);  // :224-227
        t32 = (td[2 + h] * w1) >> 14;  // :229-230 (no rounding here)
        t32 = shift_s32(t32, cshift);
        st->out_ola[i] = (int16_t)sat16(t32);
        // slide the analysis buffers, :239-245
        st->x_old[i] = st->frm_far[wrap_add(in_rd, i, AECM_FRM_RING)];
        st->d_noisy_old[i] = st->frm_near[wrap_add(in_rd, i, AECM_FRM_RING)];
        if (use_clean) st->d_clean_old[i] = st->frm_clean[wrap_add(clean_rd, i, AECM_FRM_RING)];
    }
    __syncwarp();
}

}  // namespace

#endif  // AECM_BLOCK_CUH_
