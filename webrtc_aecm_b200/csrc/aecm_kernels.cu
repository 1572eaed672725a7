// aecm_kernels.cu -- the batched AECM engine's device code for sm_100a.
//
// One warp owns one AECM instance.  A launch stages the instance record (aecm_state.h) from HBM into
// shared memory with one bulk async copy per warp (TMA unit, mbarrier-signalled), runs `n_calls` fused 10 ms API calls
// (BufferFarend + Process: far-end jitter FIFO -> 80-sample frames -> 64-sample blocks -> block DSP) and
// writes the record back.  All arithmetic is int16/int32 Q-format and bit-exact with the reference's
// generic C path; every stage cites the reference lines (relative to /root/reference/aecm) it implements.
//
// Lane mapping: 65 frequency bins -> bins lane, lane+32 and (lane 0 only) bin 64; 128 time samples ->
// 4 per lane; the 32 delay-estimator bands -> one per lane (so __ballot_sync *is* the binary spectrum);
// 100 delay candidates -> 4 strided passes.  Warp-uniform state-machine code is computed redundantly by all lanes in
// registers (no broadcasts); the scalars themselves are read from the shared-memory record and written back by lane 0
// only, with __syncwarp() between the last read and the write and after it (compute-sanitizer racecheck: 0 hazards).
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

#include "aecm_fft.cuh"
#include "aecm_state.h"
#include "aecm_tables.h"

namespace {

constexpr unsigned kFull = 0xffffffffu;
// One 28-warp CTA per SM: the warps of an SM start together and drift less through the ~50 KB of straight-line
// code than seven independent 4-warp CTAs do (measured: +26 % throughput), at identical residency.
#ifndef AECM_WARPS_PER_CTA
#define AECM_WARPS_PER_CTA 28
#endif
constexpr int kWarpsPerCta = AECM_WARPS_PER_CTA;
constexpr int kCtasPerSm = kWarpsPerCta >= 28 ? 1 : 28 / kWarpsPerCta;
#ifndef AECM_LOCKSTEP
#define AECM_LOCKSTEP 0
#endif
constexpr bool kLockstep = AECM_LOCKSTEP != 0;  // CTA-wide rendezvous before every block (I-cache locality experiment)  // 28 warps/SM: BASELINE configs[1] (4096 instances) is a single wave on 148 SMs

// ---- algorithm constants, aecm_defines.h:28-85 ----
constexpr int kConvLen = 512;
constexpr int kFarEnergyMin = 1025;
constexpr int kFarEnergyDiff = 929;
constexpr int kEnergyDevTol = 400;
constexpr int kFarEnergyVadRegion = 230;
constexpr int kMuMin = 10, kMuMax = 1, kMuDiff = 9;
constexpr int kMinMseCount = 20, kMinMseDiff = 29, kMseResolution = 5;
constexpr int kResCh16 = 12, kResCh32 = 28, kChannelVad = 16, kResSupGain = 8;
constexpr int kSupGainDefault = 256, kSupGainA = 3072, kSupGainB = 1536, kSupGainD = 256, kSupGainEpcDt = 200;
constexpr int kOneQ14 = 16384, kNlpLow = 3277;

// LCG jump-ahead: after k steps of seed' = (seed*69069 + 1) mod 2^31 the seed is mul[k-1]*seed + add[k-1]
// (signal_processing_library.cc:129-136), so lane-parallel code can draw the k-th number directly.
struct LcgTable {
    uint32_t mul[64], add[64];
};
constexpr LcgTable make_lcg_table() {
    LcgTable t{};
    uint32_t a = 1, c = 0;
    for (int k = 0; k < 64; k++) {
        c = c * 69069u + 1u;
        a = a * 69069u;
        t.mul[k] = a & 0x7fffffffu;
        t.add[k] = c & 0x7fffffffu;
    }
    return t;
}
__device__ const LcgTable kLcgJump = make_lcg_table();

// Per-lane constant words of the FFT tile layouts (aecm_fft.cuh), built at compile time.  All constant tables
// stay in global memory and are read through the read-only path: they are a few KB, hit in L1 after first touch,
// and cost neither shared memory nor a per-launch staging loop.
struct LaneTables {
    int4 words[32];  // aecm_fft::LaneWords {a, b, c, d}
    int4 bins[32];   // aecm_fft::LaneBins {k0, k1, m0, m1}
};
constexpr LaneTables make_lane_tables() {
    LaneTables t{};
    for (int L = 0; L < 32; L++) {
        t.words[L] = int4{aecm_fft::phys(4 * L), aecm_fft::phys((L >> 2) * 16 + (L & 3)),
                          aecm_fft::phys((L >> 4) * 64 + (L & 15)), aecm_fft::phys(L)};
        t.bins[L] = int4{aecm_fft::phys(aecm_fft::brev7c(L)), aecm_fft::phys(aecm_fft::brev7c(L + 32)),
                         aecm_fft::phys(aecm_fft::brev7c((128 - L) & 127)), aecm_fft::phys(aecm_fft::brev7c(96 - L))};
    }
    return t;
}
__device__ const LaneTables kLaneTables = make_lane_tables();

// Exact signed division by d = 2..65 (the NLMS normalisation by bin index + 1, aecm_core.cc:904) without a divide:
// for 0 <= n <= 2^31 and 2^s <= d < 2^(s+1), floor(n / d) = (n * M) >> (32 + s') with M = ceil(2^(32+s') / d) < 2^32,
// s' = s for non powers of two and s - 1 for powers of two (then M = 2^31 exactly).  Truncation toward zero for
// negative dividends is restored by dividing |n|.
struct DivTable {
    uint32_t mul[66];
    int32_t shift[66];
};
constexpr DivTable make_div_table() {
    DivTable t{};
    for (int d = 2; d <= 65; d++) {
        int s = 0;
        while ((2 << s) <= d) s++;          // 2^s <= d < 2^(s+1)
        const bool pow2 = (1 << s) == d;
        const int sp = pow2 ? s - 1 : s;
        const unsigned long long num = 1ull << (32 + sp);
        t.mul[d] = (uint32_t)((num + (unsigned)d - 1) / (unsigned)d);
        t.shift[d] = sp;
    }
    return t;
}
__device__ const DivTable kDivTable = make_div_table();
__device__ __forceinline__ int div_by_bin(int n, int d) {  // d = bin index + 1 in [1, 65]
    if (d == 1) return n;
    const uint32_t a = (uint32_t)abs(n);
    const uint32_t q = __umulhi(a, __ldg(&kDivTable.mul[d])) >> __ldg(&kDivTable.shift[d]);
    return n < 0 ? -(int)q : (int)q;
}

__device__ __forceinline__ const int2 *twiddle_table() { return reinterpret_cast<const int2 *>(kAecmTwiddleCS); }
__device__ __forceinline__ int hann_q14(int i) { return __ldg(&kAecmSqrtHanningQ14[i]); }

struct __align__(16) WarpScratch {
    uint32_t fft[256];     // two 128-word transpose tiles of the register FFT (packed complex int16)
    uint32_t dfw[68];      // near (or clean) spectrum after the sign flip of aecm_core_c.cc:188-190
    int32_t echo_est[68];
    uint16_t xfa[72], dfa_noisy[72], dfa_clean[72], far_row[72];
    uint64_t mbar_a;       // mbarriers of the record's two bulk loads (part A: scalars/FIFOs, part B: the rest)
    uint64_t mbar_b;
};

struct Ctx {
    AecmInstanceState *st;  // shared-memory copy of the record
    WarpScratch *sc;
    uint16_t *far_hist;     // this instance's 100 x 66 rows in HBM
    int16_t *far_ring;      // this instance's 4000-sample jitter FIFO in HBM
    int lane;
};

// wait for a bulk load (phase 0 of its mbarrier); returns at once when the data already arrived
__device__ __forceinline__ void mbar_wait(uint64_t *bar) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_RECORD:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra DONE_RECORD;\n"
        "bra WAIT_RECORD;\n"
        "DONE_RECORD:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar))
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// scalar primitives (spl_inl.h, signal_processing_library.{h,cc})
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int s16(int v) { return (int)(short)v; }  // int16 truncation + sign extension
__device__ __forceinline__ int norm_w32(int a) { return a == 0 ? 0 : __clz(a < 0 ? ~a : a) - 1; }   // spl_inl.h:96
__device__ __forceinline__ int norm_u32(unsigned a) { return a == 0 ? 0 : __clz((int)a); }          // spl_inl.h:101
__device__ __forceinline__ int norm_w16(int a) { return a == 0 ? 0 : __clz(a < 0 ? ~a : a) - 17; } // spl_inl.h:107
__device__ __forceinline__ int add_sat32(int a, int b) {  // spl_inl.h:70-82
    int r;
    asm("add.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }  // spl_inl.h:59-68
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t v) { return ((int)v) >> 16; }

// WEBRTC_SPL_SHIFT_W32 (signal_processing_library.h:130).  The shift count of the NLMS rescaling
// (aecm_core.cc:906-912) can leave [-31, 31]; the reference then executes an x86 SAR/SHL, which uses the
// count modulo 32, so counts are masked here the same way (CUDA's native shifts clamp instead).
__device__ __forceinline__ int shift_s32(int x, int c) {
    return c >= 0 ? (int)((uint32_t)x << (c & 31)) : x >> ((-c) & 31);
}
__device__ __forceinline__ uint32_t shift_u32(uint32_t x, int c) { return c >= 0 ? x << (c & 31) : x >> ((-c) & 31); }

// signal_processing_library.cc:84-105 is floor(sqrt(v)) for v in [0, 2^31).  One MUFU.SQRT
// (relative error <= 2^-23: the estimate is within 0.01 of the root on this range) instead of the IEEE sqrtf sequence
// with its slow-path call; the two integer corrections make the result exact for ANY estimate within +-1 of the root.
// aecm_sqrt_selftest_kernel below sweeps all 2^31 arguments on the device (tests/test_gpu_parity.py).
__device__ __forceinline__ int sqrt_floor(int v) {
    float f;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(__uint2float_rn((uint32_t)v)));
    uint32_t r = (uint32_t)__float2int_rz(f);
    if (r * r > (uint32_t)v) r--;
    if ((r + 1) * (r + 1) <= (uint32_t)v) r++;
    return (int)r;
}

__device__ __forceinline__ void mean_estimator_fix(int v, int factor, int &mean) {  // delay_estimator.cc:690-702
    int d = v - mean;
    d = d < 0 ? -((-d) >> factor) : d >> factor;
    mean += d;
}

__device__ __forceinline__ int log_energy_q8(uint32_t e, int q) {  // aecm_core.cc:612-628
    int r = 7 << 7;
    if (e > 0) {
        const int z = __clz((int)e);
        const int frac = (int)(((e << z) & 0x7fffffffu) >> 23);
        r = s16(r + ((31 - z) << 8) + frac - (q << 8));
    }
    return r;
}

__device__ __forceinline__ int asym_filt(int old_v, int in, int step_pos, int step_neg) {  // aecm_core.cc:588-605
    if (old_v == 32767 || old_v == -32768) return in;
    return old_v > in ? s16(old_v - ((old_v - in) >> step_neg)) : s16(old_v + ((in - old_v) >> step_pos));
}

__device__ __forceinline__ int wrap_add(int base, int k, int cap) {  // (base + k) mod cap for 0 <= base,k < cap
    const int v = base + k;
    return v >= cap ? v - cap : v;
}
__device__ __forceinline__ int brev7(int n) { return (int)(__brev((unsigned)n) >> 25); }

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
        for (int p = 0; p < 3; p++) {
            const int i = p < 2 ? lane + 32 * p : 64;
            const int re = bre[n][p];
            const int im = (i == 0 || i == 64) ? 0 : s16(-bim[n][p]);  // :188-190, :296-297
            // :305-358.  The reference short-cuts re == 0 / im == 0 to |im| / |re| (as uint16, so |-32768| = 32768);
            // floor(sqrt(v*v)) = |v| exactly, so the general formula (with the AddSatW32 clamp) covers all three cases.
            const uint32_t e = (uint32_t)(re * re) + (uint32_t)(im * im);
            const int mag = sqrt_floor((int)min(e, 0x7fffffffu));
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

// WebRtc_ProcessBinarySpectrum (delay_estimator.cc:521-664) with robust validation off, lookahead 0.  The three
// scalars of the estimator are passed in registers (warp-uniform) and updated in place.
__device__ __forceinline__ void process_binary_spectrum(AecmInstanceState *st, uint32_t near_bits, int head, int lane,
                                                        int &minimum_probability, int &last_delay_probability,
                                                        int &last_delay) {
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
    const bool valid = depth > 1024 && (best < minimum_probability || best < last_delay_probability);
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
        uint32_t *row = reinterpret_cast<uint32_t *>(c.far_hist + pos * AECM_HIST_ROW);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(sc->xfa);
        row[lane] = src[lane];
        if (lane == 0) row[32] = (uint32_t)sc->xfa[64] | ((uint32_t)far_q << 16);
    }

    // ---- delay estimation (delay_estimator_wrapper.cc:233-262, 447-476; delay_estimator.cc:369-382) ----
    mbar_wait(&sc->mbar_b);  // everything from here on lives in part B of the record (its copy ran behind the FFTs)
    int thr_far_init = S.thr_far_init, thr_near_init = S.thr_near_init;
    const int old_head = S.bin_far_head;
    const int head = old_head == 0 ? AECM_MAX_DELAY - 1 : old_head - 1;
    const uint32_t far_bits = binary_spectrum(sc->xfa, st->thr_far, far_q, thr_far_init, lane);
    const uint32_t near_bits = binary_spectrum(sc->dfa_noisy, st->thr_near, q_noisy, thr_near_init, lane);
    if (lane == 0) st->bin_far[head] = far_bits;
    int minimum_probability = S.minimum_probability, last_delay_probability = S.last_delay_probability;
    int last_delay = S.last_delay;
    __syncwarp();
    process_binary_spectrum(st, near_bits, head, lane, minimum_probability, last_delay_probability, last_delay);
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
        const uint32_t *row = reinterpret_cast<const uint32_t *>(c.far_hist + rp * AECM_HIST_ROW);
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
        for (int p = 0; p < 3; p++) {
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
            for (int p = 0; p < 3; p++) {
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
        for (int p = 0; p < 3; p++) {
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
                for (int p = 0; p < 3; p++) {
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
#pragma unroll
        for (int p = 0; p < 3; p++) {
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
    for (int p = 0; p < 3; p++) {
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
        for (int p = 0; p < 3; p++) hnl[p] = (hnl[p] * hnl[p]) >> 14;
        if (lane >= 4 && lane <= 24) part = hnl[0];
        const int avg = __reduce_add_sync(kFull, part) / 21;
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int i = lane + 32 * p;
            if (i >= 24 && hnl[p] > avg) hnl[p] = avg;
        }
    }
    // ---- NLP and gain application (aecm_core_c.cc:651-686; nlpFlag is 1 forever, aecm_core.cc:399) ----
    int ere[3], eim[3];
    const int nlp_gain = num_pos < 3 ? 0 : kOneQ14;
#pragma unroll
    for (int p = 0; p < 3; p++) {
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
        for (int p = 0; p < 3; p++) {
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
        for (int p = 0; p < 3; p++) {
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

// ------------------------------------------------------------------------------------------------
// FIFO helpers: a FIFO is (storage, capacity, read index, fill count) -- ring_buffer.c:53-231
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int fifo_move(int &rd, int &cnt, int cap, int k) {  // ring_buffer.c:177-212
    const int room = cap - cnt;
    if (k > cnt) k = cnt;
    if (k < -room) k = -room;
    int r = rd + k;
    if (r >= cap) r -= cap;
    if (r < 0) r += cap;
    rd = r;
    cnt -= k;
    return k;
}

// ------------------------------------------------------------------------------------------------
// sample movers.  All FIFO positions and plane offsets are even in every reachable state (frames of 80, blocks
// of 64, rewinds of 16, DelayComp stuffing of an even count), so 16-bit samples normally move as 32-bit words;
// the 16-bit loops remain as the general fallback.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool word_aligned(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; }

// CTA-wide rendezvous that keeps the warps of a CTA inside the same block of code (instruction-cache locality);
// every warp of the CTA executes it the same number of times per call, whatever its instance is doing.
__device__ __forceinline__ void cta_rendezvous() {
    if (kLockstep) __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// L2: one 80-sample frame (aecm_core.cc:501-572).  The far frame comes from the HBM jitter FIFO (far_idx >= 0:
// index of its first sample) or from the replay buffer farend_old[slot]; near/clean come from the input planes.
// `live` is false for warps that only keep the CTA rendezvous count (no instance, error, or start-up phase).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void process_frame(const Ctx &c, bool live, int far_idx, int slot, const int16_t *near80,
                                              const int16_t *clean80, int16_t *out80) {
    AecmInstanceState *st = c.st;
    AecmScalars &S = st->s;
    const int lane = c.lane;
    // FIFO cursors are read into registers (warp-uniform), updated there, and committed by lane 0
    if (live) {
        const int in_rd = S.frm_in_rd, clean_rd = S.frm_clean_rd;
        int in_cnt = S.frm_in_cnt, clean_cnt = S.frm_clean_cnt;
        // WebRtc_WriteBuffer x2/x3, :520-524 (a full FIFO drops the excess, ring_buffer.c:150-152)
        const int n = min(AECM_FRAME_LEN, AECM_FRM_RING - in_cnt);
        const int wr = wrap_add(in_rd, in_cnt, AECM_FRM_RING);
        if (n == AECM_FRAME_LEN && ((wr | far_idx) & 1) == 0 && word_aligned(near80)) {
            const uint32_t *nw = reinterpret_cast<const uint32_t *>(near80);
            const uint32_t *ring = reinterpret_cast<const uint32_t *>(c.far_ring);
            uint32_t *old = reinterpret_cast<uint32_t *>(st->farend_old[slot]);
            uint32_t *ff = reinterpret_cast<uint32_t *>(st->frm_far), *fn = reinterpret_cast<uint32_t *>(st->frm_near);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int k = lane + 32 * h;
                if (k < AECM_FRAME_LEN / 2) {
                    uint32_t f;
                    if (far_idx >= 0) {
                        f = ring[wrap_add(far_idx >> 1, k, AECM_FAR_RING / 2)];
                        old[k] = f;  // echo_control_mobile.cc:374
                    } else {
                        f = old[k];
                    }
                    const int w = wrap_add(wr >> 1, k, AECM_FRM_RING / 2);
                    ff[w] = f;
                    fn[w] = nw[k];
                }
            }
        } else {
            for (int k = lane; k < n; k += 32) {
                const int16_t f = far_idx >= 0 ? c.far_ring[wrap_add(far_idx, k, AECM_FAR_RING)] : st->farend_old[slot][k];
                if (far_idx >= 0) st->farend_old[slot][k] = f;
                const int w = wrap_add(wr, k, AECM_FRM_RING);
                st->frm_far[w] = f;
                st->frm_near[w] = near80[k];
            }
        }
        in_cnt += n;
        if (clean80) {
            const int nc = min(AECM_FRAME_LEN, AECM_FRM_RING - clean_cnt);
            const int wc = wrap_add(clean_rd, clean_cnt, AECM_FRM_RING);
            for (int k = lane; k < nc; k += 32) st->frm_clean[wrap_add(wc, k, AECM_FRM_RING)] = clean80[k];
            clean_cnt += nc;
        }
        __syncwarp();
        if (lane == 0) {  // commit the input cursors; nothing is kept in registers across a block
            S.frm_in_cnt = in_cnt;
            S.frm_clean_cnt = clean_cnt;
        }
        __syncwarp();
    }
    // :527-555 "while (available >= PART_LEN)": a FIFO below 64 gains 80 samples, so at most two blocks per frame
#pragma unroll 1
    for (int blk = 0; blk < 2; blk++) {
        cta_rendezvous();
        if (live && S.frm_in_cnt >= AECM_PART_LEN) {
            {
                const int o_rd = S.frm_out_rd, o_cnt = S.frm_out_cnt;
                process_block(c, S.frm_in_rd, S.frm_clean_rd, clean80 != nullptr, wrap_add(o_rd, o_cnt, AECM_FRM_RING),
                              min(AECM_PART_LEN, AECM_FRM_RING - o_cnt));
            }
            int i_rd = S.frm_in_rd, i_cnt = S.frm_in_cnt, c_rd = S.frm_clean_rd, c_cnt = S.frm_clean_cnt;
            const int o_cnt = S.frm_out_cnt + min(AECM_PART_LEN, AECM_FRM_RING - S.frm_out_cnt);
            fifo_move(i_rd, i_cnt, AECM_FRM_RING, AECM_PART_LEN);
            if (clean80) fifo_move(c_rd, c_cnt, AECM_FRM_RING, AECM_PART_LEN);
            __syncwarp();
            if (lane == 0) {
                S.frm_in_rd = i_rd; S.frm_in_cnt = i_cnt;
                S.frm_clean_rd = c_rd; S.frm_clean_cnt = c_cnt;
                S.frm_out_cnt = o_cnt;
            }
            __syncwarp();
        }
    }
    if (!live) return;
    int out_rd = S.frm_out_rd, out_cnt = S.frm_out_cnt;
    if (out_cnt < AECM_FRAME_LEN)  // :559-562: rewind, re-emitting already delivered samples
        fifo_move(out_rd, out_cnt, AECM_FRM_RING, out_cnt - AECM_FRAME_LEN);
    {
        const int n = min(AECM_FRAME_LEN, out_cnt);
        if (n == AECM_FRAME_LEN && (out_rd & 1) == 0 && word_aligned(out80)) {
            uint32_t *ow = reinterpret_cast<uint32_t *>(out80);
            const uint32_t *fo = reinterpret_cast<const uint32_t *>(st->frm_out);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int k = lane + 32 * h;
                if (k < AECM_FRAME_LEN / 2) ow[k] = fo[wrap_add(out_rd >> 1, k, AECM_FRM_RING / 2)];
            }
        } else {
            for (int k = lane; k < n; k += 32) out80[k] = st->frm_out[wrap_add(out_rd, k, AECM_FRM_RING)];
        }
        fifo_move(out_rd, out_cnt, AECM_FRM_RING, n);
    }
    __syncwarp();
    if (lane == 0) {  // commit the output cursor
        S.frm_out_rd = out_rd;
        S.frm_out_cnt = out_cnt;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// L3: WebRtcAecm_BufferFarend (echo_control_mobile.cc:215-234) and WebRtcAecm_Process (:236-408)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int api_buffer_farend(const Ctx &c, const int16_t *far, int n) {
    AecmScalars &S = c.st->s;
    if (S.init_flag != AECM_INIT_MAGIC) return AECM_UNINITIALIZED_ERROR;
    int far_rd = S.far_rd, far_cnt = S.far_cnt;
    bool stuffed = false;
    if (!S.ec_startup) {  // WebRtcAecm_DelayComp, :575-594
        const int mult = S.mult;
        const int n_snd = S.ms_in_snd_card_buf * 8 * mult;
        if (n_snd - far_cnt > 4 * AECM_PART_LEN - AECM_FRAME_LEN * mult) {
            int add = max((n_snd >> 1) - far_cnt, AECM_FRAME_LEN);
            add = min(add, 10 * AECM_FRAME_LEN);
            fifo_move(far_rd, far_cnt, AECM_FAR_RING, -add);
            stuffed = true;
        }
    }
    const int w = min(n, AECM_FAR_RING - far_cnt);
    int wr = far_rd + far_cnt;
    if (wr >= AECM_FAR_RING) wr -= AECM_FAR_RING;
    if (((wr | w) & 1) == 0 && word_aligned(far)) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(far);
        uint32_t *ring = reinterpret_cast<uint32_t *>(c.far_ring);
        for (int k = c.lane; k < (w >> 1); k += 32) ring[wrap_add(wr >> 1, k, AECM_FAR_RING / 2)] = src[k];
    } else {
        for (int k = c.lane; k < w; k += 32) c.far_ring[wrap_add(wr, k, AECM_FAR_RING)] = far[k];
    }
    far_cnt += w;
    __syncwarp();
    if (c.lane == 0) {
        S.far_rd = far_rd;
        S.far_cnt = far_cnt;
        if (stuffed) S.delay_change = 1;
    }
    __syncwarp();
    return 0;
}

// WebRtcAecm_EstBufDelay (echo_control_mobile.cc:534-573) on register copies of the wrapper scalars
struct DelayRegs {
    int far_rd, far_cnt, filt_delay, known_delay, last_delay_diff, time_for_delay_change;
};
__device__ __forceinline__ void est_buf_delay(DelayRegs &d, int ms, int mult) {
    const int n_far = s16(d.far_cnt);
    const int n_snd = s16(ms * 8 * mult);
    int delay_new = s16(n_snd - n_far);
    if (delay_new < AECM_FRAME_LEN) {
        fifo_move(d.far_rd, d.far_cnt, AECM_FAR_RING, AECM_FRAME_LEN);
        delay_new = s16(delay_new + AECM_FRAME_LEN);
    }
    d.filt_delay = s16(max(0, (8 * d.filt_delay + 2 * delay_new) / 10));
    const int diff = s16(d.filt_delay - d.known_delay);
    if (diff > 224) {
        if (d.last_delay_diff < 96) d.time_for_delay_change = 0; else d.time_for_delay_change++;
    } else if (diff < 96 && d.known_delay > 0) {
        if (d.last_delay_diff > 224) d.time_for_delay_change = 0; else d.time_for_delay_change++;
    } else {
        d.time_for_delay_change = 0;
    }
    d.last_delay_diff = diff;
    if (d.time_for_delay_change > 25) d.known_delay = max(d.filt_delay - 160, 0);  // never forwarded (:390-391)
}

// `live` = the warp owns an instance and its BufferFarend/Process calls are to be executed; dead warps only keep
// the CTA rendezvous count.
__device__ __forceinline__ int api_process(const Ctx &c, bool live, const int16_t *near_noisy, const int16_t *near_clean,
                                           int16_t *out, int n, int ms) {
    AecmScalars &S = c.st->s;
    const int lane = c.lane;
    const int n_frames = n / AECM_FRAME_LEN;
    int ret = 0;
    bool processing = false;
    if (live) {
        if (S.init_flag != AECM_INIT_MAGIC) {
            ret = AECM_UNINITIALIZED_ERROR;
            live = false;
        }
    }
    if (live) {
        if (ms < 0) { ms = 0; ret = AECM_BAD_PARAMETER_WARNING; }          // :272-278
        else if (ms > 500) { ms = 500; ret = AECM_BAD_PARAMETER_WARNING; }
        ms += 10;
        const int mult = S.mult;
        const int n_blocks10ms = n_frames / mult;
        int ec_startup = S.ec_startup;
        if (ec_startup) {  // :285-356
            const int16_t *src = near_clean ? near_clean : near_noisy;
            if (out != src)
                for (int k = lane; k < n; k += 32) out[k] = src[k];
            int far_rd = S.far_rd, far_cnt = S.far_cnt;
            const int filled = s16(far_cnt) / AECM_FRAME_LEN;
            int check_buff_size = S.check_buff_size, check_buf_size_ctr = S.check_buf_size_ctr;
            int counter = S.counter, first_val = S.first_val, sum = S.sum, buf_size_start = S.buf_size_start;
            if (check_buff_size) {
                check_buf_size_ctr = s16(check_buf_size_ctr + 1);
                if (counter == 0) { first_val = ms; sum = 0; }
                const int dev = abs(first_val - ms);
                const double tol = 0.2 * ms > 8 ? 0.2 * ms : 8;  // :312-313, a double comparison in the reference
                if ((double)dev < tol) { sum = s16(sum + ms); counter = s16(counter + 1); }
                else counter = 0;
                if (counter * n_blocks10ms >= 6) {
                    buf_size_start = s16(min((3 * sum * mult) / (counter * 40), 50));
                    check_buff_size = 0;
                }
                if (check_buf_size_ctr * n_blocks10ms > 50) {
                    buf_size_start = s16(min((3 * ms * mult) / 40, 50));
                    check_buff_size = 0;
                }
            }
            if (!check_buff_size) {
                if (filled == buf_size_start) {
                    ec_startup = 0;
                } else if (filled > buf_size_start) {
                    fifo_move(far_rd, far_cnt, AECM_FAR_RING, far_cnt - buf_size_start * AECM_FRAME_LEN);
                    ec_startup = 0;
                }
            }
            __syncwarp();
            if (lane == 0) {
                S.ms_in_snd_card_buf = ms;
                S.check_buff_size = check_buff_size; S.check_buf_size_ctr = check_buf_size_ctr;
                S.counter = counter; S.first_val = first_val; S.sum = sum; S.buf_size_start = buf_size_start;
                S.ec_startup = ec_startup;
                S.far_rd = far_rd; S.far_cnt = far_cnt;
            }
            __syncwarp();
        } else {
            processing = true;
        }
    }
    for (int i = 0; i < n_frames; i++) {  // :358-397
        int far_idx = -1;
        if (processing) {  // wrapper bookkeeping of this frame on register copies, committed before the frame is processed
            DelayRegs d;
            d.far_rd = S.far_rd; d.far_cnt = S.far_cnt;
            d.filt_delay = S.filt_delay; d.known_delay = S.wrapper_known_delay;
            d.last_delay_diff = S.last_delay_diff; d.time_for_delay_change = S.time_for_delay_change;
            const int fs = S.fs;
            const int filled = s16(d.far_cnt) / AECM_FRAME_LEN;
            if (filled > 0) {  // WebRtc_ReadBuffer of 80 samples
                far_idx = d.far_rd;
                fifo_move(d.far_rd, d.far_cnt, AECM_FAR_RING, AECM_FRAME_LEN);
            }
            if ((i == 0 && fs == 8000) || (i == 1 && fs == 16000)) est_buf_delay(d, ms, S.mult);
            __syncwarp();
            if (lane == 0) {
                S.ms_in_snd_card_buf = ms;
                S.far_rd = d.far_rd; S.far_cnt = d.far_cnt;
                S.filt_delay = d.filt_delay; S.wrapper_known_delay = d.known_delay;
                S.last_delay_diff = d.last_delay_diff; S.time_for_delay_change = d.time_for_delay_change;
            }
            __syncwarp();
        }
        process_frame(c, processing, far_idx, i, near_noisy + AECM_FRAME_LEN * i,
                      near_clean ? near_clean + AECM_FRAME_LEN * i : nullptr, out + AECM_FRAME_LEN * i);
    }
    return ret;
}

// ---- 1-D bulk async copies (TMA unit) between the HBM record and the warp's shared-memory copy ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Issues the two bulk loads of the record (lane 0) and waits for part A; part B is waited for where it is first used.
__device__ __forceinline__ void bulk_load_record(AecmInstanceState *dst_smem, const AecmInstanceState *src_gmem, uint32_t bytes,
                                                 uint64_t *bar_a, uint64_t *bar_b, int lane) {
    if (lane == 0) {
        const uint32_t a = smem_u32(bar_a), b = smem_u32(bar_b);
        const uint32_t bytes_b = bytes - AECM_STATE_PART_A;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"((uint32_t)AECM_STATE_PART_A) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(dst_smem)),
                     "l"(src_gmem), "r"((uint32_t)AECM_STATE_PART_A), "r"(a)
                     : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes_b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(reinterpret_cast<const char *>(dst_smem) + AECM_STATE_PART_A)),
                     "l"(reinterpret_cast<const char *>(src_gmem) + AECM_STATE_PART_A), "r"(bytes_b), "r"(b)
                     : "memory");
    }
    __syncwarp();
    mbar_wait(bar_a);
}
__device__ __forceinline__ void prefetch_line(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bulk_store_record(void *dst_gmem, const void *src_smem, uint32_t bytes, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                     "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the shared-memory source may be released
    }
}

// shared memory: kWarpsPerCta x (AecmInstanceState, WarpScratch)
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm) aecm_call_kernel(AecmBatchPlanes planes, AecmCallArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wbase = smem_raw + (size_t)warp * (sizeof(AecmInstanceState) + sizeof(WarpScratch));
    AecmInstanceState *st = reinterpret_cast<AecmInstanceState *>(wbase);
    WarpScratch *sc = reinterpret_cast<WarpScratch *>(wbase + sizeof(AecmInstanceState));

    const int row = blockIdx.x * kWarpsPerCta + warp;
    const bool live = row < a.count;
    if (!live && !kLockstep) return;
    int inst = a.first + (live ? row : 0);
    if (a.index && live) {  // scattered rows (combined scalar calls): the row names its instance
        inst = __ldg(&a.index[row]);
        if ((unsigned)inst >= (unsigned)planes.n_instances) {
            if (a.status && lane == 0) a.status[row] = AECM_BAD_PARAMETER_ERROR;
            if (!kLockstep) return;
            inst = 0;
        }
    }
    const bool use_clean = a.near_clean != nullptr;
    const uint32_t rec_bytes = use_clean ? AECM_STATE_BYTES : AECM_STATE_BYTES_NOCLEAN;
    AecmInstanceState *g_state = planes.state + inst;
    const int64_t off = (int64_t)row * a.row_stride;
    int16_t *const far_ring = planes.far_ring + (size_t)inst * AECM_FAR_RING;
    if (live) {
        // The audio rows of the first call are needed right after the record: start them on their way from HBM now,
        // one 128-byte line per lane, so that their latency overlaps the record copy instead of following it.
        const int row_bytes = a.samples_per_call * (int)sizeof(int16_t);
        const int at = min(lane * 128, row_bytes - 1);  // lane 0.. cover the row, the last one its final byte
        if (lane * 128 < row_bytes + 128) {
            if (a.far) prefetch_line(reinterpret_cast<const char *>(a.far + off) + at);
            if (a.near_noisy) prefetch_line(reinterpret_cast<const char *>(a.near_noisy + off) + at);
            if (use_clean) prefetch_line(reinterpret_cast<const char *>(a.near_clean + off) + at);
        }
        bulk_load_record(st, g_state, rec_bytes, &sc->mbar_a, &sc->mbar_b, lane);
        // ... and the stretch of the far-end jitter ring this call will read back (echo_control_mobile.cc:369-380:
        // up to 160 samples from the read cursor, which EstBufDelay may first move by one frame either way)
        if (lane < 6) {
            int pos = st->s.far_rd - AECM_FRAME_LEN + lane * 64;
            pos += pos < 0 ? AECM_FAR_RING : 0;
            pos -= pos >= AECM_FAR_RING ? AECM_FAR_RING : 0;
            prefetch_line(far_ring + pos);
        }
        if (use_clean) mbar_wait(&sc->mbar_b);  // the clean-path tail is needed from the first frame on
    }

    Ctx c;
    c.st = st;
    c.sc = sc;
    c.far_hist = planes.far_hist + (size_t)inst * AECM_MAX_DELAY * AECM_HIST_ROW;
    c.far_ring = far_ring;
    c.lane = lane;

    const int ms = (a.ms && live) ? (int)a.ms[row] : a.ms_uniform;
    int status = 0;
    for (int call = 0; call < a.n_calls; call++) {
        const int64_t o = off + (int64_t)call * a.samples_per_call;
        if (live && call + 1 < a.n_calls && lane * 128 < a.samples_per_call * 2) {  // the next call's audio, one line per lane
            const int64_t nx = (o + a.samples_per_call) * (int64_t)sizeof(int16_t) + lane * 128;
            if (a.far) prefetch_line(reinterpret_cast<const char *>(a.far) + nx);
            if (a.near_noisy) prefetch_line(reinterpret_cast<const char *>(a.near_noisy) + nx);
        }
        if (live && a.far && (!a.far_mask || a.far_mask[(size_t)call * a.mask_stride + row])) {
            const int r = api_buffer_farend(c, a.far + o, a.samples_per_call);
            if (status == 0) status = r;
        }
        if (a.near_noisy) {
            const int r = api_process(c, live, a.near_noisy + o, use_clean ? a.near_clean + o : nullptr, a.out + o,
                                      a.samples_per_call, ms);
            if (status == 0) status = r;
        }
    }
    if (!live) return;
    mbar_wait(&sc->mbar_b);  // (a start-up call never touched part B: make sure it has landed before it is stored back)
    bulk_store_record(g_state, st, rec_bytes, lane);
    if (a.status && lane == 0) a.status[row] = status;
}

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

__global__ void aecm_init_kernel(AecmBatchPlanes planes, int first, int count, int fs) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= count) return;
    AecmInstanceState *g = planes.state + first + row;
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

__global__ void aecm_set_config_kernel(AecmBatchPlanes planes, int first, int count, int cng_mode, int echo_mode,
                                       int32_t *status) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= count) return;
    AecmScalars &S = planes.state[first + row].s;
    int r = 0;
    if (S.init_flag != AECM_INIT_MAGIC) {
        r = AECM_UNINITIALIZED_ERROR;
    } else if (cng_mode != 0 && cng_mode != 1) {
        r = AECM_BAD_PARAMETER_ERROR;
    } else if (echo_mode < 0 || echo_mode > 4) {
        S.cng_mode = cng_mode;  // the reference assigns cngMode before validating echoMode (:424-428)
        r = AECM_BAD_PARAMETER_ERROR;
    } else {
        apply_config(S, cng_mode, echo_mode);
    }
    if (status) status[row] = r;
}

__global__ void aecm_init_echo_path_kernel(AecmBatchPlanes planes, int first, int count, const int16_t *path,
                                           int path_stride, int32_t *status) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= count) return;
    AecmInstanceState *g = planes.state + first + row;
    int r = 0;
    if (g->s.init_flag != AECM_INIT_MAGIC) r = AECM_UNINITIALIZED_ERROR;  // echo_control_mobile.cc:496-498
    else init_echo_path_core(g, path + (size_t)row * path_stride, lane);
    if (status && lane == 0) status[row] = r;
}

std::atomic<int> g_launches{0};  // launches may come from several host threads (one per device, or the scalar-API leader)

}  // namespace

// Exhaustive check of sqrt_floor against the reference's bit-by-bit construction (signal_processing_library.cc:84-105:
// sixteen trial-and-subtract steps from bit 15 down).  Counts the arguments in [first, first + count) that disagree.
__global__ void aecm_sqrt_selftest_kernel(uint32_t first, uint32_t count, unsigned long long *bad) {
    unsigned long long mine = 0;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = first + (uint32_t)k;
        uint32_t root = 0, rem = v;
        for (int bit = 15; bit >= 0; bit--) {
            const uint32_t trial = ((root << 1) + (1u << bit)) << bit;  // (root + 2^bit)^2 - root^2
            if (rem >= trial) {
                rem -= trial;
                root |= 1u << bit;
            }
        }
        mine += (uint32_t)sqrt_floor((int)v) != root;
    }
    if (mine) atomicAdd(bad, mine);
}

extern "C" {

int aecm_cuda_kernel_launch_count(void) { return g_launches.load(); }

int aecm_cuda_launch_init(const AecmBatchPlanes *p, int first, int count, int samp_freq, void *stream) {
    if (count <= 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(p->far_hist + (size_t)first * AECM_MAX_DELAY * AECM_HIST_ROW, 0,
                                    (size_t)count * AECM_MAX_DELAY * AECM_HIST_ROW * sizeof(uint16_t), s);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(p->far_ring + (size_t)first * AECM_FAR_RING, 0, (size_t)count * AECM_FAR_RING * sizeof(int16_t), s);
    if (e != cudaSuccess) return (int)e;
    const int threads = 128, warps_per_block = threads / 32;
    aecm_init_kernel<<<(count + warps_per_block - 1) / warps_per_block, threads, 0, s>>>(*p, first, count, samp_freq);
    g_launches++;
    return (int)cudaGetLastError();
}

int aecm_cuda_launch_set_config(const AecmBatchPlanes *p, int first, int count, int cng_mode, int echo_mode,
                                int32_t *status, void *stream) {
    if (count <= 0) return 0;
    aecm_set_config_kernel<<<(count + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, first, count, cng_mode,
                                                                                                echo_mode, status);
    g_launches++;
    return (int)cudaGetLastError();
}

int aecm_cuda_launch_init_echo_path(const AecmBatchPlanes *p, int first, int count, const int16_t *d_path,
                                    int path_stride, int32_t *status, void *stream) {
    if (count <= 0) return 0;
    const int threads = 128, warps_per_block = threads / 32;
    aecm_init_echo_path_kernel<<<(count + warps_per_block - 1) / warps_per_block, threads, 0,
                                 static_cast<cudaStream_t>(stream)>>>(*p, first, count, d_path, path_stride, status);
    g_launches++;
    return (int)cudaGetLastError();
}

int aecm_cuda_launch_call(const AecmBatchPlanes *p, const AecmCallArgs *a, void *stream) {
    if (a->count <= 0 || a->n_calls <= 0) return 0;
    static std::atomic<bool> attr_set[64];  // function attributes are per device
    const size_t smem = (size_t)kWarpsPerCta * (sizeof(AecmInstanceState) + sizeof(WarpScratch));
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(aecm_call_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(aecm_call_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev] = true;
    }
    const int blocks = (a->count + kWarpsPerCta - 1) / kWarpsPerCta;
    aecm_call_kernel<<<blocks, kWarpsPerCta * 32, smem, static_cast<cudaStream_t>(stream)>>>(*p, *a);
    g_launches++;
    return (int)cudaGetLastError();
}

// diagnostics: number of arguments in [first, first + count) for which the device sqrt_floor differs from the
// reference construction; negative = CUDA error
long long aecm_cuda_selftest_sqrt(uint32_t first, uint32_t count) {
    unsigned long long *d_bad = nullptr, bad = 0;
    if (cudaMalloc(&d_bad, sizeof(bad)) != cudaSuccess) return -1;
    cudaMemset(d_bad, 0, sizeof(bad));
    aecm_sqrt_selftest_kernel<<<148 * 8, 256>>>(first, count, d_bad);
    cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    return e == cudaSuccess ? (long long)bad : -(long long)e;
}

}  // extern "C"
