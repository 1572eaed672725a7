// aecm_prims.cuh -- launch shape, algorithm constants, compile-time tables, per-warp scratch, scalar primitives
// (part of the single device translation unit aecm_kernels.cu; see that file's header for the mapping)
#ifndef AECM_PRIMS_CUH_
#define AECM_PRIMS_CUH_

#ifndef AECM_HOST_EMU  // tests/host_emu supplies the CUDA vocabulary itself
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "aecm_fft.cuh"
#include "aecm_state.h"
#include "aecm_tables.h"

namespace {

constexpr unsigned kFull = 0xffffffffu;
// One 28-warp CTA per SM: the warps of an SM start together and drift less through the ~50 KB of straight-line
// code than seven independent 4-warp CTAs do (measured: +26 % throughput), at identical residency.  The kernel is
// persistent: at most one CTA per SM is launched and every warp walks the rows r, r + (warps in the grid), ...
// 28 warps/SM: BASELINE configs[1] (4096 instances) is a single pass on 148 SMs.
#ifndef AECM_WARPS_PER_CTA
#define AECM_WARPS_PER_CTA 28
#endif
constexpr int kWarpsPerCta = AECM_WARPS_PER_CTA;
constexpr int kCtasPerSm = kWarpsPerCta >= 28 ? 1 : 28 / kWarpsPerCta;
// CTA-wide rendezvous between two rows of the persistent loop: re-aligns the warps of an SM on the instruction stream
// once per row.  The call path is ~110 KB of straight-line SASS against a 32 KB L1.5 instruction cache, so the warps
// of an SM have to walk it together; left to run free (0) they drift apart and starve on instruction fetch (measured at
// 262 144 instances: stall_no_inst 35 % of the samples, 2.34e8 blocks/s; with the rendezvous 2.76e8).  DESIGN.md 4.1.
#ifndef AECM_ROW_RESYNC
#define AECM_ROW_RESYNC 1
#endif
constexpr bool kRowResync = AECM_ROW_RESYNC != 0;
// Finer rendezvous points inside a row: 1 = also before every 80-sample frame, 2 = also before every (potential) 64-sample
// block.  Every warp of the CTA executes the same number of them per call whatever its instance is doing (start-up,
// un-initialised, no row at all): see cta_rendezvous_* in aecm_frame_api.cuh.  Measured, DESIGN.md 4.1.
#ifndef AECM_SYNC_POINTS
#define AECM_SYNC_POINTS 0
#endif
constexpr int kSyncPoints = kRowResync ? AECM_SYNC_POINTS : 0;
// MEASUREMENT ONLY (tools/build_variants.sh): 2 drops the one-live-lane pass for bin 64 from every 65-bin loop.  The output
// is then WRONG; the build exists to bound what removing that pass could buy.  The product is built with 3.
#ifndef AECM_BIN_PASSES
#define AECM_BIN_PASSES 3
#endif
constexpr int kBinPasses = AECM_BIN_PASSES;

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
    const uint32_t a = n < 0 ? 0u - (uint32_t)n : (uint32_t)n;  // |n| as unsigned, defined for INT_MIN as well (-> 2^31)
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
    uint32_t stored_dirty; // ch_stored was rewritten during this launch (StoreAdaptiveChannel): it has to be written back
    uint32_t phase;        // parity of the mbarrier phase of this row's bulk loads (flips from row to row)
    uint32_t pad_[2];
    int32_t lane_row[32];  // the persistent loop's row counter, one word per lane (see aecm_call_kernel)
    int32_t lane_ms[32];   // msInSndCardBuf of the current row, one word per lane
};

struct Ctx {
    AecmInstanceState *st;  // shared-memory copy of the record
    WarpScratch *sc;
    // The three HBM planes are addressed as (plane base, instance): the bases are kernel parameters (uniform, never
    // held in per-thread registers), so an instance costs one register here instead of three 64-bit pointers.
    AecmInstanceState *state_base;
    uint16_t *far_hist_base;
    int16_t *far_ring_base;
    int inst;
    int lane;
    AecmDelayExt *ext;      // SURVEY 8-f4 record of this instance in HBM (kExt instantiation only; otherwise unused)
    // the record in HBM (write-through target of the fields that are not bulk-stored back)
    __device__ __forceinline__ AecmInstanceState *g() const { return state_base + inst; }
    // this instance's 100 x 66 rows in HBM
    __device__ __forceinline__ uint16_t *far_hist() const { return far_hist_base + (size_t)inst * (AECM_MAX_DELAY * AECM_HIST_ROW); }
    // this instance's 4000-sample jitter FIFO in HBM
    __device__ __forceinline__ int16_t *far_ring() const { return far_ring_base + (size_t)inst * AECM_FAR_RING; }
};

// wait for a bulk load (the phase of its mbarrier with the given parity: the persistent loop alternates 0, 1, 0, ...
// from row to row); returns at once when the data already arrived
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#ifdef AECM_HOST_EMU  // tests/host_emu: the record is already in place
    (void)bar;
    (void)parity;
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_RECORD:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_RECORD;\n"
        "bra WAIT_RECORD;\n"
        "DONE_RECORD:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
#endif
}

// ------------------------------------------------------------------------------------------------
// scalar primitives (spl_inl.h, signal_processing_library.{h,cc})
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int s16(int v) { return (int)(short)v; }  // int16 truncation + sign extension
__device__ __forceinline__ int norm_w32(int a) { return a == 0 ? 0 : __clz(a < 0 ? ~a : a) - 1; }   // spl_inl.h:96
__device__ __forceinline__ int norm_u32(unsigned a) { return a == 0 ? 0 : __clz((int)a); }          // spl_inl.h:101
__device__ __forceinline__ int norm_w16(int a) { return a == 0 ? 0 : __clz(a < 0 ? ~a : a) - 17; } // spl_inl.h:107
__device__ __forceinline__ int add_sat32(int a, int b) {  // spl_inl.h:70-82
#ifdef AECM_HOST_EMU
    const long long w = (long long)a + b;
    return (int)(w > 2147483647LL ? 2147483647LL : (w < -2147483648LL ? -2147483648LL : w));
#else
    int r;
    asm("add.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#endif
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
#ifdef AECM_HOST_EMU  // any estimate within +-1 of the root will do (see above): the host's sqrtf
    f = sqrtf((float)(uint32_t)v);
#else
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(__uint2float_rn((uint32_t)v)));
#endif
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
    const unsigned v = (unsigned)(base + k);
    return (int)min(v, v - (unsigned)cap);  // v - cap wraps to a huge value exactly when v < cap
}

}  // namespace

#endif  // AECM_PRIMS_CUH_
