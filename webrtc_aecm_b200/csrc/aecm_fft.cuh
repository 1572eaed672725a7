// aecm_fft.cuh -- bit-exact 128-point fixed-point complex FFT / IFFT of the reference (complex_fft.c:241-491,
// mode 1, after the bit reversal of :181-209) held in the registers of one warp.
//
// Data layout.  Position p (0..127) is the index into the reference's bit-reversed work buffer.  Each lane
// owns four positions; the set changes between stage groups so that consecutive radix-2 stages pair positions
// that live in the same lane ("layouts" A..D).  Between groups the 128 packed complex words take one trip
// through a 512-byte shared-memory tile whose word index is XOR-swizzled so that every access pattern used
// here (128-bit rows, stride 4, stride 16, stride 32) is bank-conflict free.
//
//   layout A  stages 0,1   p = 4*L + r                        (L = lane, r = 0..3)
//   layout B  stages 2,3   p = (L>>2)*16 + r*4 + (L&3)
//   layout C  stages 4,5   p = (L>>4)*64 + r*16 + (L&15)
//   layout D  stage  6     p = (r>>1)*64 + (r&1)*32 + L        -> lane L ends with bins/samples L, L+32, L+64, L+96
//
// Butterfly arithmetic.  Per radix-2 butterfly with Q15 twiddle (wr, wi) the reference computes
// (complex_fft.c:332-350 forward, :465-482 inverse)
//       tr = (wr*xr - wi*xi + 1) >> 1            out_i = (q*2^14 + tr + rnd) >> sh   (truncated to int16)
//       ti = (wr*xi + wi*xr + 1) >> 1            out_j = (q*2^14 - tr + rnd) >> sh
// with (rnd, sh) = (2^14, 15) forward and (2^13 << s, 14 + s), s in {0,1,2}, inverse.  With floor shifts,
// (A + (B >> 1)) >> n == (2A + B) >> (n+1) and (A - (B >> 1)) >> n == (2A - B + 1) >> (n+1) for integers A, B.
// With P = wr*xr - wi*xi every output is therefore bits 16..31 of one 32-bit value
//       forward / s=1 :  V_i = q*2^15 + 2^15 + P + 1          V_j = q*2^15 + 2^15 - P
//       s = 0        :  V_i = q*2^16 + 2^15 + 2P + 2         V_j = q*2^16 + 2^15 - 2P
//       s = 2        :  V_i = q*2^14 + 2^15 + ((P+1)>>1)     V_j = q*2^14 + 2^15 - ((P+1)>>1)
// evaluated in wrapping 32-bit arithmetic: the reference's int16 truncation (including its wrap-around when a
// sum leaves the int16 range) is exactly "take the high half", so no shift and no re-truncation is issued per
// output and packing a complex result is a single PRMT.
#ifndef AECM_B200_FFT_CUH_
#define AECM_B200_FFT_CUH_

#include <stdint.h>

namespace aecm_fft {

constexpr unsigned kFull = 0xffffffffu;

// XOR swizzle: logical position p -> word index in the 128-word transpose tile (linear over GF(2))
__host__ __device__ constexpr int phys(int p) {
    int g = p >> 2;
    g ^= ((g >> 3) & 3) ^ (((g >> 4) & 1) << 2);
    return (g << 2) | (p & 3);
}
__host__ __device__ constexpr int brev7c(int n) {
    int r = 0;
    for (int b = 0; b < 7; b++) r |= ((n >> b) & 1) << (6 - b);
    return r;
}

// per-lane word offsets into a tile, computed once per CTA (see fill_lane_words)
struct LaneWords {
    int a;   // layout A row: word of position 4L (128-bit aligned)
    int b;   // layout B, r = 0; r adds XOR 4r
    int c;   // layout C, r = 0; r adds XOR {0, 16, 36, 52}
    int d;   // layout D, r = 0; r adds XOR {0, 36, 88, 124}
};
struct LaneBins {
    int k0, k1;  // words of the bit-reversed positions of bins L and L+32
    int m0, m1;  // ... of their conjugate-symmetric partners 128-L and 96-L
};
__device__ __forceinline__ LaneWords lane_words(int L) {
    LaneWords w;
    w.a = phys(4 * L);
    w.b = phys((L >> 2) * 16 + (L & 3));
    w.c = phys((L >> 4) * 64 + (L & 15));
    w.d = phys(L);
    return w;
}
__device__ __forceinline__ LaneBins lane_bins(int L) {
    LaneBins w;
    w.k0 = phys((int)(__brev((unsigned)L) >> 25));
    w.k1 = phys((int)(__brev((unsigned)(L + 32)) >> 25));
    w.m0 = phys((int)(__brev((unsigned)((128 - L) & 127)) >> 25));
    w.m1 = phys((int)(__brev((unsigned)(96 - L)) >> 25));
    return w;
}
// XOR deltas of the four points of a lane under each layout (phys() is linear, so phys(base ^ delta) =
// phys(base) ^ phys(delta) when base and delta occupy disjoint bits)
struct DxB { static __host__ __device__ constexpr int d(int r) { return phys(r * 4); } };
struct DxC { static __host__ __device__ constexpr int d(int r) { return phys(r * 16); } };
struct DxD { static __host__ __device__ constexpr int d(int r) { return phys((r >> 1) * 64 + (r & 1) * 32); } };
static_assert(DxB::d(1) == 4 && DxB::d(3) == 12 && DxC::d(1) == 16 && DxC::d(2) == 36 && DxC::d(3) == 52 &&
                  DxD::d(1) == 36 && DxD::d(2) == 88 && DxD::d(3) == 124, "swizzle deltas");

__device__ __forceinline__ uint32_t pack_hi(int vr, int vi) {  // (vr >> 16) | (vi & 0xffff0000)
    return __byte_perm((uint32_t)vr, (uint32_t)vi, 0x7632);
}
// sign-extended low half (the compiler emits PRMT 0x9910; the __byte_perm intrinsic would drop the sign bits)
__device__ __forceinline__ int sx_lo(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int sx_hi(uint32_t v) { return ((int)v) >> 16; }

// twiddle table entry t = j/8 of the reference's kSinTable1024: c = S[8t + 256] (cos), s = S[8t] (sin)
struct Tw {
    int c, s;
};

// AECM_FMA_BIAS (bit 0: forward butterflies, bit 1: inverse butterflies): the butterflies are written as multiply-add
// chains whose power-of-two factors the compiler cannot see (opaque()), so that it emits IMAD (fma pipe) where it would
// otherwise strength-reduce to SHF / LOP3 / IADD3 (alu pipe).
// Same values bit for bit (wrapping 32-bit arithmetic); only the pipe the work lands on changes: the kernel's mix is 50 %
// alu-pipe against 25 % fma-pipe instructions and each pipe takes one warp-instruction every two cycles.
#ifndef AECM_FMA_BIAS
#define AECM_FMA_BIAS 2
#endif
__device__ __forceinline__ int opaque(int v) {
    asm("" : "+r"(v));
    return v;
}

// warp-uniform block-floating parameters of one inverse stage (complex_fft.c:381-396)
struct InvScale {
    int qs;     // 2^(16 - shift)  (forward: 2^15, see fwd_scale)
    int mult;   // 2 for shift 0, else 1
    int d;      // 1 for shift 2, else 0
    int c_i;    // 2^15 + addi
    int addi;   // mult for shift 0/1, 0 for shift 2
    int shift;
};
__device__ __forceinline__ InvScale inv_scale(int peak) {
    InvScale s;
    s.shift = (peak > 13573) + (peak > 27146);
    s.qs = 65536 >> s.shift;
    s.mult = s.shift == 0 ? 2 : 1;
    s.d = s.shift == 2;
    s.addi = s.shift == 2 ? 0 : s.mult;
    s.c_i = 32768 + s.addi;
#if AECM_FMA_BIAS & 2
    s.qs = opaque(s.qs);
    s.mult = opaque(s.mult);
    s.c_i = opaque(s.c_i);
#endif
    return s;
}
__device__ __forceinline__ InvScale fwd_scale() {  // the forward transform only uses qs (as the opaque factor 2^15)
    InvScale s{};
    s.qs = (AECM_FMA_BIAS & 1) ? opaque(32768) : 32768;
    return s;
}

// One radix-2 butterfly on sign-extended int16 inputs; outputs are the raw 32-bit V values (see header).
// kInv=false: forward (wi = -s).  kInv=true: inverse (wi = +s).
template <bool kInv>
__device__ __forceinline__ void butterfly(int qr, int qi, int xr, int xi, Tw w, const InvScale &sc, int &vir, int &vii,
                                          int &vjr, int &vji) {
    if (!kInv && (AECM_FMA_BIAS & 1)) {  // V_i = q*2^15 + 2^15 + 1 + P as a chain of three IMADs, V_j = (2 q*2^15 + 2^16 + 1) - V_i
        const int br = qr * sc.qs + 32769, bi = qi * sc.qs + 32769;
        const int cr = qr * sc.qs + 32768, ci = qi * sc.qs + 32768;
        vir = (int)((uint32_t)(w.c * xr) + (uint32_t)opaque(w.s * xi + br));
        vii = (int)((uint32_t)(w.c * xi) - (uint32_t)opaque(w.s * xr - bi));
        vjr = (int)((uint32_t)br + (uint32_t)cr - (uint32_t)vir);
        vji = (int)((uint32_t)bi + (uint32_t)ci - (uint32_t)vii);
    } else if (kInv && (AECM_FMA_BIAS & 2)) {
        const int P = (w.c * xr + sc.d - w.s * xi) >> sc.d;
        const int Q = (w.c * xi + sc.d + w.s * xr) >> sc.d;
        const int br = qr * sc.qs + sc.c_i, bi = qi * sc.qs + sc.c_i;
        const int cr = qr * sc.qs + 32768, ci = qi * sc.qs + 32768;  // c_i - addi = 2^15 for every shift
        vir = (int)((uint32_t)P * (uint32_t)sc.mult + (uint32_t)br);
        vii = (int)((uint32_t)Q * (uint32_t)sc.mult + (uint32_t)bi);
        vjr = (int)((uint32_t)br + (uint32_t)cr - (uint32_t)vir);
        vji = (int)((uint32_t)bi + (uint32_t)ci - (uint32_t)vii);
    } else if (!kInv) {
        const int P = w.c * xr + w.s * xi;
        const int Q = w.c * xi - w.s * xr;
        const int br = qr * 32768 + 32769, bi = qi * 32768 + 32769;
        vir = br + P;
        vii = bi + Q;
        vjr = br - 1 - P;
        vji = bi - 1 - Q;
    } else {
        const int P = (w.c * xr + sc.d - w.s * xi) >> sc.d;
        const int Q = (w.c * xi + sc.d + w.s * xr) >> sc.d;
        const int br = qr * sc.qs + sc.c_i, bi = qi * sc.qs + sc.c_i;
        const int pm = P * sc.mult, qm = Q * sc.mult;
        vir = br + pm;
        vii = bi + qm;
        vjr = br - sc.addi - pm;
        vji = bi - sc.addi - qm;
    }
}

// Four complex points per lane for each of NS independent signals.
template <int NS>
struct Points {
    int re[NS][4], im[NS][4];
};

// one local stage on sign-extended values.  kPair = 0 pairs (0,1),(2,3); kPair = 1 pairs (0,2),(1,3).
// w0 / w1 are the twiddles of the first / second pair.  Leaves raw V values in z.
template <int NS, bool kInv, int kPair>
__device__ __forceinline__ void local_stage(Points<NS> &z, Tw w0, Tw w1, const InvScale &sc) {
    constexpr int i0 = 0, j0 = kPair ? 2 : 1, i1 = kPair ? 1 : 2, j1 = 3;
#pragma unroll
    for (int n = 0; n < NS; n++) {
        int a, b, c, d;
        butterfly<kInv>(z.re[n][i0], z.im[n][i0], z.re[n][j0], z.im[n][j0], w0, sc, a, b, c, d);
        z.re[n][i0] = a; z.im[n][i0] = b; z.re[n][j0] = c; z.im[n][j0] = d;
        butterfly<kInv>(z.re[n][i1], z.im[n][i1], z.re[n][j1], z.im[n][j1], w1, sc, a, b, c, d);
        z.re[n][i1] = a; z.im[n][i1] = b; z.re[n][j1] = c; z.im[n][j1] = d;
    }
}

template <int NS>
__device__ __forceinline__ void extract(Points<NS> &z) {  // raw V -> sign-extended int16 value
#pragma unroll
    for (int n = 0; n < NS; n++)
#pragma unroll
        for (int r = 0; r < 4; r++) {
            z.re[n][r] >>= 16;
            z.im[n][r] >>= 16;
        }
}

// per-lane max |value| over the extracted points (result still needs the warp reduction)
template <int NS>
__device__ __forceinline__ int lane_peak(const Points<NS> &z, int n) {
    int mx = 0, mn = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        mx = __vimax3_s32(mx, z.re[n][r], z.im[n][r]);
        mn = __vimin3_s32(mn, z.re[n][r], z.im[n][r]);
    }
    return max(mx, -mn);
}

// tile round trip: store raw V values packed under one layout, load (and sign-extend) under the next.
template <int NS, class Dx>
__device__ __forceinline__ void store_words(const Points<NS> &z, uint32_t *tile, int w0) {
#pragma unroll
    for (int n = 0; n < NS; n++)
#pragma unroll
        for (int r = 0; r < 4; r++) tile[n * 128 + (w0 ^ Dx::d(r))] = pack_hi(z.re[n][r], z.im[n][r]);
}
template <int NS, class Dx>
__device__ __forceinline__ void load_words(Points<NS> &z, const uint32_t *tile, int w0) {
#pragma unroll
    for (int n = 0; n < NS; n++)
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t v = tile[n * 128 + (w0 ^ Dx::d(r))];
            z.re[n][r] = sx_lo(v);
            z.im[n][r] = sx_hi(v);
        }
}

// lane-dependent twiddles of stages 2..6 (t = m << (6 - s), m = p & (2^s - 1))
struct LaneTwiddles {
    Tw s2, s3a, s3b, s4, s5a, s5b, s6a, s6b;
};
__device__ __forceinline__ LaneTwiddles load_twiddles(const int2 *tab, int L) {
    LaneTwiddles t;
    auto get = [&](int i) { const int2 v = tab[i]; return Tw{v.x, v.y}; };
    const int lo2 = L & 3, lo4 = L & 15;
    t.s2 = get(lo2 * 16);
    t.s3a = get(lo2 * 8);
    t.s3b = get(32 + lo2 * 8);
    t.s4 = get(lo4 * 4);
    t.s5a = get(lo4 * 2);
    t.s5b = get(32 + lo4 * 2);
    t.s6a = get(L);
    t.s6b = get(32 + L);
    return t;
}

// ---------------------------------------------------------------------------------------------------------
// Forward transform of NS real signals (real_fft.c:47-72 + complex_fft.c:293-357).
//   x[n][k]  : windowed sample of signal n at time index brev5(lane) + 32*k, k = 0..3 (sign-extended int16)
//   out_*[n][0..2] : bins lane, lane+32 and 64+lane (sign-extended), before the sign flip of aecm_core_c.cc:188-190
// tile: NS*128 words of shared memory private to the warp.
// ---------------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void forward128(const int (&x)[NS][4], uint32_t *tile, const int2 *tw_tab, const LaneWords lw,
                                           int lane, int (&out_re)[NS][3], int (&out_im)[NS][3]) {
    Points<NS> z;
    const InvScale none = fwd_scale();
    // layout A: position 4L + r holds time index brev7(4L + r) = brev5(L) + 32 * {0, 2, 1, 3}[r]
#pragma unroll
    for (int n = 0; n < NS; n++) {
        z.re[n][0] = x[n][0]; z.re[n][1] = x[n][2]; z.re[n][2] = x[n][1]; z.re[n][3] = x[n][3];
#pragma unroll
        for (int r = 0; r < 4; r++) z.im[n][r] = 0;
    }
    constexpr Tw w_one{32767, 0}, w_quarter{0, 32767};  // t = 0 and t = 32
    local_stage<NS, false, 0>(z, w_one, w_one, none);      // stage 0: m = 0
    extract(z);
    local_stage<NS, false, 1>(z, w_one, w_quarter, none);  // stage 1: m = p & 1 = r & 1
    // A -> B
    {
        uint4 *row = reinterpret_cast<uint4 *>(tile);
#pragma unroll
        for (int n = 0; n < NS; n++)
            row[n * 32 + (lw.a >> 2)] = make_uint4(pack_hi(z.re[n][0], z.im[n][0]), pack_hi(z.re[n][1], z.im[n][1]),
                                                   pack_hi(z.re[n][2], z.im[n][2]), pack_hi(z.re[n][3], z.im[n][3]));
    }
    const LaneTwiddles t = load_twiddles(tw_tab, lane);
    __syncwarp();
    load_words<NS, DxB>(z, tile, lw.b);
    local_stage<NS, false, 0>(z, t.s2, t.s2, none);
    extract(z);
    local_stage<NS, false, 1>(z, t.s3a, t.s3b, none);
    __syncwarp();
    store_words<NS, DxB>(z, tile, lw.b);
    __syncwarp();
    load_words<NS, DxC>(z, tile, lw.c);
    local_stage<NS, false, 0>(z, t.s4, t.s4, none);
    extract(z);
    local_stage<NS, false, 1>(z, t.s5a, t.s5b, none);
    __syncwarp();
    store_words<NS, DxC>(z, tile, lw.c);
    __syncwarp();
    load_words<NS, DxD>(z, tile, lw.d);
    // stage 6: only bins 0..64 are kept (real_fft.c:69): the "+" outputs of both pairs and the "-" output of pair (0,2)
    local_stage<NS, false, 1>(z, t.s6a, t.s6b, none);
#pragma unroll
    for (int n = 0; n < NS; n++) {
        out_re[n][0] = z.re[n][0] >> 16; out_im[n][0] = z.im[n][0] >> 16;
        out_re[n][1] = z.re[n][1] >> 16; out_im[n][1] = z.im[n][1] >> 16;
        out_re[n][2] = z.re[n][2] >> 16; out_im[n][2] = z.im[n][2] >> 16;
    }
    __syncwarp();  // the tile may be reused by the caller
}

// ---------------------------------------------------------------------------------------------------------
// Inverse transform (real_fft.c:74-102 + complex_fft.c:361-491).  The caller has written the 128 packed
// complex inputs (conjugate-symmetric extension included) to the tile at phys(bit-reversed position) and
// passes the peak |component| of that data.  Returns the accumulated scale; out_re[k] = real part of time
// sample lane + 32k (sign-extended), k = 0..3.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_peak(int lane_max) { return min(__reduce_max_sync(kFull, lane_max), 32767); }

__device__ __forceinline__ int inverse128(uint32_t *tile, const int2 *tw_tab, const LaneWords lw, int lane, int peak,
                                          int (&out_re)[4]) {
    Points<1> z;
    int scale = 0;
    constexpr Tw w_one{32767, 0}, w_quarter{0, 32767};
    {
        const uint4 v = reinterpret_cast<const uint4 *>(tile)[lw.a >> 2];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int r = 0; r < 4; r++) { z.re[0][r] = sx_lo(w[r]); z.im[0][r] = sx_hi(w[r]); }
    }
    const LaneTwiddles t = load_twiddles(tw_tab, lane);
    InvScale sc = inv_scale(peak);
    scale += sc.shift;
    local_stage<1, true, 0>(z, w_one, w_one, sc);  // stage 0
    extract(z);
    sc = inv_scale(warp_peak(lane_peak(z, 0)));
    scale += sc.shift;
    local_stage<1, true, 1>(z, w_one, w_quarter, sc);  // stage 1
    __syncwarp();
    reinterpret_cast<uint4 *>(tile)[lw.a >> 2] =
        make_uint4(pack_hi(z.re[0][0], z.im[0][0]), pack_hi(z.re[0][1], z.im[0][1]), pack_hi(z.re[0][2], z.im[0][2]),
                   pack_hi(z.re[0][3], z.im[0][3]));
    __syncwarp();
    load_words<1, DxB>(z, tile, lw.b);
    sc = inv_scale(warp_peak(lane_peak(z, 0)));
    scale += sc.shift;
    local_stage<1, true, 0>(z, t.s2, t.s2, sc);  // stage 2
    extract(z);
    sc = inv_scale(warp_peak(lane_peak(z, 0)));
    scale += sc.shift;
    local_stage<1, true, 1>(z, t.s3a, t.s3b, sc);  // stage 3
    __syncwarp();
    store_words<1, DxB>(z, tile, lw.b);
    __syncwarp();
    load_words<1, DxC>(z, tile, lw.c);
    sc = inv_scale(warp_peak(lane_peak(z, 0)));
    scale += sc.shift;
    local_stage<1, true, 0>(z, t.s4, t.s4, sc);  // stage 4
    extract(z);
    sc = inv_scale(warp_peak(lane_peak(z, 0)));
    scale += sc.shift;
    local_stage<1, true, 1>(z, t.s5a, t.s5b, sc);  // stage 5
    __syncwarp();
    store_words<1, DxC>(z, tile, lw.c);
    __syncwarp();
    load_words<1, DxD>(z, tile, lw.d);
    sc = inv_scale(warp_peak(lane_peak(z, 0)));
    scale += sc.shift;
    local_stage<1, true, 1>(z, t.s6a, t.s6b, sc);  // stage 6 (imaginary outputs are dead code: real_fft.c:97-99)
#pragma unroll
    for (int r = 0; r < 4; r++) out_re[r] = z.re[0][r] >> 16;
    __syncwarp();
    return scale;
}

}  // namespace aecm_fft
#endif  // AECM_B200_FFT_CUH_
