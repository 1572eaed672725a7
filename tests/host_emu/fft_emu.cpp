// fft_emu.cpp -- TEST INFRASTRUCTURE.  Executes aecm_fft.cuh on an emulated warp and exposes C entry points
// with the same contract as the oracle's oracle_real_fft128_forward / _inverse, so tests can compare them.
#define AECM_HOST_EMU 1
#include "warp_emu.h"

#include <string.h>

#include "../../webrtc_aecm_b200/csrc/aecm_fft.cuh"
#include "../../webrtc_aecm_b200/csrc/aecm_tables.h"

using namespace aecm_fft;

static void make_tw(int2 *tw) {
    for (int t = 0; t < 64; t++) {
        const uint32_t w = kAecmTwiddle128[t];
        tw[t] = make_int2((int)(int16_t)(w & 0xffff), (int)(int16_t)(w >> 16));
    }
}

extern "C" {

// ns signals of 128 windowed real samples each -> per signal bins 0..64 interleaved re,im (130 values)
void emu_fft128_forward(const int16_t *in, int16_t *out, int ns) {
    alignas(16) static uint32_t tile[256];
    static int2 tw[64];
    make_tw(tw);
    warp_emu::run([&](int lane) {
        const int n0 = (int)(__brev((unsigned)lane) >> 27);
        const LaneWords lw = lane_words(lane);
        if (ns == 1) {
            int x[1][4], re[1][3], im[1][3];
            for (int k = 0; k < 4; k++) x[0][k] = in[n0 + 32 * k];
            forward128<1>(x, tile, tw, lw, lane, re, im);
            for (int p = 0; p < 3; p++) {
                const int bin = p < 2 ? lane + 32 * p : 64;
                if (p < 2 || lane == 0) { out[2 * bin] = (int16_t)re[0][p]; out[2 * bin + 1] = (int16_t)im[0][p]; }
            }
        } else {
            int x[2][4], re[2][3], im[2][3];
            for (int n = 0; n < 2; n++) for (int k = 0; k < 4; k++) x[n][k] = in[n * 128 + n0 + 32 * k];
            forward128<2>(x, tile, tw, lw, lane, re, im);
            for (int n = 0; n < 2; n++) for (int p = 0; p < 3; p++) {
                const int bin = p < 2 ? lane + 32 * p : 64;
                if (p < 2 || lane == 0) { out[n * 130 + 2 * bin] = (int16_t)re[n][p]; out[n * 130 + 2 * bin + 1] = (int16_t)im[n][p]; }
            }
        }
    });
}

// 65 complex bins (interleaved, as handed to WebRtcSpl_RealInverseFFT) -> 128 real samples; returns scale
int emu_fft128_inverse(const int16_t *in, int16_t *out) {
    alignas(16) static uint32_t tile[128];
    static int2 tw[64];
    static int scale_out;
    make_tw(tw);
    warp_emu::run([&](int lane) {
        const LaneWords lw = lane_words(lane);
        const LaneBins lb = lane_bins(lane);
        const int words[2] = {lb.k0, lb.k1}, mirrors[2] = {lb.m0, lb.m1};
        int peak = 0;
        for (int p = 0; p < 3; p++) {
            const int i = p < 2 ? lane + 32 * p : 64;
            if (p < 2 || lane == 0) {
                const int re = in[2 * i], im = in[2 * i + 1];
                tile[p < 2 ? words[p] : phys(1)] = (uint32_t)(uint16_t)re | ((uint32_t)(uint16_t)im << 16);
                peak = max(peak, max(abs(re), abs(im)));
                if (p < 2 && i > 0) tile[mirrors[p]] = (uint32_t)(uint16_t)re | ((uint32_t)(uint16_t)(int16_t)(-im) << 16);
            }
        }
        peak = min(__reduce_max_sync(kFull, peak), 32767);
        __syncwarp();
        int td[4];
        const int sc = inverse128(tile, tw, lw, lane, peak, td);
        for (int k = 0; k < 4; k++) out[lane + 32 * k] = (int16_t)td[k];
        if (lane == 0) scale_out = sc;
    });
    return scale_out;
}
}
