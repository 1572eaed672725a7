// warp_emu.h -- TEST INFRASTRUCTURE.  Runs per-lane CUDA device code of webrtc_aecm_b200/csrc on the CPU: one
// std::thread per lane, __syncwarp()/__reduce_*_sync()/__ballot_sync() implemented with a barrier.  Lets the
// CPU test-suite check the register FFT (aecm_fft.cuh) bit for bit against the oracle without a GPU.
#pragma once
#include <pthread.h>
#include <stdint.h>

#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__

struct uint4 { uint32_t x, y, z, w; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

namespace warp_emu {
inline pthread_barrier_t g_bar;
inline int g_xchg[32];
inline thread_local int t_lane = 0;
inline void run(const std::function<void(int)> &body) {
    pthread_barrier_init(&g_bar, nullptr, 32);
    std::vector<std::thread> th;
    for (int l = 0; l < 32; l++) th.emplace_back([&, l] { t_lane = l; body(l); });
    for (auto &t : th) t.join();
    pthread_barrier_destroy(&g_bar);
}
template <class Op>
inline int reduce(int v, Op op) {
    g_xchg[t_lane] = v;
    pthread_barrier_wait(&g_bar);
    int r = g_xchg[0];
    for (int i = 1; i < 32; i++) r = op(r, g_xchg[i]);
    pthread_barrier_wait(&g_bar);
    return r;
}
}  // namespace warp_emu

static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&warp_emu::g_bar); }
static inline int __reduce_max_sync(unsigned, int v) { return warp_emu::reduce(v, [](int a, int b) { return std::max(a, b); }); }
static inline int __reduce_min_sync(unsigned, int v) { return warp_emu::reduce(v, [](int a, int b) { return std::min(a, b); }); }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    return (unsigned)warp_emu::reduce((int)v, [](int a, int b) { return (int)((unsigned)a + (unsigned)b); });
}
static inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline int __vimax3_s32(int a, int b, int c) { return std::max(a, std::max(b, c)); }
static inline int __vimin3_s32(int a, int b, int c) { return std::min(a, std::min(b, c)); }
// PTX prmt.b32 default mode: nibble bits 0..2 pick a byte of {b:a}, bit 3 replicates that byte's sign
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned n = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(v >> (8 * (n & 7))) & 0xff;
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
using std::max;
using std::min;
