// warp_emu.h -- TEST INFRASTRUCTURE.  Runs per-lane CUDA device code of webrtc_aecm_b200/csrc on the CPU.
//
// A warp is 32 cooperative fibers on ONE host thread (ucontext), scheduled round-robin in lane order: a lane runs
// until its next warp-level operation (__syncwarp, __reduce_*_sync, __ballot_sync, __any_sync), then the next lane
// runs up to the same point; when lane 31 arrives, lane 0 continues.  This is exactly the guarantee the device code
// relies on (all lanes reach every warp-level operation; shared-memory writes before it are visible after it), with
// the execution between two such points fully serialised, so the emulation is deterministic.  Collectives exchange
// their operands through two alternating buffers: a lane that runs ahead to the NEXT collective must not overwrite
// what slower lanes still have to read from the current one.
//
// Lets the CPU test-suite run the device code -- the register FFT, and the whole BufferFarend/Process call path --
// bit for bit against the oracle without a GPU.  (Code that executes a different number of warp-level operations on
// different lanes is a bug on the device too; here it shows up as a mismatch or a hang.)
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __align__(n) alignas(n)

struct uint4 { uint32_t x, y, z, w; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

namespace warp_emu {
constexpr int kLanes = 32;
constexpr size_t kStack = 1 << 20;
inline ucontext_t g_ctx[kLanes], g_main;
inline char *g_stack[kLanes];
inline int g_cur = 0;
inline const std::function<void(int)> *g_body = nullptr;
inline uint32_t g_xchg[2][kLanes];
inline unsigned g_epoch[kLanes];  // collectives executed so far, per lane (selects the exchange buffer)
inline long g_barriers = 0;

inline void barrier() {  // hand the thread to the next lane; returns when every lane has arrived here
    const int me = g_cur, next = (me + 1) % kLanes;
    g_barriers++;
    g_cur = next;
    swapcontext(&g_ctx[me], &g_ctx[next]);
}
inline void trampoline(int lane) {
    (*g_body)(lane);
    // lanes finish in order (they all executed the same number of barriers): pass the thread on, lane 31 returns it
    g_cur = lane + 1;
    if (lane + 1 < kLanes) setcontext(&g_ctx[lane + 1]);
    setcontext(&g_main);
}
inline void run(const std::function<void(int)> &body) {
    g_body = &body;
    for (int l = 0; l < kLanes; l++) {
        if (!g_stack[l]) g_stack[l] = (char *)malloc(kStack);
        getcontext(&g_ctx[l]);
        g_ctx[l].uc_stack.ss_sp = g_stack[l];
        g_ctx[l].uc_stack.ss_size = kStack;
        g_ctx[l].uc_link = nullptr;
        makecontext(&g_ctx[l], (void (*)())trampoline, 1, l);
        g_epoch[l] = 0;
    }
    g_cur = 0;
    swapcontext(&g_main, &g_ctx[0]);
}
inline int lane() { return g_cur; }
// every lane contributes v; returns a pointer to all 32 contributions
inline const uint32_t *exchange(uint32_t v) {
    const int me = g_cur;
    uint32_t *buf = g_xchg[g_epoch[me]++ & 1];
    buf[me] = v;
    barrier();
    return buf;
}
}  // namespace warp_emu

static inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::barrier(); }
static inline void __syncthreads() { warp_emu::barrier(); }  // only reached by experiments that are compiled out
static inline int __reduce_max_sync(unsigned, int v) {
    const uint32_t *x = warp_emu::exchange((uint32_t)v);
    int r = (int)x[0];
    for (int i = 1; i < 32; i++) r = std::max(r, (int)x[i]);
    return r;
}
static inline int __reduce_min_sync(unsigned, int v) {
    const uint32_t *x = warp_emu::exchange((uint32_t)v);
    int r = (int)x[0];
    for (int i = 1; i < 32; i++) r = std::min(r, (int)x[i]);
    return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    const uint32_t *x = warp_emu::exchange(v);
    unsigned r = x[0];
    for (int i = 1; i < 32; i++) r = std::min(r, x[i]);
    return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    const uint32_t *x = warp_emu::exchange(v);
    unsigned r = x[0];
    for (int i = 1; i < 32; i++) r = std::max(r, x[i]);
    return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    const uint32_t *x = warp_emu::exchange(v);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r += x[i];
    return r;
}
static inline int __reduce_add_sync(unsigned, int v) { return (int)__reduce_add_sync(0xffffffffu, (unsigned)v); }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    const uint32_t *x = warp_emu::exchange(v);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= x[i];
    return r;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    const uint32_t *x = warp_emu::exchange(pred ? 1u : 0u);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (x[i] & 1u) << i;
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)warp_emu::exchange((uint32_t)v)[src & 31]; }

template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __mulhi(int a, int b) { return (int)(((int64_t)a * b) >> 32); }
static inline int __float2int_rz(float f) { return (int)f; }
// single-precision operations with one rounding each (the emulation is built with -ffp-contract=off)
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __int2float_rn(int v) { return (float)v; }
static inline float __uint2float_rn(unsigned v) { return (float)v; }
static inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline int __vimax3_s32(int a, int b, int c) { return std::max(a, std::max(b, c)); }
static inline int __vimin3_s32(int a, int b, int c) { return std::min(a, std::min(b, c)); }
// __byte_perm: the CUDA intrinsic masks every selector nibble to 3 bits (no sign replication, unlike raw PTX prmt)
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= ((unsigned)(v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
using std::max;
using std::min;
static inline int max(int a, unsigned b) { return std::max(a, (int)b); }
static inline unsigned min(unsigned a, int b) { return std::min(a, (unsigned)b); }
