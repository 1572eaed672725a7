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
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#include "aecm_frame_api.cuh"  // -> aecm_block.cuh -> aecm_prims.cuh -> aecm_fft.cuh, aecm_state.h, aecm_tables.h
#include "aecm_init.cuh"

namespace {

// ---- 1-D bulk async copies (TMA unit) between the HBM record and the warp's shared-memory copy ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
}
// Issues the two bulk loads of the record (lane 0).  Part A is waited for after the row's CTA rendezvous, part B where it
// is first used.
__device__ __forceinline__ void bulk_issue_record(AecmInstanceState *dst_smem, const AecmInstanceState *src_gmem, uint32_t bytes,
                                                  uint64_t *bar_a, uint64_t *bar_b, int lane) {
    if (lane == 0) {
        const uint32_t a = smem_u32(bar_a), b = smem_u32(bar_b);
        const uint32_t bytes_b = bytes - AECM_STATE_PART_A;
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
}
__device__ __forceinline__ void prefetch_line(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_line_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {  // bytes: multiple of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_piece(void *rec_gmem, const void *rec_smem, uint32_t from, uint32_t to) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char *>(rec_gmem) + from),
                 "r"(smem_u32(reinterpret_cast<const char *>(rec_smem) + from)), "r"(to - from)
                 : "memory");
}
// Write the record back.  Two fields are left out: the binary far-end history (one word changes per block: written
// through where it is produced, aecm_block.cuh) and the stored channel unless StoreAdaptiveChannel ran (aecm_core.cc:286-306
// fires once per >= 30 qualifying blocks after start-up).
__device__ __forceinline__ void bulk_store_record(AecmInstanceState *dst_gmem, const AecmInstanceState *src_smem, uint32_t bytes,
                                                  bool stored_dirty, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
    __syncwarp();
    if (lane == 0) {
        constexpr uint32_t o_bin = offsetof(AecmInstanceState, bin_far), o_mean = offsetof(AecmInstanceState, mean_bit_counts);
        constexpr uint32_t o_st = offsetof(AecmInstanceState, ch_stored), o_a16 = offsetof(AecmInstanceState, ch_adapt16);
        static_assert(o_bin % 16 == 0 && o_mean % 16 == 0 && o_st % 16 == 0 && o_a16 % 16 == 0, "bulk copies move 16-byte units");
        bulk_store_piece(dst_gmem, src_smem, 0, o_bin);
        if (stored_dirty) {
            bulk_store_piece(dst_gmem, src_smem, o_mean, bytes);
        } else {
            bulk_store_piece(dst_gmem, src_smem, o_mean, o_st);
            bulk_store_piece(dst_gmem, src_smem, o_a16, bytes);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the shared-memory source may be reused
    }
}

// shared memory: kWarpsPerCta x (AecmInstanceState, WarpScratch).  Persistent: warp w of CTA b owns the rows
// b * kWarpsPerCta + w + k * (warps in the grid), k = 0, 1, ...; the mbarriers of its record buffer flip phase per row.
template <bool kExt>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm) aecm_call_kernel(AecmBatchPlanes planes, AecmCallArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wbase = smem_raw + (size_t)warp * (sizeof(AecmInstanceState) + sizeof(WarpScratch));
    AecmInstanceState *st = reinterpret_cast<AecmInstanceState *>(wbase);
    WarpScratch *sc = reinterpret_cast<WarpScratch *>(wbase + sizeof(AecmInstanceState));

    const bool use_clean = a.near_clean != nullptr;
    const uint32_t rec_bytes = use_clean ? AECM_STATE_BYTES : AECM_STATE_BYTES_NOCLEAN;
    if (lane == 0) {
        mbar_init(&sc->mbar_a);
        mbar_init(&sc->mbar_b);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sc->phase = 0;  // parity of the mbarrier phase the next bulk load completes; kept in shared memory, not in a register
    }
    __syncwarp();

    // The row counter lives in shared memory between rows (written and read by every lane for itself through its own
    // word), not in a register: the block code below needs all 72 registers and nothing may be spilled to local memory.
    volatile int *const my_row = reinterpret_cast<volatile int *>(&sc->lane_row[lane]);
    *my_row = blockIdx.x * kWarpsPerCta + warp;
#pragma unroll 1
    for (;;) {
        const int stride = (int)gridDim.x * kWarpsPerCta;
        const int row = *my_row;
        if (row - warp >= a.count) break;  // the CTA's rows are exhausted (same trip count for all its warps)
        *my_row = row + stride;
        // (derived from the volatile row word on purpose: otherwise the compiler hoists the per-lane prefetch offsets and
        // predicates out of the row loop and, with no register to keep them in, spills them to local memory)
        const int row_bytes = (a.samples_per_call + (row < 0 ? 1 : 0)) * (int)sizeof(int16_t);
        // which instance?  -1: this warp has no row in the CTA's last pass; -2: the row names an instance that does not exist
        int inst = row < a.count ? a.first + row : -1;
        if (inst >= 0 && a.index) {  // scattered rows (combined scalar calls): the row names its instance
            inst = __ldg(&a.index[row]);
            if ((unsigned)inst >= (unsigned)planes.n_instances) inst = -2;
        }
        if (inst >= 0) {
            // The audio rows of the first call are needed right after the record: start them on their way from HBM now,
            // one 128-byte line per lane, so that their latency overlaps the record copy instead of following it.
            const int64_t off = (int64_t)row * a.row_stride;
            const int at = min(lane * 128, row_bytes - 1);  // lane 0.. cover the row, the last one its final byte
            if (lane * 128 < row_bytes + 128) {
                if (a.far) prefetch_line(reinterpret_cast<const char *>(a.far + off) + at);
                if (a.near_noisy) prefetch_line(reinterpret_cast<const char *>(a.near_noisy + off) + at);
                if (use_clean) prefetch_line(reinterpret_cast<const char *>(a.near_clean + off) + at);
            }
            if (lane == 0) sc->stored_dirty = 0;
            // the record's bulk loads go out BEFORE the rendezvous: their latency runs behind the wait for the slowest warp
            // (+0.3 ... +0.9 %, profiles/r02_variants_session12.txt)
            bulk_issue_record(st, planes.state + inst, rec_bytes, &sc->mbar_a, &sc->mbar_b, lane);
        }
        // The CTA's rendezvous of this row: ONE call site that every warp of the CTA passes once per row, with or without
        // a row of its own (AECM_ROW_RESYNC = k > 1: only before every k-th row; measured).
        if (kRowResync && row - warp != (int)blockIdx.x * kWarpsPerCta &&
            (AECM_ROW_RESYNC == 1 || ((row - warp) / stride) % AECM_ROW_RESYNC == 0))
            __syncthreads();
        if (inst < 0) {
            if (inst == -2 && a.status) a.status[row] = AECM_BAD_PARAMETER_ERROR;  // (every lane, same word: rare path)
            if (a.near_noisy)  // keep the CTA's count of the finer rendezvous points (AECM_SYNC_POINTS experiments)
                for (int call = 0; call < a.n_calls; call++) idle_call(a.samples_per_call / AECM_FRAME_LEN);
            continue;
        }
        mbar_wait(&sc->mbar_a, sc->phase);
        // ... and the stretch of the far-end jitter ring this call will read back (echo_control_mobile.cc:369-380:
        // up to 160 samples from the read cursor, which EstBufDelay may first move by one frame either way)
        if (lane < 6) {
            int pos = st->s.far_rd - AECM_FRAME_LEN + lane * 64;
            pos += pos < 0 ? AECM_FAR_RING : 0;
            pos -= pos >= AECM_FAR_RING ? AECM_FAR_RING : 0;
            prefetch_line(planes.far_ring + (size_t)inst * AECM_FAR_RING + pos);
        }
        // The row this warp takes next: pull its record and its first audio rows into L2 behind this row's arithmetic, so
        // that the bulk load at the top of the next iteration finds them there instead of in HBM.
        {
            const int nrow = row + stride;
            if (nrow < a.count && !a.index) {
                if (lane == 0) bulk_prefetch_l2(planes.state + a.first + nrow, rec_bytes);
                const int64_t noff = (int64_t)nrow * a.row_stride;
                if (lane >= 8 && (lane - 8) * 128 < row_bytes) {
                    if (a.far) prefetch_line_l2(reinterpret_cast<const char *>(a.far + noff) + (lane - 8) * 128);
                    if (a.near_noisy) prefetch_line_l2(reinterpret_cast<const char *>(a.near_noisy + noff) + (lane - 8) * 128);
                }
            }
        }
        if (use_clean) mbar_wait(&sc->mbar_b, sc->phase);  // the clean-path tail is needed from the first frame on

        Ctx c;
        c.st = st;
        c.sc = sc;
        c.state_base = planes.state;
        c.far_hist_base = planes.far_hist;
        c.far_ring_base = planes.far_ring;
        c.inst = inst;
        c.lane = lane;
        c.ext = kExt ? planes.dly_ext + inst : nullptr;

        // (through the lane's own shared-memory word: the clamped value would otherwise be hoisted out of the call loop and,
        // with no register left for it, spilled to local memory)
        volatile int *const my_ms = reinterpret_cast<volatile int *>(&sc->lane_ms[lane]);
        *my_ms = a.ms ? (int)a.ms[row] : a.ms_uniform;
        int status = 0;
#pragma unroll 1
        for (int call = 0; call < a.n_calls; call++) {
            // (recomputed from the row and call numbers instead of carried: keeps two registers free across the block code)
            const int64_t o = (int64_t)row * a.row_stride + (int64_t)call * a.samples_per_call;
            if (call + 1 < a.n_calls && lane * 128 < row_bytes) {  // the next call's audio, one line per lane
                const int64_t nx = (o + a.samples_per_call) * (int64_t)sizeof(int16_t) + lane * 128;
                if (a.far) prefetch_line(reinterpret_cast<const char *>(a.far) + nx);
                if (a.near_noisy) prefetch_line(reinterpret_cast<const char *>(a.near_noisy) + nx);
            }
            if (a.far && (!a.far_mask || a.far_mask[(size_t)call * a.mask_stride + row])) {
                const int r = api_buffer_farend(c, a.far + o, a.samples_per_call);
                if (status == 0) status = r;
            }
            if (a.near_noisy) {
                const int r = api_process<kExt>(c, a.near_noisy + o, use_clean ? a.near_clean + o : nullptr, a.out + o,
                                          a.samples_per_call, *my_ms);
                if (status == 0) status = r;
            }
        }
        mbar_wait(&sc->mbar_b, sc->phase);  // (a start-up call never touched part B: make sure it has landed before it is stored back)
        const bool stored_dirty = sc->stored_dirty != 0;  // written by lane 0 before the __syncwarp that ended its block
        bulk_store_record(c.g(), st, rec_bytes, stored_dirty, lane);
        if (lane == 0) {
            if (a.status) a.status[row] = status;
            sc->phase ^= 1;
        }
        __syncwarp();  // the new parity is read by every lane at the top of the next row
    }
}

__global__ void aecm_init_kernel(AecmBatchPlanes planes, int first, int count, int fs) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= count) return;
    init_instance(planes.state + first + row, fs, lane);
    init_delay_ext(planes.dly_ext + first + row, lane);
}

// SURVEY 8-f4 control kernels, one warp per instance, straight on the HBM records
__global__ void aecm_dly_options_kernel(AecmBatchPlanes planes, int first, int count, int robust, int allowed_offset,
                                        int max_lookahead, int lookahead, int32_t *status) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= count) return;
    AecmInstanceState *g = planes.state + first + row;
    int r;
    if (g->s.init_flag != AECM_INIT_MAGIC) r = AECM_UNINITIALIZED_ERROR;  // the estimator's state belongs to an initialised core
    else r = set_delay_options(g, planes.dly_ext + first + row, robust, allowed_offset, max_lookahead, lookahead, lane);
    if (status && lane == 0) status[row] = r;
}

__global__ void aecm_dly_soft_reset_kernel(AecmBatchPlanes planes, int first, int count, int delay_shift, int32_t *applied) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= count) return;
    const int r = soft_reset_delay_estimator(planes.state + first + row, planes.dly_ext + first + row, delay_shift, lane);
    if (applied && lane == 0) applied[row] = r;
}

__global__ void aecm_dly_query_kernel(AecmBatchPlanes planes, int first, int count, int32_t *last_delay, float *quality) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= count) return;
    const AecmInstanceState *g = planes.state + first + row;
    if (last_delay) last_delay[row] = g->s.last_delay;
    if (quality) quality[row] = last_delay_quality(g, planes.dly_ext + first + row);
}

__global__ void aecm_set_config_kernel(AecmBatchPlanes planes, int first, int count, int cng_mode, int echo_mode,
                                       int32_t *status) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= count) return;
    const int r = set_config_checked(planes.state[first + row].s, cng_mode, echo_mode);
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

// Known-answer harness for the warp FFT exactly as aecm_call_kernel uses it (same tables, same tile layout, same
// bin -> lane mapping): one warp per vector.  mode 0: forward128<1>, 1: forward128<2> on vectors (2v, 2v+1), 2: inverse128
// with the conjugate-symmetric fill and the peak scan of process_block.  in/out are device arrays:
//   forward  in [n][128] windowed real samples   out [n][130] bins 0..64 interleaved re,im (before the sign flip)
//   inverse  in [n][130] bins 0..64 interleaved  out [n][128] real samples, scale[n] the accumulated block-floating shift
__global__ void aecm_fft_selftest_kernel(int mode, const int16_t *in, int16_t *out, int32_t *scale, int n) {
    __shared__ __align__(16) uint32_t tiles[4][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *tile = tiles[warp];
    const int v = (blockIdx.x * 4 + warp) * (mode == 1 ? 2 : 1);
    if (v >= n) return;
    const int4 lwv = __ldg(&kLaneTables.words[lane]);
    const aecm_fft::LaneWords lw{lwv.x, lwv.y, lwv.z, lwv.w};
    const int n0 = (int)(__brev((unsigned)lane) >> 27);
    if (mode == 0) {
        int x[1][4], re[1][3], im[1][3];
        for (int k = 0; k < 4; k++) x[0][k] = in[(size_t)v * 128 + n0 + 32 * k];
        aecm_fft::forward128<1>(x, tile, twiddle_table(), lw, lane, re, im);
        for (int p = 0; p < 3; p++) {
            const int bin = p < 2 ? lane + 32 * p : 64;
            if (p < 2 || lane == 0) {
                out[(size_t)v * 130 + 2 * bin] = (int16_t)re[0][p];
                out[(size_t)v * 130 + 2 * bin + 1] = (int16_t)im[0][p];
            }
        }
    } else if (mode == 1) {
        int x[2][4], re[2][3], im[2][3];
        for (int s = 0; s < 2; s++)
            for (int k = 0; k < 4; k++) x[s][k] = in[(size_t)min(v + s, n - 1) * 128 + n0 + 32 * k];
        aecm_fft::forward128<2>(x, tile, twiddle_table(), lw, lane, re, im);
        for (int s = 0; s < 2 && v + s < n; s++)
            for (int p = 0; p < 3; p++) {
                const int bin = p < 2 ? lane + 32 * p : 64;
                if (p < 2 || lane == 0) {
                    out[(size_t)(v + s) * 130 + 2 * bin] = (int16_t)re[s][p];
                    out[(size_t)(v + s) * 130 + 2 * bin + 1] = (int16_t)im[s][p];
                }
            }
    } else {
        const int4 lb = __ldg(&kLaneTables.bins[lane]);
        const int words[2] = {lb.x, lb.y}, mirrors[2] = {lb.z, lb.w};
        int peak = 0;
        for (int p = 0; p < 3; p++) {
            const int i = p < 2 ? lane + 32 * p : 64;
            if (p < 2 || lane == 0) {
                const int re = in[(size_t)v * 130 + 2 * i], im = in[(size_t)v * 130 + 2 * i + 1];
                tile[p < 2 ? words[p] : aecm_fft::phys(1)] = pack16(re, im);
                peak = max(peak, max(abs(re), abs(im)));
                if (p < 2 && i > 0) tile[mirrors[p]] = pack16(re, s16(-im));  // real_fft.c:87-91
            }
        }
        peak = min(__reduce_max_sync(kFull, peak), 32767);
        __syncwarp();
        int td[4];
        const int sc = aecm_fft::inverse128(tile, twiddle_table(), lw, lane, peak, td);
        for (int k = 0; k < 4; k++) out[(size_t)v * 128 + lane + 32 * k] = (int16_t)td[k];
        if (lane == 0) scale[v] = sc;
    }
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

int aecm_cuda_launch_dly_options(const AecmBatchPlanes *p, int first, int count, int robust, int allowed_offset, int max_lookahead,
                                 int lookahead, int32_t *status, void *stream) {
    if (count <= 0) return 0;
    aecm_dly_options_kernel<<<(count + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, first, count, robust, allowed_offset,
                                                                                            max_lookahead, lookahead, status);
    g_launches++;
    return (int)cudaGetLastError();
}

int aecm_cuda_launch_dly_soft_reset(const AecmBatchPlanes *p, int first, int count, int delay_shift, int32_t *applied, void *stream) {
    if (count <= 0) return 0;
    aecm_dly_soft_reset_kernel<<<(count + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, first, count, delay_shift, applied);
    g_launches++;
    return (int)cudaGetLastError();
}

int aecm_cuda_launch_dly_query(const AecmBatchPlanes *p, int first, int count, int32_t *last_delay, float *quality, void *stream) {
    if (count <= 0) return 0;
    aecm_dly_query_kernel<<<(count + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, first, count, last_delay, quality);
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
        for (int v = 0; v < 2; v++) {
            const void *fn = v ? (const void *)aecm_call_kernel<true> : (const void *)aecm_call_kernel<false>;
            cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return (int)e;
        }
        attr_set[dev] = true;
    }
    static std::atomic<int> sm_count[64];  // persistent grid: at most (resident CTAs per SM) x (SMs) CTAs
    if (!sm_count[dev]) {
        int n = 0;
        cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess || n <= 0) return (int)(e != cudaSuccess ? e : cudaErrorInvalidDevice);
        sm_count[dev] = n;
    }
    // AECM_B200_MAX_CTAS caps the persistent grid (diagnostics: with 1, a batch of 70 instances walks three rows per warp,
    // the last one partial, under compute-sanitizer in seconds)
    static const int max_ctas = [] {
        const char *e = getenv("AECM_B200_MAX_CTAS");
        const int v = e ? atoi(e) : 0;
        return v > 0 ? v : 1 << 30;
    }();
    const int blocks = std::min(std::min((a->count + kWarpsPerCta - 1) / kWarpsPerCta, sm_count[dev] * kCtasPerSm), max_ctas);
    if (a->ext) aecm_call_kernel<true><<<blocks, kWarpsPerCta * 32, smem, static_cast<cudaStream_t>(stream)>>>(*p, *a);
    else aecm_call_kernel<false><<<blocks, kWarpsPerCta * 32, smem, static_cast<cudaStream_t>(stream)>>>(*p, *a);
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

// diagnostics: the kernel's warp FFT on host vectors (see aecm_fft_selftest_kernel); returns a cudaError_t value
int aecm_cuda_selftest_fft(int mode, const int16_t *in, int16_t *out, int32_t *scale, int n) {
    if (n <= 0) return 0;
    const size_t in_w = mode == 2 ? 130 : 128, out_w = mode == 2 ? 128 : 130;
    int16_t *d_in = nullptr, *d_out = nullptr;
    int32_t *d_scale = nullptr;
    cudaError_t e = cudaMalloc(&d_in, (size_t)n * in_w * sizeof(int16_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)n * out_w * sizeof(int16_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_scale, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(d_in, in, (size_t)n * in_w * sizeof(int16_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_out, 0, (size_t)n * out_w * sizeof(int16_t));
    if (e == cudaSuccess) e = cudaMemset(d_scale, 0, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess) {
        const int vectors_per_block = mode == 1 ? 8 : 4;
        aecm_fft_selftest_kernel<<<(n + vectors_per_block - 1) / vectors_per_block, 128>>>(mode, d_in, d_out, d_scale, n);
        g_launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)n * out_w * sizeof(int16_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && scale) e = cudaMemcpy(scale, d_scale, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    cudaFree(d_scale);
    return (int)e;
}

}  // extern "C"
