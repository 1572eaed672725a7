// aecm_frame_api.cuh -- L2 frame->block FIFOs (aecm_core.cc:501-572) and L3 API wrapper (echo_control_mobile.cc:195-408, 534-594)
// (part of the single device translation unit aecm_kernels.cu; see that file's header for the mapping)
#ifndef AECM_FRAME_API_CUH_
#define AECM_FRAME_API_CUH_

#include "aecm_block.cuh"

namespace {

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

// CTA-wide rendezvous points that keep the warps of an SM on the same stretch of code (the call path is 110 KB of
// straight-line SASS against a 32 KB instruction cache).  They sit at fixed places of a call -- before every frame and
// before every potential block -- so every warp of the CTA passes the same number of them per call; a warp whose instance
// does nothing in this call (un-initialised, start-up phase) or that has no row walks through them with idle_call().
__device__ __forceinline__ void cta_rendezvous_frame() {
    if (kSyncPoints >= 1) __syncthreads();
}
__device__ __forceinline__ void cta_rendezvous_block() {
    if (kSyncPoints >= 2) __syncthreads();
}
__device__ __forceinline__ void idle_call(int n_frames) {
    if (kSyncPoints >= 1)
        for (int i = 0; i < n_frames; i++) {
            cta_rendezvous_frame();
            cta_rendezvous_block();
            cta_rendezvous_block();
        }
}

// ------------------------------------------------------------------------------------------------
// L2: one 80-sample frame (aecm_core.cc:501-572).  The far frame comes from the HBM jitter FIFO (far_idx >= 0:
// index of its first sample) or from the replay buffer farend_old[slot]; near/clean come from the input planes.
// ------------------------------------------------------------------------------------------------
template <bool kExt>
__device__ __forceinline__ void process_frame(const Ctx &c, int far_idx, int slot, const int16_t *near80,
                                              const int16_t *clean80, int16_t *out80) {
    AecmInstanceState *st = c.st;
    AecmScalars &S = st->s;
    const int lane = c.lane;
    // FIFO cursors are read into registers (warp-uniform), updated there, and committed by lane 0
    {
        const int in_rd = S.frm_in_rd, clean_rd = S.frm_clean_rd;
        int in_cnt = S.frm_in_cnt, clean_cnt = S.frm_clean_cnt;
        // WebRtc_WriteBuffer x2/x3, :520-524 (a full FIFO drops the excess, ring_buffer.c:150-152)
        const int n = min(AECM_FRAME_LEN, AECM_FRM_RING - in_cnt);
        const int wr = wrap_add(in_rd, in_cnt, AECM_FRM_RING);
        if (n == AECM_FRAME_LEN && ((wr | far_idx) & 1) == 0 && word_aligned(near80)) {
            const uint32_t *nw = reinterpret_cast<const uint32_t *>(near80);
            const uint32_t *ring = reinterpret_cast<const uint32_t *>(c.far_ring());
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
                const int16_t f = far_idx >= 0 ? c.far_ring()[wrap_add(far_idx, k, AECM_FAR_RING)] : st->farend_old[slot][k];
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
        cta_rendezvous_block();
        if (S.frm_in_cnt >= AECM_PART_LEN) {
            {
                const int o_rd = S.frm_out_rd, o_cnt = S.frm_out_cnt;
                process_block<kExt>(c, S.frm_in_rd, S.frm_clean_rd, clean80 != nullptr, wrap_add(o_rd, o_cnt, AECM_FRM_RING),
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
        uint32_t *ring = reinterpret_cast<uint32_t *>(c.far_ring());
        for (int k = c.lane; k < (w >> 1); k += 32) ring[wrap_add(wr >> 1, k, AECM_FAR_RING / 2)] = src[k];
    } else {
        for (int k = c.lane; k < w; k += 32) c.far_ring()[wrap_add(wr, k, AECM_FAR_RING)] = far[k];
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

template <bool kExt>
__device__ __forceinline__ int api_process(const Ctx &c, const int16_t *near_noisy, const int16_t *near_clean,
                                           int16_t *out, int n, int ms) {
    AecmScalars &S = c.st->s;
    const int lane = c.lane;
    const int n_frames = n / AECM_FRAME_LEN;
    int ret = 0;
    if (S.init_flag != AECM_INIT_MAGIC) {  // :264
        idle_call(n_frames);
        return AECM_UNINITIALIZED_ERROR;
    }
    {
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
            idle_call(n_frames);
            return ret;
        }
    }
    // aecm->msInSndCardBuf = msInSndCardBuf + 10 (:280): kept in the record from here on, not in a register across the frames
    __syncwarp();
    if (lane == 0) S.ms_in_snd_card_buf = ms;
    __syncwarp();
    for (int i = 0; i < n_frames; i++) {  // :358-397
        cta_rendezvous_frame();
        int far_idx = -2;  // "replay farend_old" (even, so that the word-wide mover still applies)
        {  // wrapper bookkeeping of this frame on register copies, committed before the frame is processed
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
            if ((i == 0 && fs == 8000) || (i == 1 && fs == 16000)) est_buf_delay(d, S.ms_in_snd_card_buf, S.mult);
            __syncwarp();
            if (lane == 0) {
                S.far_rd = d.far_rd; S.far_cnt = d.far_cnt;
                S.filt_delay = d.filt_delay; S.wrapper_known_delay = d.known_delay;
                S.last_delay_diff = d.last_delay_diff; S.time_for_delay_change = d.time_for_delay_change;
            }
            __syncwarp();
        }
        process_frame<kExt>(c, far_idx, i, near_noisy + AECM_FRAME_LEN * i,
                      near_clean ? near_clean + AECM_FRAME_LEN * i : nullptr, out + AECM_FRAME_LEN * i);
    }
    return ret;
}

}  // namespace

#endif  // AECM_FRAME_API_CUH_
