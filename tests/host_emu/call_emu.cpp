// call_emu.cpp -- TEST INFRASTRUCTURE.  Compiles the DEVICE code of the hot path (aecm_frame_api.cuh -> aecm_block.cuh
// -> aecm_prims.cuh -> aecm_fft.cuh, plus aecm_init.cuh) for the host against warp_emu.h and drives one instance through
// whole streams of 10 ms API calls exactly as aecm_call_kernel does for one warp: the record lives in ordinary memory
// instead of shared memory, the far history and the far-end jitter ring in ordinary memory instead of HBM, the TMA
// copies are not needed.  Everything else -- every line of arithmetic and every warp-level operation -- is the code the
// GPU runs.  tests/test_call_path_emulated_warp.py compares the output with the oracle / golden streams.
#define AECM_HOST_EMU 1
#include "warp_emu.h"

#include "../../webrtc_aecm_b200/csrc/aecm_frame_api.cuh"
#include "../../webrtc_aecm_b200/csrc/aecm_init.cuh"

namespace {

struct Emu {
    alignas(16) AecmInstanceState st;
    alignas(16) WarpScratch sc;
    alignas(16) AecmDelayExt ext;
    uint16_t far_hist[AECM_MAX_DELAY * AECM_HIST_ROW];
    int16_t far_ring[AECM_FAR_RING];
};

// the 65-entry channel estimate an InitEchoPath event imports (tests/test_fuzz_api.event_path is the same rule in numpy)
void emu_event_path(int seed, int16_t *path) {
    for (int i = 0; i < AECM_BINS; i++) path[i] = (int16_t)(((uint32_t)seed * 1103515245u + 12345u + (uint32_t)i * 7919u) & 0x3fffu);
}

}  // namespace

extern "C" {

size_t emu_state_bytes(void) { return sizeof(AecmInstanceState); }

// the scalar primitives of aecm_prims.cuh, one value at a time (known-answer tests against the reference's SPL functions)
int emu_div_by_bin(int n, int d) { return div_by_bin(n, d); }
int emu_sqrt_floor(int v) { return sqrt_floor(v); }
int emu_norm_w32(int a) { return norm_w32(a); }
int emu_norm_u32(unsigned a) { return norm_u32(a); }
int emu_norm_w16(int a) { return norm_w16(a); }
int emu_log_energy_q8(unsigned e, int q) { return log_energy_q8(e, q); }
int emu_add_sat32(int a, int b) { return add_sat32(a, b); }
// the comfort-noise generator as the device draws it (aecm_block.cuh: lane-parallel jump-ahead): out[k] = the (k+1)-th number
// of WebRtcSpl_RandUArray started at `seed`, k = 0..63; returns the seed after 64 draws
uint32_t emu_lcg_draws(uint32_t seed, int16_t *out) {
    for (int k = 0; k < 64; k++) out[k] = (int16_t)(((kLcgJump.mul[k] * seed + kLcgJump.add[k]) & 0x7fffffffu) >> 16);
    return (kLcgJump.mul[63] * seed + kLcgJump.add[63]) & 0x7fffffffu;
}

// One instance, n_calls API calls of `spc` samples.  ms: per-call msInSndCardBuf [n_calls].  far_ops: per call how
// many times BufferFarend runs before Process (0 = underrun, 1 = normal, 2 = overrun).  Returns the OR of all return
// codes; *state_out (optional) receives the final record.
// spc_calls (optional): per-call frame sizes.  events / event_rc / quality_trace (optional): see the loop below.
// dly (optional): {use, robust, allowed_offset, max_lookahead, lookahead, soft_reset_every, soft_reset_shift}: with
// use != 0 the kExt instantiation runs, the options are applied after Init/set_config as AecmB200_SetDelayEstimatorOptions
// does, and the delay estimator is soft-reset before the calls with (call % every) == every - 1, the sign of the shift
// alternating (the schedule of oracle/aecm_runner.inc).  delay_trace (optional) [n_calls]: last_delay after every call.
int emu_run_stream(int fs, int cng_mode, int echo_mode, const int16_t *far, const int16_t *near_noisy, const int16_t *near_clean,
                   int16_t *out, int n_calls, int spc_all, const int16_t *ms, const uint8_t *far_ops, void *state_out,
                   long *barriers, const int32_t *dly, int32_t *delay_trace, const int32_t *spc_calls, const int32_t *events,
                   int n_events, int32_t *event_rc, float *quality_trace) {
    Emu *e = new Emu();
    memset(e, 0, sizeof(*e));
    int acc = 0;
    const bool ext = dly && dly[0];
    warp_emu::g_barriers = 0;
    warp_emu::run([&](int lane) {
        init_instance(&e->st, fs, lane);
        init_delay_ext(&e->ext, lane);
        __syncwarp();
        if (lane == 0) apply_config(e->st.s, cng_mode, echo_mode);
        __syncwarp();
        if (ext) {
            const int r = set_delay_options(&e->st, &e->ext, dly[1], dly[2], dly[3], dly[4], lane);
            if (lane == 0) acc |= r;
            __syncwarp();
        }
        Ctx c;
        c.st = &e->st;
        c.sc = &e->sc;
        c.state_base = &e->st;  // the "HBM" record is the record itself here: the write-through stores hit the same words
        c.far_hist_base = e->far_hist;
        c.far_ring_base = e->far_ring;
        c.inst = 0;
        c.ext = &e->ext;
        c.lane = lane;
        size_t o = 0;  // spc_calls (optional): the frame size of every call (80 or 160), overriding spc
        for (int call = 0; call < n_calls; o += (size_t)(spc_calls ? spc_calls[call] : spc_all), call++) {
            const int spc = spc_calls ? spc_calls[call] : spc_all;
            int r = 0;
            if (ext && dly[5] > 0 && (call % dly[5]) == dly[5] - 1) {
                soft_reset_delay_estimator(&e->st, &e->ext, ((call / dly[5]) & 1) ? -dly[6] : dly[6], lane);
                __syncwarp();
            }
            // events (optional, sorted by call): {call, kind, a, b, c, d} applied BEFORE that call; kind 0 = set_delay_options(a, b, c, d),
            // kind 1 = soft_reset_delay_estimator(a), kind 2 = set_config(cngMode a, echoMode b), kind 3 = Init(fs a),
            // kind 4 = InitEchoPath(emu_event_path(a)); event_rc[k] receives what the device function returned
            for (int k = 0; k < n_events; k++) {
                const int32_t *ev = events + 6 * k;
                if (ev[0] != call) continue;
                int er = 0;
                if (ev[1] == 0) {
                    er = set_delay_options(&e->st, &e->ext, ev[2], ev[3], ev[4], ev[5], lane);
                } else if (ev[1] == 1) {
                    er = soft_reset_delay_estimator(&e->st, &e->ext, ev[2], lane);
                } else if (ev[1] == 2) {  // aecm_set_config_kernel (one thread per instance there)
                    if (lane == 0) er = set_config_checked(e->st.s, ev[2], ev[3]);
                } else if (ev[1] == 3) {  // AecmB200_Init: aecm_init_kernel + the memsets of the two HBM planes
                    init_instance(&e->st, ev[2], lane);
                    init_delay_ext(&e->ext, lane);
                    __syncwarp();
                    if (lane == 0) {
                        memset(e->far_hist, 0, sizeof(e->far_hist));
                        memset(e->far_ring, 0, sizeof(e->far_ring));
                    }
                } else if (ev[1] == 4) {  // aecm_init_echo_path_kernel with the path emu_event_path(seed) below
                    int16_t path[AECM_BINS];
                    emu_event_path(ev[2], path);
                    init_echo_path_core(&e->st, path, lane);
                }
                if (lane == 0 && event_rc) event_rc[k] = er;
                __syncwarp();
            }
            for (int rep = 0; rep < (far_ops ? far_ops[call] : 1); rep++) r |= api_buffer_farend(c, far + o, spc);
            if (ext) r |= api_process<true>(c, near_noisy + o, near_clean ? near_clean + o : nullptr, out + o, spc, ms ? ms[call] : 40);
            else r |= api_process<false>(c, near_noisy + o, near_clean ? near_clean + o : nullptr, out + o, spc, ms ? ms[call] : 40);
            if (lane == 0) {
                acc |= r;
                if (delay_trace) delay_trace[call] = e->st.s.last_delay;
                if (quality_trace) quality_trace[call] = last_delay_quality(&e->st, &e->ext);
            }
        }
    });
    if (state_out) memcpy(state_out, &e->st, sizeof(e->st));
    if (barriers) *barriers = warp_emu::g_barriers;
    delete e;
    return acc;
}

}  // extern "C"
