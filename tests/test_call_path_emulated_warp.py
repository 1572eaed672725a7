"""CPU check of DEVICE code, whole call path: the headers the GPU kernel is made of (aecm_frame_api.cuh ->
aecm_block.cuh -> aecm_prims.cuh -> aecm_fft.cuh, aecm_init.cuh) are compiled for the host against
tests/host_emu/warp_emu.h (a warp = 32 cooperative fibers, warp-level operations = rendezvous points) and driven
through whole streams of WebRtcAecm_BufferFarend / WebRtcAecm_Process calls.  Output, return codes and the final
instance record must equal the oracle's bit for bit -- the same comparison the GPU tests make, without a GPU, so the
CPU suite already catches an arithmetic slip in the kernel source."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from cases import CASE_BY_NAME, CASES, DLY_CONFIGS, dly_inputs, far_mask, make_inputs, run_kwargs
from webrtc_aecm_b200.synth import synth_streams

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "webrtc_aecm_b200", "csrc")
SRC = os.path.join(HERE, "host_emu", "call_emu.cpp")
LIB = os.path.join(HERE, "host_emu", "libcall_emu.so")
DEPS = [SRC, os.path.join(HERE, "host_emu", "warp_emu.h")] + [os.path.join(CSRC, f) for f in (
    "aecm_frame_api.cuh", "aecm_block.cuh", "aecm_prims.cuh", "aecm_fft.cuh", "aecm_init.cuh", "aecm_state.h", "aecm_tables.h")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(HERE, "..", "include"),
                               "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.emu_state_bytes.restype = C.c_size_t
    lib.emu_run_stream.restype = C.c_int
    return lib


def emu_run(lib, case, far, near, clean, ms, instance, want_state=False, misalign=False, dly=None, soft_reset=None,
            delay_trace=None):
    """One instance of a case through the emulated warp -> (out, or-of-return-codes, state bytes or None).
    dly / soft_reset as in oracle_lib.run_streams (the kExt instantiation); delay_trace: int32 [calls] to fill."""
    spc = case["spc"]
    calls = far.shape[1] // spc
    if ms is None:
        ms_calls = np.full(calls, 40, dtype=np.int16)
    elif np.ndim(ms) == 2:
        ms_calls = np.ascontiguousarray(ms[instance], dtype=np.int16)
    else:
        ms_calls = np.full(calls, int(np.resize(ms, far.shape[0])[instance]), dtype=np.int16)
    ops = np.ones(calls, dtype=np.uint8)
    if case["skip_far_every"]:
        ops[far_mask(case, calls)[:, 0] == 0] = 0
    k = case["extra_far_every"]
    if k:
        ops[(np.arange(calls) % k == k - 1) & (ops > 0)] = 2
    def plane(a):  # misalign: the first sample sits at an odd int16 position, so 32-bit movers cannot be used
        if not misalign:
            return np.ascontiguousarray(a)
        buf = np.zeros(a.size + 2, dtype=np.int16)
        off = 1 if buf.ctypes.data % 4 == 0 else 2
        buf[off:off + a.size] = a
        return buf[off:off + a.size]

    f, x = plane(far[instance]), plane(near[instance])
    cl = None if clean is None else plane(clean[instance])
    out = plane(np.zeros_like(near[instance]))
    state = np.zeros(lib.emu_state_bytes(), dtype=np.uint8) if want_state else None
    nbar = C.c_long(0)
    dly_arr = None
    if dly is not None or soft_reset is not None:
        d, sr = dly or {}, soft_reset or (0, 0)
        dly_arr = np.array([1, d.get("robust", 0), d.get("allowed_offset", 0), d.get("max_lookahead", -1), d.get("lookahead", -1),
                            sr[0], sr[1]], dtype=np.int32)
    acc = lib.emu_run_stream(case["fs"], case["cng"], case["echo_mode"], f.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
                             None if cl is None else cl.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), calls, spc,
                             ms_calls.ctypes.data_as(C.c_void_p), ops.ctypes.data_as(C.c_void_p),
                             None if state is None else state.ctypes.data_as(C.c_void_p), C.byref(nbar),
                             None if dly_arr is None else dly_arr.ctypes.data_as(C.c_void_p),
                             None if delay_trace is None else delay_trace.ctypes.data_as(C.c_void_p), None, None, 0, None, None)
    return out, acc, state


# every case family; two instances each (the GPU suite runs all instances of all cases)
EMU_CASES = [c["name"] for c in CASES]


@pytest.mark.parametrize("name", EMU_CASES)
def test_device_code_on_emulated_warp_matches_oracle(emu, name):
    case = CASE_BY_NAME[name]
    far, near, clean, ms = make_inputs(case)
    want, ret, _ = ol.run_streams(far, near, impl="oracle", clean=clean, ms=ms, **run_kwargs(case))
    n = far.shape[0]
    for instance in sorted({1 % n, n - 1}):
        out, acc, _ = emu_run(emu, case, far, near, clean, ms, instance)
        bad = np.nonzero(out != want[instance])[0]
        assert bad.size == 0, f"instance {instance}: first differing sample {bad[0]} (call {bad[0] // case['spc']})"
        assert (acc != 0) == (ret[instance] != 0)


@pytest.mark.parametrize("name", ["16k_ms_mix", "8k_far_underrun_ms_mix", "16k_clean_path"])
def test_planes_that_are_not_32_bit_aligned(emu, name):
    """Audio rows at odd int16 offsets (odd row strides, sub-ranges): the sample movers fall back from 32-bit words to
    16-bit accesses and must produce the same stream."""
    case = CASE_BY_NAME[name]
    far, near, clean, ms = make_inputs(case)
    want, _, _ = ol.run_streams(far, near, impl="oracle", clean=clean, ms=ms, **run_kwargs(case))
    out, _, _ = emu_run(emu, case, far, near, clean, ms, 0, misalign=True)
    np.testing.assert_array_equal(out, want[0])


def test_final_record_equals_oracle_state_field_by_field(emu):
    """The record the emulated warp leaves behind is the device's AecmInstanceState: compare it with the oracle's state
    through the same field map the GPU state test uses."""
    from webrtc_aecm_b200 import api

    case = CASE_BY_NAME["16k_stress"]
    far, near, clean, ms = make_inputs(case)
    spc, calls = case["spc"], 300
    far, near = far[:, :calls * spc], near[:, :calls * spc]
    out, acc, state = emu_run(emu, case, far, near, None, None, 2, want_state=True)
    o = ol.OracleInstance(case["fs"], case["cng"], case["echo_mode"])
    for c in range(calls):
        o.buffer_farend(far[2, c * spc:(c + 1) * spc])
        rc, w = o.process(near[2, c * spc:(c + 1) * spc], None, 40)
        np.testing.assert_array_equal(out[c * spc:(c + 1) * spc], w)
    dev = api.InstanceState.from_buffer_copy(state.tobytes())
    assert ol.compare_states(dev, o.state()) == []
    o.close()


@pytest.mark.parametrize("cfg", DLY_CONFIGS, ids=[c["id"] for c in DLY_CONFIGS])
def test_delay_estimator_options_on_emulated_warp(emu, cfg):
    """SURVEY 8-f4 on the device code (kExt instantiation: robust validation with its float histogram, near-end lookahead,
    soft resets), run on the emulated warp: output AND the delay estimate after every call equal the oracle's, which
    tests/test_oracle.py pins to the reference driven through its own setters."""
    fs, spc, far, near, ms, kw = dly_inputs(cfg)
    case = dict(CASE_BY_NAME["c5_delay_sweep_0_120ms"], fs=fs, spc=spc)
    calls = far.shape[1] // spc
    want, ret, _, trace = ol.run_streams(far, near, impl="oracle", ms=ms, dly=cfg["dly"], soft_reset=cfg["soft_reset"],
                                         delay_trace=True, **kw)
    for instance in cfg["instances"]:
        got_trace = np.zeros(calls, dtype=np.int32)
        out, acc, _ = emu_run(emu, case, far, near, None, ms, instance, dly=cfg["dly"], soft_reset=cfg["soft_reset"],
                              delay_trace=got_trace)
        bad = np.nonzero(out != want[instance])[0]
        assert bad.size == 0, f"instance {instance}: first differing sample {bad[0]} (call {bad[0] // spc})"
        np.testing.assert_array_equal(got_trace, trace[instance])
        assert acc == 0 and ret[instance] == 0


@pytest.mark.parametrize("seed,fs", [(s, f) for s, f in __import__("test_fuzz_api").SCRIPTS if s % 3 != 2][:14] + __import__("test_fuzz_api").MUTANT_KILLERS)
def test_random_call_sequences_on_emulated_warp(emu, seed, fs):
    """The randomised API scripts of tests/test_fuzz_api.py (mixed 80/160-sample calls, msInSndCardBuf jumps and clamps,
    under- and overruns, clean plane or not, all echo modes) through the device code on the emulated warp: output and the OR
    of the return codes equal the oracle's, which the same scripts pin to the reference."""
    import test_fuzz_api as fz

    s = fz.make_script(seed, 500, fs)
    want, rcs = fz.run_oracle(s)
    n_calls = len(s["n"])
    out = np.zeros_like(s["near"])
    ms16 = np.ascontiguousarray(s["ms"].astype(np.int16))
    spc_calls = np.ascontiguousarray(s["n"].astype(np.int32))
    ops = np.ascontiguousarray(s["far_ops"])
    cl = np.ascontiguousarray(s["cln"]) if s["clean"].all() else None
    far, near = np.ascontiguousarray(s["far"]), np.ascontiguousarray(s["near"])
    nbar = C.c_long(0)
    acc = emu.emu_run_stream(fs, s["cng"], s["echo_mode"], far.ctypes.data_as(C.c_void_p), near.ctypes.data_as(C.c_void_p),
                             None if cl is None else cl.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n_calls, 0,
                             ms16.ctypes.data_as(C.c_void_p), ops.ctypes.data_as(C.c_void_p), None, C.byref(nbar), None, None,
                             spc_calls.ctypes.data_as(C.c_void_p), None, 0, None, None)
    bad = np.nonzero(out != want)[0]
    assert bad.size == 0, f"first differing sample {bad[0]}"
    assert acc == int(np.bitwise_or.reduce(rcs))


@pytest.mark.parametrize("seed,fs", [(s, f) for s, f in __import__("test_fuzz_api").SCRIPTS if s % 2 == 0][:10])
def test_random_control_calls_in_mid_stream_on_emulated_warp(emu, seed, fs):
    """set_config (valid and refused), Init (same or other rate) and InitEchoPath at random points of the random scripts through
    the device functions the control kernels run (set_config_checked, init_instance + init_delay_ext, init_echo_path_core):
    return values, output and the OR of the return codes equal the oracle's, which test_fuzz_api.py pins to the reference."""
    import test_fuzz_api as fz

    s = fz.make_script(seed, 500, fs)
    events = fz.make_control_events(seed, 500)
    want_ev = []
    want, rcs = fz.run_oracle(s, events, want_ev)
    ev = np.ascontiguousarray(np.array([(c, k, a, b, 0, 0) for c, k, a, b in events], dtype=np.int32))
    got_ev = np.full(len(events), 99, dtype=np.int32)
    out = np.zeros_like(s["near"])
    ms16 = np.ascontiguousarray(s["ms"].astype(np.int16))
    spc_calls = np.ascontiguousarray(s["n"].astype(np.int32))
    ops = np.ascontiguousarray(s["far_ops"])
    cl = np.ascontiguousarray(s["cln"]) if s["clean"].all() else None
    far, near = np.ascontiguousarray(s["far"]), np.ascontiguousarray(s["near"])
    nbar = C.c_long(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    acc = emu.emu_run_stream(fs, s["cng"], s["echo_mode"], vp(far), vp(near), None if cl is None else vp(cl), vp(out), len(s["n"]), 0,
                             vp(ms16), vp(ops), None, C.byref(nbar), None, None, vp(spc_calls), vp(ev), len(events), vp(got_ev), None)
    assert got_ev.tolist() == want_ev and len(events) >= 3
    bad = np.nonzero(out != want)[0]
    assert bad.size == 0, f"first differing sample {bad[0]}"
    assert acc == int(np.bitwise_or.reduce(rcs))


@pytest.mark.parametrize("fs,base_ms", [(16000, 20), (8000, 12), (16000, 29), (16000, 25)])
def test_startup_stability_check_at_exactly_the_tolerance_on_emulated_warp(emu, fs, base_ms):
    """tests/test_fuzz_api.tolerance_script (deviation == tolerance during start-up) through the device code; a tolerance of 9
    instead of 8 in aecm_frame_api.cuh survives everything else in the suite."""
    import test_fuzz_api as fz

    s = fz.tolerance_script(fs, base_ms)
    want, rcs = fz.run_oracle(s)
    out = np.zeros_like(s["near"])
    ms16 = np.ascontiguousarray(s["ms"].astype(np.int16))
    spc_calls = np.ascontiguousarray(s["n"].astype(np.int32))
    ops = np.ascontiguousarray(s["far_ops"])
    cl = np.ascontiguousarray(s["cln"]) if s["clean"].all() else None
    far, near = np.ascontiguousarray(s["far"]), np.ascontiguousarray(s["near"])
    nbar = C.c_long(0)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    acc = emu.emu_run_stream(fs, s["cng"], s["echo_mode"], vp(far), vp(near), vp(cl), vp(out), len(s["n"]), 0, vp(ms16), vp(ops), None, C.byref(nbar),
                             None, None, vp(spc_calls), None, 0, None, None)
    np.testing.assert_array_equal(out, want)
    assert acc == int(np.bitwise_or.reduce(rcs))


@pytest.mark.parametrize("seed", range(200, 208))
def test_random_delay_estimator_option_changes_on_emulated_warp(emu, seed):
    """SURVEY 8-f4 device functions under random use: the estimator's setters, soft resets by random shifts and re-creations
    with a new max_lookahead hit the device code (set_delay_options, soft_reset_delay_estimator, process_binary_spectrum<true>)
    at random calls; output, every setter's return value, the delay estimate and its float quality after every call equal the
    oracle's (which tests/test_fuzz_api.py pins to the reference under the same kind of fire)."""
    rng = np.random.default_rng(seed)
    fs, spc, calls = 16000, 160, 500
    far, near = synth_streams(1, calls * spc, fs, seed=seed, delay_ms=int(rng.integers(0, 115)))
    far, near = np.ascontiguousarray(far[0]), np.ascontiguousarray(near[0])
    max_la = int(rng.integers(0, 9))
    events = [(0, 0, int(rng.integers(0, 2)), int(rng.integers(0, 4)), max_la, -1)]
    for c in range(1, calls):
        u = rng.random()
        if u < 0.02:
            events.append((c, 0, int(rng.integers(0, 2)), int(rng.integers(0, 5)), -1, int(rng.integers(-1, max_la + 3))))
        elif u < 0.035:
            events.append((c, 1, int(rng.integers(-30, 31)), 0, 0, 0))
        elif u < 0.04:
            max_la = int(rng.integers(0, 12))
            events.append((c, 0, int(rng.integers(0, 2)), 0, max_la, -1))
    events += [(50, 1, 1, 0, 0, 0), (51, 1, -1, 0, 0, 0), (52, 1, 0, 0, 0, 0)]  # the smallest shifts: one block forward, one back, none
    events.sort(key=lambda e: e[0])
    ev = np.ascontiguousarray(np.array(events, dtype=np.int32))
    # the oracle, stepped call by call with the same events
    o = ol.OracleInstance(fs, 1, 3)
    soft = ol.oracle().oracle_aecm_soft_reset_delay_estimator
    soft.argtypes = [C.c_void_p, C.c_int]
    want = np.zeros_like(near)
    want_rc = np.zeros(len(events), dtype=np.int32)
    want_d, want_q = np.zeros(calls, dtype=np.int32), np.zeros(calls, dtype=np.float32)
    k = 0
    for c in range(calls):
        while k < len(events) and events[k][0] == c:
            e = events[k]
            want_rc[k] = o.set_delay_estimator_options(e[2], e[3], e[4], e[5]) if e[1] == 0 else soft(o.h, e[2])
            k += 1
        o.buffer_farend(far[c * spc:(c + 1) * spc])
        rc, w = o.process(near[c * spc:(c + 1) * spc], None, 40)
        want[c * spc:(c + 1) * spc] = w
        want_d[c], want_q[c] = o.delay_estimate()
    o.close()
    out = np.zeros_like(near)
    got_rc = np.full(len(events), 99, dtype=np.int32)
    got_d, got_q = np.zeros(calls, dtype=np.int32), np.zeros(calls, dtype=np.float32)
    dly = np.array([1, 0, 0, -1, -1, 0, 0], dtype=np.int32)  # the kExt instantiation, nothing set up front
    nbar = C.c_long(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    acc = emu.emu_run_stream(fs, 1, 3, vp(far), vp(near), None, vp(out), calls, spc, None, None, None, C.byref(nbar), vp(dly),
                             vp(got_d), None, vp(ev), len(events), vp(got_rc), vp(got_q))
    assert acc == 0
    np.testing.assert_array_equal(got_rc, want_rc)
    bad = np.nonzero(out != want)[0]
    assert bad.size == 0, f"first differing sample {bad[0]} (call {bad[0] // spc})"
    np.testing.assert_array_equal(got_d, want_d)
    np.testing.assert_array_equal(got_q, want_q)
    assert len(events) >= 8
