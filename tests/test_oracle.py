"""Pins the CPU oracle (oracle/aecm_oracle.c): against the committed golden fixtures generated from the
unmodified reference (tests/golden/, made by tools/make_golden.py) and, when oracle/_ref is present, against
the reference library itself.  CPU only."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from cases import CASE_BY_NAME, CASES, DLY_CONFIGS, dly_inputs, make_inputs, run_kwargs
from webrtc_aecm_b200.shard import stream_checksum

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STREAMS = json.load(open(os.path.join(GOLD, "streams.json")))["cases"]
VEC = np.load(os.path.join(GOLD, "vectors.npz"))


def run_case(case, impl):
    far, near, clean, ms = make_inputs(case)
    out, ret, _ = ol.run_streams(far, near, impl=impl, clean=clean, ms=ms, **run_kwargs(case))
    return far, near, out, ret


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_golden_streams(case):
    far, near, out, ret = run_case(case, "oracle")
    g = STREAMS[case["name"]]
    assert stream_checksum(np.concatenate([far, near])) == g["inputs_sha256"], "synthetic inputs drifted"
    assert list(out.shape) == g["shape"]
    assert [int(r) for r in ret] == g["ret_or"]
    assert stream_checksum(out) == g["sha256"]
    key = "out__" + case["name"]
    if key in VEC.files:
        np.testing.assert_array_equal(out, VEC[key])


@pytest.mark.skipif(ol.ref_path() is None, reason="oracle/_ref/libaecm_ref.so not built")
@pytest.mark.parametrize("case", CASES[::3], ids=[c["name"] for c in CASES[::3]])
def test_oracle_matches_reference_library(case):
    _, _, out, ret = run_case(case, "oracle")
    _, _, ref_out, ref_ret = run_case(case, "ref")
    np.testing.assert_array_equal(out, ref_out)
    np.testing.assert_array_equal(ret, ref_ret)


def test_fft_forward_known_answers():
    lib = ol.oracle()
    fin, fout = VEC["fft_fwd_in"], VEC["fft_fwd_out"]
    got = np.zeros_like(fout)
    for i in range(len(fin)):
        lib.oracle_real_fft128_forward(fin[i].ctypes.data_as(C.c_void_p), got[i].ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(got, fout)


def test_fft_inverse_known_answers_and_scale():
    lib = ol.oracle()
    iin, iout, iscale = VEC["fft_inv_in"], VEC["fft_inv_out"], VEC["fft_inv_scale"]
    got = np.zeros_like(iout)
    scale = np.zeros_like(iscale)
    for i in range(len(iin)):
        scale[i] = lib.oracle_real_fft128_inverse(iin[i].ctypes.data_as(C.c_void_p), got[i].ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(scale, iscale)
    np.testing.assert_array_equal(got, iout)
    assert len(set(iscale.tolist())) >= 4 and iscale.min() == 0  # the vectors exercise the block-floating scaling


def test_fft_adversarial_family_known_answers():
    """The oracle's FFT/IFFT on the larger adversarial family (tests/cases.py:fft_adversarial_vectors), against the
    reference's answers in tests/golden/fft_adversarial.npz; and the committed inputs are what cases.py generates."""
    from cases import fft_adversarial_vectors

    lib = ol.oracle()
    A = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fft_adversarial.npz"))
    fwd, inv = fft_adversarial_vectors()
    np.testing.assert_array_equal(fwd, A["fwd_in"])
    np.testing.assert_array_equal(inv, A["inv_in"])
    got = np.zeros_like(A["fwd_out"])
    for i in range(len(fwd)):
        lib.oracle_real_fft128_forward(fwd[i].ctypes.data_as(C.c_void_p), got[i].ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(got, A["fwd_out"])
    got, sc = np.zeros_like(A["inv_out"]), np.zeros(len(inv), dtype=np.int32)
    for i in range(len(inv)):
        sc[i] = lib.oracle_real_fft128_inverse(inv[i].ctypes.data_as(C.c_void_p), got[i].ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(sc, A["inv_scale"])
    np.testing.assert_array_equal(got, A["inv_out"])
    assert set(A["inv_scale"].tolist()) >= {0, 1, 2, 3, 4}


def test_sqrt_floor_and_lcg_known_answers():
    lib = ol.oracle()
    got = np.array([lib.oracle_sqrt_floor(int(v)) for v in VEC["sqrt_in"]], dtype=np.int32)
    np.testing.assert_array_equal(got, VEC["sqrt_out"])
    for s, want, after in zip(VEC["rand_seed"], VEC["rand_out"], VEC["rand_seed_after"]):
        buf = np.zeros(64, dtype=np.int16)
        seed = C.c_uint32(int(s))
        lib.oracle_rand_u_array(buf.ctypes.data_as(C.c_void_p), 64, C.byref(seed))
        np.testing.assert_array_equal(buf, want)
        assert seed.value == int(after)


def test_error_codes_and_call_protocol():
    lib = ol.oracle()
    h = lib.oracle_aecm_create()
    z = np.zeros(160, dtype=np.int16)
    p = z.ctypes.data_as(C.c_void_p)
    # un-initialised instance (echo_control_mobile.cc:206, 264, 417)
    assert lib.oracle_aecm_buffer_farend(h, p, 160) == 12002
    assert lib.oracle_aecm_process(h, p, None, p, 160, 0) == 12002
    assert lib.oracle_aecm_set_config(h, 1, 3) == 12002
    assert lib.oracle_aecm_init(h, 44100) == 12004
    assert lib.oracle_aecm_init(h, 16000) == 0
    assert lib.oracle_aecm_buffer_farend(h, None, 160) == 12003
    assert lib.oracle_aecm_buffer_farend(h, p, 100) == 12004
    assert lib.oracle_aecm_process(h, None, None, p, 160, 0) == 12003
    assert lib.oracle_aecm_process(h, p, None, None, 160, 0) == 12003
    assert lib.oracle_aecm_process(h, p, None, p, 128, 0) == 12004
    assert lib.oracle_aecm_set_config(h, 2, 3) == 12004
    assert lib.oracle_aecm_set_config(h, 1, 5) == 12004
    assert lib.oracle_aecm_process(h, p, None, p, 160, -1) == 12100  # warning, still processed
    assert lib.oracle_aecm_process(h, p, None, p, 160, 501) == 12100
    path = np.zeros(65, dtype=np.int16)
    assert lib.oracle_aecm_get_echo_path(h, path.ctypes.data_as(C.c_void_p), 130) == 0
    assert path[0] == 2040 and path[64] == 3153  # kChannelStored16kHz, aecm_core.cc:37-43
    assert lib.oracle_aecm_get_echo_path(h, path.ctypes.data_as(C.c_void_p), 128) == 12004
    assert lib.oracle_aecm_init_echo_path(h, None, 130) == 12003
    lib.oracle_aecm_free(h)


def test_startup_gate_and_output_lag():
    """calls 0-5 are pass-through (ECstartup clears inside call 5); afterwards the near end
    reaches the output with a lag of 112 samples when the far end is silent and CNG is off."""
    fs, spc, calls = 16000, 160, 40
    rng = np.random.default_rng(3)
    near = rng.integers(-3000, 3000, size=(1, calls * spc)).astype(np.int16)
    far = np.zeros_like(near)
    out, ret, _ = ol.run_streams(far, near, fs=fs, impl="oracle", cng_mode=0)
    assert not ret.any()
    np.testing.assert_array_equal(out[0, :6 * spc], near[0, :6 * spc])
    assert not np.array_equal(out[0, 6 * spc:7 * spc], near[0, 6 * spc:7 * spc])
    seg = slice(20 * spc, 30 * spc)
    err = {lag: np.abs(out[0, seg].astype(int) - near[0, seg.start - lag:seg.stop - lag].astype(int)).mean()
           for lag in (0, 64, 96, 112, 128)}
    assert min(err, key=err.get) == 112 and err[112] < 0.05 * err[0]


def test_adversarial_cases_reach_the_wide_shift_corner():
    """The adv_* cases exist to exercise aecm_core.cc:906-912 with |shift2ResChan| > 31 (undefined in C, x86 SAR
    semantics in the reference as built); make sure they really do."""
    from cases import CASE_BY_NAME

    lib = ol.oracle()
    lib.oracle_debug_wide_shift_events.restype = C.c_uint
    hit = 0
    for name in ("16k_adv_gated_tone_dc_near", "8k_adv_gated_tone_dc_near_mode0"):
        before = lib.oracle_debug_wide_shift_events()
        case = CASE_BY_NAME[name]
        far, near, clean, ms = make_inputs(case)
        ol.run_streams(far, near, fs=case["fs"], samples_per_call=case["spc"], impl="oracle", ms=ms, cng_mode=case["cng"],
                       echo_mode=case["echo_mode"], threads=1)
        hit += lib.oracle_debug_wide_shift_events() - before
    assert hit >= 10


@pytest.mark.skipif(ol.ref_path() is None, reason="oracle/_ref/libaecm_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("fs", [16000, 8000])
def test_two_minute_soak_oracle_equals_reference(fs):
    """120 s per instance: the slow state machines (channel store/reset, VAD update halt, delay-change counters) stay
    bit-identical between the restatement and the unmodified reference far beyond the length of the golden cases."""
    from cases import soak_inputs

    far, near = soak_inputs(fs)
    ms = (np.arange(far.shape[0]) * 37) % 300
    a, ra, _ = ol.run_streams(far, near, fs=fs, impl="oracle", ms=ms, echo_mode=3)
    b, rb, _ = ol.run_streams(far, near, fs=fs, impl="ref", ms=ms, echo_mode=3)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ra, rb)


@pytest.mark.skipif(ol.ref_path() is None, reason="oracle/_ref/libaecm_ref.so not built (needs /root/reference)")
def test_live_config_and_echo_path_changes_oracle_equals_reference():
    """Control calls on a running instance (set_config with every mode, InitEchoPath, a second Init): the restatement
    and the reference library, driven call by call through their C APIs, stay bit-identical."""
    from webrtc_aecm_b200.synth import synth_streams

    class Cfg(C.Structure):
        _fields_ = [("cngMode", C.c_int16), ("echoMode", C.c_int16)]

    ref = C.CDLL(ol.ref_path())
    ref.WebRtcAecm_Create.restype = C.c_void_p
    for name, args in (("WebRtcAecm_Init", [C.c_void_p, C.c_int32]), ("WebRtcAecm_Free", [C.c_void_p]),
                       ("WebRtcAecm_BufferFarend", [C.c_void_p, C.c_void_p, C.c_size_t]),
                       ("WebRtcAecm_Process", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int16]),
                       ("WebRtcAecm_set_config", [C.c_void_p, Cfg]), ("WebRtcAecm_InitEchoPath", [C.c_void_p, C.c_void_p, C.c_size_t]),
                       ("WebRtcAecm_GetEchoPath", [C.c_void_p, C.c_void_p, C.c_size_t])):
        getattr(ref, name).argtypes = args
    lib = ol.oracle()
    fs, spc, calls = 8000, 80, 400
    far, near = synth_streams(1, calls * spc, fs, seed=2024, delay_ms=35)
    r = ref.WebRtcAecm_Create()
    o = ol.OracleInstance(fs, 1, 3)
    assert ref.WebRtcAecm_Init(r, fs) == 0
    path = (np.arange(65) * 53 % 3000 + 200).astype(np.int16)
    schedule = {40: (0, 1), 90: (1, 4), 91: (1, 0), 170: (0, 3), 300: (1, 2)}
    for c in range(calls):
        if c in schedule:
            cng, mode = schedule[c]
            assert ref.WebRtcAecm_set_config(r, Cfg(cng, mode)) == lib.oracle_aecm_set_config(o.h, cng, mode) == 0
        if c == 130:
            assert ref.WebRtcAecm_InitEchoPath(r, path.ctypes.data_as(C.c_void_p), 130) == 0
            assert lib.oracle_aecm_init_echo_path(o.h, path.ctypes.data_as(C.c_void_p), 130) == 0
        if c == 250:  # re-Init in the middle of a stream
            assert ref.WebRtcAecm_Init(r, fs) == lib.oracle_aecm_init(o.h, fs) == 0
        f = np.ascontiguousarray(far[0, c * spc:(c + 1) * spc])
        x = np.ascontiguousarray(near[0, c * spc:(c + 1) * spc])
        out = np.zeros_like(x)
        assert ref.WebRtcAecm_BufferFarend(r, f.ctypes.data_as(C.c_void_p), spc) == o.buffer_farend(f) == 0
        rc = ref.WebRtcAecm_Process(r, x.ctypes.data_as(C.c_void_p), None, out.ctypes.data_as(C.c_void_p), spc, 30)
        want_rc, want = o.process(x, None, 30)
        assert rc == want_rc
        np.testing.assert_array_equal(out, want, err_msg=f"call {c}")
    a, b = np.zeros(65, np.int16), np.zeros(65, np.int16)
    assert ref.WebRtcAecm_GetEchoPath(r, a.ctypes.data_as(C.c_void_p), 130) == 0
    assert lib.oracle_aecm_get_echo_path(o.h, b.ctypes.data_as(C.c_void_p), 130) == 0
    np.testing.assert_array_equal(a, b)
    ref.WebRtcAecm_Free(r)
    o.close()


DLY_GOLD = json.load(open(os.path.join(GOLD, "dly_streams.json")))["cases"]


@pytest.mark.parametrize("cfg", DLY_CONFIGS, ids=[c["id"] for c in DLY_CONFIGS])
def test_delay_estimator_options_oracle_equals_reference(cfg):
    """robust validation (float histogram), lookahead and soft resets.  The oracle must reproduce the checksums
    of the reference driven through its own setters (tests/golden/dly_streams.json, written by tools/make_golden.py via
    oracle/ref_shim.cc) -- output and the delay estimate after every call -- and, where the shim library is present, the
    arrays themselves."""
    _, _, far, near, ms, kw = dly_inputs(cfg)
    out, ret, _, trace = ol.run_streams(far, near, impl="oracle", ms=ms, dly=cfg["dly"], soft_reset=cfg["soft_reset"], delay_trace=True, **kw)
    gold = DLY_GOLD[cfg["id"]]
    assert stream_checksum(out) == gold["sha256"] and stream_checksum(trace) == gold["delay_trace_sha256"]
    assert [int(r) for r in ret] == gold["ret_or"]
    if ol.ref_shim_path():
        rout, rret, _, rtrace = ol.run_streams(far, near, impl="ref", ms=ms, dly=cfg["dly"], soft_reset=cfg["soft_reset"],
                                               delay_trace=True, **kw)
        np.testing.assert_array_equal(out, rout)
        np.testing.assert_array_equal(trace, rtrace)
        np.testing.assert_array_equal(ret, rret)


def test_delay_estimator_options_change_the_estimate():
    """The configurations are not vacuous: robust validation, lookahead and soft resets each move the output away from the
    default wiring on these inputs, and the setters refuse what the reference's refuse."""
    base = json.load(open(os.path.join(GOLD, "streams.json")))["cases"]["c5_delay_sweep_0_120ms"]["sha256"]
    shas = {k: v["sha256"] for k, v in DLY_GOLD.items() if k != "robust_long_steady"}  # (that one has its own inputs)
    assert shas["statistics_only"] == base  # collecting the statistics alone changes nothing (aecm_core.cc:225)
    assert len(set(shas.values())) == len(shas) and all(v != base for k, v in shas.items() if k != "statistics_only")
    lib = ol.oracle()
    a = lib.oracle_aecm_create()
    lib.oracle_aecm_init(C.c_void_p(a), 16000)
    f = lib.oracle_aecm_set_delay_estimator_options
    f.argtypes = [C.c_void_p] + [C.c_int] * 4
    assert f(a, 2, 0, -1, -1) == -1 and f(a, 1, -1, -1, -1) == -1  # delay_estimator_wrapper.cc:404, 427
    assert f(a, 0, 0, -1, 1) == -1  # lookahead beyond max_lookahead (0): :388-391
    assert f(a, 1, 2, 6, 4) == 0 and f(a, 1, 2, -1, 7) == -1
    lib.oracle_aecm_soft_reset_delay_estimator.argtypes = [C.c_void_p, C.c_int]
    assert lib.oracle_aecm_soft_reset_delay_estimator(a, 3) == 3  # lookahead 4 -> 1
    assert lib.oracle_aecm_soft_reset_delay_estimator(a, 3) == 1  # 1 -> 0 (clamped)
    assert lib.oracle_aecm_soft_reset_delay_estimator(a, -9) == -6  # 0 -> 6 (clamped to max_lookahead)
    lib.oracle_aecm_free(C.c_void_p(a))


@pytest.mark.skipif(ol.ref_shim_path() is None, reason="oracle/_ref/libaecm_ref_shim.so not built (needs /root/reference)")
@pytest.mark.parametrize("robust", [0, 1])
def test_last_delay_quality_oracle_equals_reference(robust):
    """WebRtc_last_delay / WebRtc_last_delay_quality (delay_estimator_wrapper.cc:498-512, delay_estimator.cc:666-688): a float that
    is the histogram height under robust validation and a function of last_delay_probability otherwise.  The reference
    instance is driven call by call through its own API (the shim library) next to the oracle; both numbers must be
    identical after every 50th call."""
    fs, spc, _, far, near, _, _ = (16000, 160) + dly_inputs(DLY_CONFIGS[1])[1:]
    R = C.CDLL(ol.ref_shim_path())
    R.WebRtcAecm_Create.restype = C.c_void_p
    R.ref_shim_last_delay_quality.restype = C.c_float
    for name in ("WebRtcAecm_Init", "WebRtcAecm_BufferFarend", "WebRtcAecm_Process", "WebRtcAecm_Free",
                 "ref_shim_set_delay_estimator_options", "ref_shim_last_delay", "ref_shim_last_delay_quality"):
        getattr(R, name).argtypes = None
    inst = 9
    h = C.c_void_p(R.WebRtcAecm_Create())
    assert R.WebRtcAecm_Init(h, fs) == 0
    assert R.ref_shim_set_delay_estimator_options(h, robust, 0, -1, -1) == 0
    o = ol.OracleInstance(fs, 1, 3)
    assert o.set_delay_estimator_options(robust=robust) == 0
    out = np.zeros(spc, dtype=np.int16)
    seen = set()
    for c in range(far.shape[1] // spc):
        f = np.ascontiguousarray(far[inst, c * spc:(c + 1) * spc])
        x = np.ascontiguousarray(near[inst, c * spc:(c + 1) * spc])
        assert R.WebRtcAecm_BufferFarend(h, f.ctypes.data_as(C.c_void_p), C.c_size_t(spc)) == 0
        assert R.WebRtcAecm_Process(h, x.ctypes.data_as(C.c_void_p), None, out.ctypes.data_as(C.c_void_p), C.c_size_t(spc), C.c_int16(40)) == 0
        o.buffer_farend(f)
        rc, w = o.process(x, None, 40)
        np.testing.assert_array_equal(out, w)
        if c % 50 == 49:
            d, q = o.delay_estimate()
            assert d == R.ref_shim_last_delay(h)
            assert q == R.ref_shim_last_delay_quality(h), (c, q)
            seen.add(q)
    assert len(seen) > 3  # the quality moves
    R.WebRtcAecm_Free(h)
    o.close()
