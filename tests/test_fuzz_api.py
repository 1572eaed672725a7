"""Randomised API call sequences: what the fixed case matrix cannot enumerate.

A script is a list of 10 ms calls; every call draws its frame size (80 or 160 samples -- the reference accepts either at
either rate, echo_control_mobile.cc:209, 268), its msInSndCardBuf (mostly steady, sometimes jumping, sometimes out of
range) and how often BufferFarend runs before Process (0 = underrun, 1, 2 = overrun); a script passes a clean plane on
every call or on none.
The oracle and the unmodified reference are stepped call by call through their own entry points; output samples and
return codes must agree call for call.  The same scripts drive the device code on the emulated warp
(test_call_path_emulated_warp.py) and the GPU (test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from webrtc_aecm_b200.synth import synth_streams


def make_script(seed, n_calls, fs):
    """-> dict(n[n_calls], ms[n_calls], far_ops[n_calls], clean[n_calls] (bool), far, near, cln: int16 streams)"""
    rng = np.random.default_rng(seed)
    mixed = rng.random() < 0.7
    base_n = 160 if fs == 16000 else 80
    n = np.where(rng.random(n_calls) < (0.35 if mixed else 0.0), 240 - base_n, base_n).astype(np.int32)
    ms = np.full(n_calls, int(rng.integers(0, 200)), dtype=np.int32)
    k = 0
    while k < n_calls:  # piecewise constant with jumps, a few out-of-range values
        k += int(rng.integers(20, 120))
        if k < n_calls:
            ms[k:] = int(rng.choice([0, 10, 40, 80, 120, 250, 500, -5, 600, int(rng.integers(0, 500))]))
    ms = ms + (rng.random(n_calls) < 0.1) * rng.integers(-3, 4, n_calls)
    far_ops = rng.choice([0, 1, 2], size=n_calls, p=[0.06, 0.88, 0.06]).astype(np.uint8)
    if rng.random() < 0.3:  # a long underrun
        a = int(rng.integers(0, n_calls - 30))
        far_ops[a:a + int(rng.integers(5, 30))] = 0
    # a clean plane on every call or on none: switching it on and off in mid-stream leaves the reference's clean-frame FIFO
    # short, and WebRtcAecm_ProcessFrame then hands ProcessBlock stale ring-buffer bytes or an uninitialised stack array
    # (aecm_core.cc:537-546 with ring_buffer.c:113-131) -- there is no defined behaviour to be bit-exact with
    clean = np.full(n_calls, rng.random() < 0.35)
    total = int(n.sum())
    far, near = synth_streams(1, total, fs, seed=seed, delay_ms=int(rng.integers(0, 110)), kind="stress" if rng.random() < 0.3 else "echo")
    cln = synth_streams(1, total, fs, seed=seed + 1, delay_ms=5)[1]
    return dict(n=n, ms=ms.astype(np.int32), far_ops=far_ops, clean=clean, far=far[0], near=near[0], cln=cln[0], fs=fs,
                cng=int(rng.integers(0, 2)), echo_mode=int(rng.integers(0, 5)))


def event_path(seed):
    """The channel estimate an InitEchoPath event imports (tests/host_emu/call_emu.cpp emu_event_path is the same rule)."""
    return ((seed * 1103515245 + 12345 + np.arange(65, dtype=np.int64) * 7919) & 0x3FFF).astype(np.int16)


def make_control_events(seed, n_calls):
    """Control calls between the 10 ms calls of a script -> sorted list of (call, kind, a, b), applied BEFORE that call:
    kind 2 = WebRtcAecm_set_config(cngMode a, echoMode b) -- sometimes with values the reference refuses (:421-428),
    kind 3 = WebRtcAecm_Init(fs a) in mid-stream (same or the other rate; the default config comes back, :183-188),
    kind 4 = WebRtcAecm_InitEchoPath(event_path(a)) (:481-503).  (Kinds 0 and 1 are the delay-estimator events of the emulation.)"""
    rng = np.random.default_rng(seed + 7_000_000)
    ev = []
    for c in range(1, n_calls):
        u = rng.random()
        if u < 0.012:
            ev.append((c, 2, int(rng.choice([0, 1, 0, 1, 2, -1])), int(rng.choice([0, 1, 2, 3, 4, 3, 5, -1]))))
        elif u < 0.016:
            ev.append((c, 3, int(rng.choice([8000, 16000])), 0))
        elif u < 0.024:
            ev.append((c, 4, int(rng.integers(0, 1 << 20)), 0))
    return ev


def _events_at(events, c):
    return [e for e in events if e[0] == c]


def run_oracle(s, events=(), event_rc=None):
    o = ol.OracleInstance(s["fs"], s["cng"], s["echo_mode"])
    L = o.lib
    out = np.zeros_like(s["near"])
    rcs = []
    pos = 0
    for c in range(len(s["n"])):
        n = int(s["n"][c])
        sl = slice(pos, pos + n)
        r = 0
        for _, kind, a, b in _events_at(events, c):
            if kind == 2:
                er = L.oracle_aecm_set_config(o.h, a, b)
            elif kind == 3:
                er = L.oracle_aecm_init(o.h, a)
            else:
                pth = event_path(a)
                er = L.oracle_aecm_init_echo_path(o.h, pth.ctypes.data_as(C.c_void_p), 130)
            if event_rc is not None:
                event_rc.append(er)
        for _ in range(int(s["far_ops"][c])):
            r |= o.buffer_farend(s["far"][sl])
        rc, w = o.process(s["near"][sl], s["cln"][sl] if s["clean"][c] else None, int(s["ms"][c]))
        out[sl] = w
        rcs.append(r | rc)
        pos += n
    o.close()
    return out, np.array(rcs)


def run_reference(s, events=(), event_rc=None):
    R = C.CDLL(ol.ref_path())
    R.WebRtcAecm_Create.restype = C.c_void_p

    class Cfg(C.Structure):
        _fields_ = [("cngMode", C.c_int16), ("echoMode", C.c_int16)]

    h = C.c_void_p(R.WebRtcAecm_Create())
    assert R.WebRtcAecm_Init(h, s["fs"]) == 0
    assert R.WebRtcAecm_set_config(h, Cfg(s["cng"], s["echo_mode"])) == 0
    out = np.zeros_like(s["near"])
    rcs = []
    pos = 0
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for c in range(len(s["n"])):
        n = int(s["n"][c])
        f = np.ascontiguousarray(s["far"][pos:pos + n])
        x = np.ascontiguousarray(s["near"][pos:pos + n])
        cl = np.ascontiguousarray(s["cln"][pos:pos + n]) if s["clean"][c] else None
        o = np.zeros(n, dtype=np.int16)
        r = 0
        for _, kind, a, b in _events_at(events, c):
            if kind == 2:
                er = R.WebRtcAecm_set_config(h, Cfg(a, b))
            elif kind == 3:
                er = R.WebRtcAecm_Init(h, a)
            else:
                pth = event_path(a)
                er = R.WebRtcAecm_InitEchoPath(h, vp(pth), C.c_size_t(130))
            if event_rc is not None:
                event_rc.append(er)
        for _ in range(int(s["far_ops"][c])):
            r |= R.WebRtcAecm_BufferFarend(h, vp(f), C.c_size_t(n))
        r |= R.WebRtcAecm_Process(h, vp(x), None if cl is None else vp(cl), vp(o), C.c_size_t(n), C.c_int16(int(s["ms"][c])))
        out[pos:pos + n] = o
        rcs.append(r)
        pos += n
    R.WebRtcAecm_Free(h)
    return out, np.array(rcs)


SCRIPTS = [(seed, 8000 if seed % 3 == 0 else 16000) for seed in range(1, 41)]
# scripts found by tools/mutation_score.py's follow-up search: each kills a one-token mutant of the device code that the first 40 let live
# (the delay estimator's threshold update with a poor best match, aecm_block.cuh:138-139; the first-VAD correction of the echo log energy, :399)
MUTANT_KILLERS = [(seed, 8000 if seed % 3 == 0 else 16000) for seed in (46, 50, 87, 191, 269)]
SCRIPTS += MUTANT_KILLERS


@pytest.mark.skipif(ol.ref_path() is None, reason="oracle/_ref/libaecm_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed,fs", SCRIPTS)
def test_random_call_sequences_oracle_equals_reference(seed, fs):
    s = make_script(seed, 500, fs)
    a, ra = run_oracle(s)
    b, rb = run_reference(s)
    np.testing.assert_array_equal(ra, rb)
    bad = np.nonzero(a != b)[0]
    assert bad.size == 0, f"first differing sample {bad[0]}"


@pytest.mark.skipif(ol.ref_path() is None, reason="oracle/_ref/libaecm_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed,fs", [(s_, f_) for s_, f_ in SCRIPTS if s_ % 2 == 0])
def test_random_control_calls_in_mid_stream_oracle_equals_reference(seed, fs):
    """set_config (valid and refused), Init (same or other rate) and InitEchoPath at random points of the random scripts:
    every control call's return value, every 10 ms call's return code and every output sample equal the reference's."""
    s = make_script(seed, 500, fs)
    ev = make_control_events(seed, 500)
    ea, eb = [], []
    a, ra = run_oracle(s, ev, ea)
    b, rb = run_reference(s, ev, eb)
    assert ea == eb and len(ea) == len(ev)
    np.testing.assert_array_equal(ra, rb)
    bad = np.nonzero(a != b)[0]
    assert bad.size == 0, f"first differing sample {bad[0]}"


def test_control_events_cover_their_kinds():
    kinds, refused = set(), 0
    for seed, _ in SCRIPTS:
        for _, kind, a, b in make_control_events(seed, 500):
            kinds.add(kind)
            refused += kind == 2 and (a not in (0, 1) or not 0 <= b <= 4)
    assert kinds == {2, 3, 4} and refused >= 10


def tolerance_script(fs, base_ms):
    """msInSndCardBuf alternating between base_ms and base_ms + 8 during start-up: the deviation from the first value is EXACTLY the
    tolerance of the buffer-size stability check for small values (max(0.2 * ms, 8), echo_control_mobile.cc:305-337; `dev < tol`
    is a double comparison there), so every second call restarts the count until the script settles on base_ms."""
    s = make_script(7, 300, fs)
    ms = np.full(300, base_ms, dtype=np.int32)
    ms[1:80:2] = base_ms + 8
    return dict(s, ms=ms, far_ops=np.ones(300, dtype=np.uint8))


@pytest.mark.skipif(ol.ref_path() is None, reason="oracle/_ref/libaecm_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("fs,base_ms", [(16000, 20), (8000, 12), (16000, 29), (16000, 25)])
def test_startup_stability_check_at_exactly_the_tolerance(fs, base_ms):
    s = tolerance_script(fs, base_ms)
    a, ra = run_oracle(s)
    b, rb = run_reference(s)
    np.testing.assert_array_equal(ra, rb)
    np.testing.assert_array_equal(a, b)
    steady = dict(s, ms=np.full(300, base_ms, dtype=np.int32))
    assert not np.array_equal(run_oracle(steady)[0], a)  # the restarts really delay the end of start-up


def test_scripts_exercise_what_they_claim():
    """Across the scripts: both frame sizes at both rates, under- and overruns, clamped ms, clean planes, all echo modes."""
    seen = {"mixed": 0, "under": 0, "over": 0, "clamp": 0, "clean": 0}
    modes = set()
    for seed, fs in SCRIPTS:
        s = make_script(seed, 500, fs)
        seen["mixed"] += len(set(s["n"].tolist())) == 2
        seen["under"] += (s["far_ops"] == 0).any()
        seen["over"] += (s["far_ops"] == 2).any()
        seen["clamp"] += ((s["ms"] < 0) | (s["ms"] > 500)).any()
        seen["clean"] += s["clean"].any()
        modes.add(s["echo_mode"])
    assert all(v >= 5 for v in seen.values()), seen
    assert modes == {0, 1, 2, 3, 4}


@pytest.mark.skipif(ol.ref_shim_path() is None, reason="oracle/_ref/libaecm_ref_shim.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(100, 112))
def test_random_delay_estimator_option_changes_oracle_equals_reference(seed):
    """SURVEY 8-f4 under fire: at random calls of a stream the estimator's knobs are turned -- robust validation on/off, allowed
    offset, lookahead (within a max_lookahead chosen at the start), soft resets by random shifts, even a re-creation with a new
    max_lookahead -- on the reference (through oracle/ref_shim.cc) and on the oracle alike.  Output, the delay estimate and its
    float quality must agree after every call."""
    rng = np.random.default_rng(seed)
    fs, spc, calls = 16000, 160, 700
    far, near = synth_streams(1, calls * spc, fs, seed=seed, delay_ms=int(rng.integers(0, 115)))
    far, near = far[0], near[0]
    R = C.CDLL(ol.ref_shim_path())
    R.WebRtcAecm_Create.restype = C.c_void_p
    R.ref_shim_last_delay_quality.restype = C.c_float
    h = C.c_void_p(R.WebRtcAecm_Create())
    assert R.WebRtcAecm_Init(h, fs) == 0
    o = ol.OracleInstance(fs, 1, 3)
    soft = ol.oracle().oracle_aecm_soft_reset_delay_estimator
    soft.argtypes = [C.c_void_p, C.c_int]
    max_la = int(rng.integers(0, 9))
    first = dict(robust=int(rng.integers(0, 2)), allowed_offset=int(rng.integers(0, 4)), max_lookahead=max_la, lookahead=-1)
    assert R.ref_shim_set_delay_estimator_options(h, first["robust"], first["allowed_offset"], max_la, -1) == 0
    assert o.set_delay_estimator_options(**first) == 0
    out = np.zeros(spc, dtype=np.int16)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    events = 0
    for c in range(calls):
        if c in (50, 51, 52):  # the smallest shifts, which the random draws rarely hit: one block forward, one back, none
            shift = (1, -1, 0)[c - 50]
            assert R.ref_shim_soft_reset_delay_estimator(h, shift) == soft(o.h, shift)
        u = rng.random()
        if u < 0.02:  # turn the setters (a lookahead beyond max_lookahead must be refused by both, identically)
            args = (int(rng.integers(0, 2)), int(rng.integers(0, 5)), -1, int(rng.integers(-1, max_la + 3)))
            assert R.ref_shim_set_delay_estimator_options(h, *args) == o.set_delay_estimator_options(*args)
            events += 1
        elif u < 0.035:
            shift = int(rng.integers(-30, 31))
            assert R.ref_shim_soft_reset_delay_estimator(h, shift) == soft(o.h, shift)
            events += 1
        elif u < 0.04:  # a fresh estimator with another max_lookahead
            max_la = int(rng.integers(0, 12))
            args = (int(rng.integers(0, 2)), 0, max_la, -1)
            assert R.ref_shim_set_delay_estimator_options(h, *args) == o.set_delay_estimator_options(*args) == 0
            events += 1
        f = np.ascontiguousarray(far[c * spc:(c + 1) * spc])
        x = np.ascontiguousarray(near[c * spc:(c + 1) * spc])
        assert R.WebRtcAecm_BufferFarend(h, vp(f), C.c_size_t(spc)) == 0
        assert R.WebRtcAecm_Process(h, vp(x), None, vp(out), C.c_size_t(spc), C.c_int16(40)) == 0
        o.buffer_farend(f)
        rc, w = o.process(x, None, 40)
        assert rc == 0
        np.testing.assert_array_equal(out, w, err_msg=f"call {c}")
        d, q = o.delay_estimate()
        assert d == R.ref_shim_last_delay(h) and q == R.ref_shim_last_delay_quality(h), (c, d, q)
    assert events >= 10
    R.WebRtcAecm_Free(h)
    o.close()
