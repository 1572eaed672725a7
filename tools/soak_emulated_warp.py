#!/usr/bin/env python
"""Soak: randomised API scripts (tests/test_fuzz_api.make_script, longer and with fresh seeds) through the DEVICE code on the
emulated warp (tests/host_emu) against the UNMODIFIED reference (oracle/_ref/libaecm_ref.so), call by call, on all host cores
for a given number of minutes; every other script also carries control calls in mid-stream (set_config valid and refused, Init
at the same or the other rate, InitEchoPath: test_fuzz_api.make_control_events).
Test infrastructure: needs /root/reference to have been compiled into oracle/_ref.

    python tools/soak_emulated_warp.py --minutes 20 [--start-seed 100000] [--calls 1500] [--jobs 8]

Prints one line per divergence (seed, fs, first differing sample) and a summary; exit 1 if anything differed."""
import argparse
import ctypes as C
import multiprocessing as mp
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def emu_lib(build=False):
    import test_call_path_emulated_warp as t

    if build and (not os.path.exists(t.LIB) or any(os.path.getmtime(d) > os.path.getmtime(t.LIB) for d in t.DEPS)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                               "-o", t.LIB, t.SRC])
    lib = C.CDLL(t.LIB)
    lib.emu_run_stream.restype = C.c_int
    return lib


def one(args):
    seed, calls = args
    import test_fuzz_api as fz

    fs = 8000 if seed % 3 == 0 else 16000
    s = fz.make_script(seed, calls, fs)
    events = fz.make_control_events(seed, calls) if seed % 2 else []  # every other script carries mid-stream control calls
    want_ev = []
    want, rcs = fz.run_reference(s, events, want_ev)
    ev = np.ascontiguousarray(np.array([(c, k, a, b, 0, 0) for c, k, a, b in events], dtype=np.int32).reshape(-1, 6))
    got_ev = np.full(len(events), 99, dtype=np.int32)
    emu = emu_lib()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    out = np.zeros_like(s["near"])
    ms16 = np.ascontiguousarray(s["ms"].astype(np.int16))
    spc_calls = np.ascontiguousarray(s["n"].astype(np.int32))
    ops = np.ascontiguousarray(s["far_ops"])
    cl = np.ascontiguousarray(s["cln"]) if s["clean"].all() else None
    far, near = np.ascontiguousarray(s["far"]), np.ascontiguousarray(s["near"])
    nbar = C.c_long(0)
    acc = emu.emu_run_stream(fs, s["cng"], s["echo_mode"], vp(far), vp(near), None if cl is None else vp(cl), vp(out), calls, 0, vp(ms16),
                             vp(ops), None, C.byref(nbar), None, None, vp(spc_calls), vp(ev) if events else None, len(events),
                             vp(got_ev) if events else None, None)
    bad = np.nonzero(out != want)[0]
    rc_ok = acc == int(np.bitwise_or.reduce(rcs)) and got_ev.tolist() == want_ev
    return seed, fs, (int(bad[0]) if bad.size else -1), rc_ok, int(want.size)


def one_dly(args):
    """SURVEY 8-f4 under fire, against the reference itself: the estimator's setters, soft resets and re-creations at random calls
    (the event mix of tests/test_call_path_emulated_warp.py) on the kExt device code and on oracle/_ref/libaecm_ref_shim.so."""
    seed, calls = args
    import oracle_lib as ol
    from webrtc_aecm_b200.synth import synth_streams

    rng = np.random.default_rng(seed)
    fs = 8000 if seed % 3 == 0 else 16000
    spc = 80 if fs == 8000 else 160
    far, near = synth_streams(1, calls * spc, fs, seed=seed, delay_ms=int(rng.integers(0, 115)), kind="stress" if rng.random() < 0.2 else "echo")
    far, near = np.ascontiguousarray(far[0]), np.ascontiguousarray(near[0])
    max_la = int(rng.integers(0, 16))
    events = [(0, 0, int(rng.integers(0, 2)), int(rng.integers(0, 4)), max_la, -1)]
    for c in range(1, calls):
        u = rng.random()
        if u < 0.01:
            events.append((c, 0, int(rng.integers(0, 2)), int(rng.integers(0, 5)), -1, int(rng.integers(-1, max_la + 3))))
        elif u < 0.018:
            events.append((c, 1, int(rng.integers(-40, 41)), 0, 0, 0))
        elif u < 0.021:
            max_la = int(rng.integers(0, 16))
            events.append((c, 0, int(rng.integers(0, 2)), 0, max_la, -1))
    R = C.CDLL(ol.ref_shim_path())
    R.WebRtcAecm_Create.restype = C.c_void_p
    R.ref_shim_last_delay_quality.restype = C.c_float
    h = C.c_void_p(R.WebRtcAecm_Create())
    assert R.WebRtcAecm_Init(h, fs) == 0
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    want = np.zeros_like(near)
    want_rc = np.zeros(len(events), dtype=np.int32)
    want_d, want_q = np.zeros(calls, dtype=np.int32), np.zeros(calls, dtype=np.float32)
    o = np.zeros(spc, dtype=np.int16)
    k = 0
    for c in range(calls):
        while k < len(events) and events[k][0] == c:
            e = events[k]
            want_rc[k] = (R.ref_shim_set_delay_estimator_options(h, e[2], e[3], e[4], e[5]) if e[1] == 0
                          else R.ref_shim_soft_reset_delay_estimator(h, e[2]))
            k += 1
        f = np.ascontiguousarray(far[c * spc:(c + 1) * spc])
        x = np.ascontiguousarray(near[c * spc:(c + 1) * spc])
        R.WebRtcAecm_BufferFarend(h, vp(f), C.c_size_t(spc))
        R.WebRtcAecm_Process(h, vp(x), None, vp(o), C.c_size_t(spc), C.c_int16(40))
        want[c * spc:(c + 1) * spc] = o
        want_d[c], want_q[c] = R.ref_shim_last_delay(h), R.ref_shim_last_delay_quality(h)
    R.WebRtcAecm_Free(h)
    emu = emu_lib()
    ev = np.ascontiguousarray(np.array(events, dtype=np.int32))
    out = np.zeros_like(near)
    got_rc = np.full(len(events), 99, dtype=np.int32)
    got_d, got_q = np.zeros(calls, dtype=np.int32), np.zeros(calls, dtype=np.float32)
    dly = np.array([1, 0, 0, -1, -1, 0, 0], dtype=np.int32)
    nbar = C.c_long(0)
    acc = emu.emu_run_stream(fs, 1, 3, vp(far), vp(near), None, vp(out), calls, spc, None, None, None, C.byref(nbar), vp(dly), vp(got_d), None,
                             vp(ev), len(events), vp(got_rc), vp(got_q))
    bad = np.nonzero(out != want)[0]
    ok = acc == 0 and (got_rc == want_rc).all() and (got_d == want_d).all() and (got_q == want_q).all()
    return seed, fs, (int(bad[0]) if bad.size else -1), bool(ok), int(want.size)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--start-seed", type=int, default=100000)
    ap.add_argument("--calls", type=int, default=1500)
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--dly", action="store_true", help="the delay-estimator options (SURVEY 8-f4) against oracle/_ref/libaecm_ref_shim.so")
    a = ap.parse_args()
    import oracle_lib as ol

    if ol.ref_path() is None or (a.dly and ol.ref_shim_path() is None):
        sys.exit("oracle/_ref/libaecm_ref.so is not built (python -c 'import __graft_entry__ as g; g.build()' with /root/reference present)")
    emu_lib(build=True)  # the parent builds; the workers only load
    t0 = time.time()
    done = samples = failed = 0
    seed = a.start_seed
    with mp.Pool(a.jobs) as pool:
        while time.time() - t0 < a.minutes * 60:
            batch = [(seed + k, a.calls) for k in range(a.jobs * 4)]
            seed += len(batch)
            for sd, fs, bad, rc_ok, n in pool.imap_unordered(one_dly if a.dly else one, batch):
                done += 1
                samples += n
                if bad >= 0 or not rc_ok:
                    failed += 1
                    print(f"DIVERGENCE seed {sd} fs {fs}: first differing sample {bad}, return codes / delay estimates {'equal' if rc_ok else 'DIFFER'}", flush=True)
    print(f"{done} scripts of {a.calls} calls (seeds {a.start_seed}..{seed - 1}), {samples} output samples compared with the reference, "
          f"{failed} divergences, {time.time() - t0:.0f} s on {a.jobs} processes")
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
