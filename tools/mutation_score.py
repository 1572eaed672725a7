#!/usr/bin/env python
"""How sharp is the parity suite?  Mutation testing of the DEVICE code on the emulated warp.

Each mutant is the device source with ONE small change -- an integer literal off by one, a shift count off by one, < for <=,
+ for -, && for || -- in aecm_block.cuh, aecm_frame_api.cuh, aecm_prims.cuh, aecm_fft.cuh or aecm_init.cuh.  The mutant is
compiled for the host (tests/host_emu/call_emu.cpp, as the CPU suite does) and run on the case families of tests/cases.py
(one instance each), the delay-estimator configurations, scripts with control calls and option changes in mid-stream, against
results of the ORACLE computed once, plus the reference's FFT known answers, the scalar-primitive known answers and three final
records field by field; a mutant is KILLED when anything differs, when it crashes or when it hangs (lanes that disagree about
the number of warp-level operations).  Code the default build compiles out (measurement knobs, the butterfly variants not in
use) is not mutated.  Survivors are listed with their source line: each is either an equivalent mutant (the
change cannot alter a result) or a hole in the cases.   ->  profiles/mutation_score.md

    python tools/mutation_score.py [--mutants 240] [--seed 1] [--jobs 8]

TEST INFRASTRUCTURE; nothing here touches the product build.
"""
import argparse
import ctypes as C
import multiprocessing as mp
import os
import random
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
FILES = ["aecm_block.cuh", "aecm_frame_api.cuh", "aecm_prims.cuh", "aecm_fft.cuh", "aecm_init.cuh"]
# cases a worker runs, cheapest and most diverse first (a mutant usually dies on the first one)
CASE_ORDER = ["16k_echomode3", "8k_clean_path", "16k_stress_clean_cng_off_mode4", "8k_far_underrun_ms_mix", "16k_80_sample_calls", "c5_delay_sweep_0_120ms",
              "16k_full_scale_alternating", "8k_echomode0", "16k_ms_unstable_startup", "8k_far_overrun_ms_jumps", "16k_adv_loud_far_zero_near_mode4",
              "8k_adv_tone_far_whisper_near", "c2_16k_default_10s", "8k_constant_far_level", "16k_silence", "8k_stress", "16k_ms_jumps", "16k_cng_off",
              "8k_160_sample_calls", "16k_far_only", "16k_echomode0", "16k_echomode4", "16k_adv_gated_tone_dc_near", "8k_near_few_lsb_30s"]


def _branch_lines():
    """line ranges of the two butterfly branches the default build does not use"""
    src = open(os.path.join(ROOT, "webrtc_aecm_b200", "csrc", "aecm_fft.cuh")).read().split("\n")
    a = next(i for i, l in enumerate(src) if "if (!kInv && (AECM_FMA_BIAS & 1))" in l) + 1
    b = next(i for i, l in enumerate(src) if "} else if (kInv && (AECM_FMA_BIAS & 2)) {" in l) + 1
    c = next(i for i, l in enumerate(src) if l.strip() == "} else {" and i > b) + 1
    d = next(i for i, l in enumerate(src) if l.startswith("}") and i > c) + 1
    return (a + 1, b - 1), (c + 1, d - 1)


FMA_FWD, OLD_INV = _branch_lines()


def strip_comment(line):
    i = line.find("//")
    return line if i < 0 else line[:i]


def candidates(path):
    """[(line index, start, end, replacement, kind)] over the code part of every line inside function bodies"""
    out = []
    lines = open(path).read().split("\n")
    in_block_comment = False
    for k, line in enumerate(lines):
        code = strip_comment(line)
        if "/*" in code:
            in_block_comment = "*/" not in code
            continue
        if in_block_comment:
            in_block_comment = "*/" not in code
            continue
        s = code.strip()
        if not s or s.startswith(("#", "static_assert", "template", "struct", "constexpr", "namespace", "}")) or "asm" in s or "static_assert" in s or '"' in s:
            continue
        if not line.startswith("    "):  # top-level declarations, tables
            continue
        if "kSyncPoints" in code or "kRowResync" in code or '"swizzle deltas"' in code or "DxD::d(" in code and "==" in code:
            continue  # measurement-only build knobs (off in the product build), static_assert continuation lines
        if os.path.basename(path) == "aecm_fft.cuh" and (FMA_FWD[0] <= k + 1 <= FMA_FWD[1] or OLD_INV[0] <= k + 1 <= OLD_INV[1]):
            continue  # butterfly variants that the default build (AECM_FMA_BIAS=2) compiles out
        for m in re.finditer(r"(?<![\w.])(\d{1,5})(?![\w.])", code):
            v = int(m.group(1))
            if code[:m.start()].rstrip().endswith(("[", "<", "kFull")) and "]" in code[m.end():m.end() + 1]:
                continue  # array extents
            out.append((k, m.start(), m.end(), str(v + 1), "literal+1"))
        for m in re.finditer(r">> (\d{1,2})\b", code):
            out.append((k, m.start(1), m.end(1), str(int(m.group(1)) + 1), "shift+1"))
        for m in re.finditer(r" (<=|>=|<|>) ", code):
            op = m.group(1)
            if "template" in code or "<<" in code[m.start() - 1:m.end() + 1] or ">>" in code[m.start() - 1:m.end() + 1]:
                continue
            out.append((k, m.start(1), m.end(1), {"<": "<=", "<=": "<", ">": ">=", ">=": ">"}[op], "relational"))
        for m in re.finditer(r" (\+|-) ", code):
            out.append((k, m.start(1), m.end(1), "-" if m.group(1) == "+" else "+", "plus/minus"))
        for m in re.finditer(r" (&&|\|\|) ", code):
            out.append((k, m.start(1), m.end(1), "||" if m.group(1) == "&&" else "&&", "and/or"))
    return lines, out


_cache = {}


def expected():
    """oracle outputs of the cases, computed once per process tree (forked workers inherit them)"""
    if _cache:
        return _cache
    import oracle_lib as ol
    from cases import CASES, make_inputs, run_kwargs

    by = {c["name"]: c for c in CASES}
    for name in CASE_ORDER:
        case = by[name]
        far, near, clean, ms = make_inputs(case)
        want, ret, _ = ol.run_streams(far, near, impl="oracle", clean=clean, ms=ms, **run_kwargs(case))
        _cache[name] = (case, far, near, clean, ms, want, ret)
    return _cache


_extra = {}


def expected_extra():
    """the delay-estimator configurations (kExt code), scripts with control calls in mid-stream, and scripts with
    random option changes: the oracle's outputs, delay traces, float qualities and return values"""
    if _extra:
        return _extra
    import oracle_lib as ol
    import test_fuzz_api as fz
    from cases import CASES, DLY_CONFIGS, dly_inputs
    from webrtc_aecm_b200.synth import synth_streams

    by = {c["name"]: c for c in CASES}
    dly = []
    for cfg in DLY_CONFIGS[:8]:
        fs, spc, far, near, ms, kw = dly_inputs(cfg)
        want, ret, _, trace = ol.run_streams(far, near, impl="oracle", ms=ms, dly=cfg["dly"], soft_reset=cfg["soft_reset"], delay_trace=True, **kw)
        dly.append((cfg, dict(by["c5_delay_sweep_0_120ms"], fs=fs, spc=spc), far, near, ms, want, trace))
    _extra["dly"] = dly
    ctl = []
    for seed, fs in [(2, 16000), (6, 8000), (14, 16000)] + [(sd, 8000 if sd % 3 == 0 else 16000) for sd in range(21, 45)] + list(fz.MUTANT_KILLERS):
        killer = (seed, fs) in fz.MUTANT_KILLERS  # these run as the test suite runs them: 500 calls, no control events
        sc = fz.make_script(seed, 500 if killer else 400, fs)
        ev = (fz.make_control_events(seed, 400) + [(77, 2, 1, 0), (78, 2, 0, 4), (79, 2, 1, -1), (80, 2, 1, 5)]) if seed % 2 == 0 and not killer else []
        ev.sort(key=lambda e: e[0])
        rcs = []
        want, want_rc = fz.run_oracle(sc, ev, rcs)
        ctl.append((sc, ev, want, int(np.bitwise_or.reduce(want_rc)), rcs))
    for fs, base in ((16000, 20), (8000, 12), (16000, 29), (16000, 25)):  # the start-up stability check at exactly its tolerance
        sc = fz.tolerance_script(fs, base)
        want, want_rc = fz.run_oracle(sc)
        ctl.append((sc, [], want, int(np.bitwise_or.reduce(want_rc)), []))
    _extra["ctl"] = ctl
    opt = []
    for seed in (201, 203):  # two of the suite's own option-change scripts (tests/test_call_path_emulated_warp.py), same generator
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
        events += [(50, 1, 1, 0, 0, 0), (51, 1, -1, 0, 0, 0), (52, 1, 0, 0, 0, 0)]  # the smallest shifts
        events.sort(key=lambda e: e[0])
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
            _, w = o.process(near[c * spc:(c + 1) * spc], None, 40)
            want[c * spc:(c + 1) * spc] = w
            want_d[c], want_q[c] = o.delay_estimate()
        o.close()
        opt.append((fs, spc, calls, far, near, np.ascontiguousarray(np.array(events, dtype=np.int32)), want, want_rc, want_d, want_q))
    _extra["opt"] = opt
    # final records of three cases, field by field (state that cannot reach the output is state all the same)
    fin = []
    for name, inst, calls in (("16k_stress", 2, 300), ("8k_far_underrun_ms_mix", 1, 300), ("16k_ms_jumps", 0, 500)):
        case = by[name]
        far, near, clean, ms = __import__("cases").make_inputs(case)
        spc = case["spc"]
        far, near = np.ascontiguousarray(far[:, :calls * spc]), np.ascontiguousarray(near[:, :calls * spc])
        fin.append((dict(case, skip_far_every=0, extra_far_every=0), far, near, inst, calls))
    _extra["final"] = fin
    # quality of the (non-robust) estimate from the very first call of a plain stream
    far, near = synth_streams(1, 300 * 160, 16000, seed=31, delay_ms=55)
    o = ol.OracleInstance(16000, 1, 3)
    q = np.zeros(300, dtype=np.float32)
    dd = np.zeros(300, dtype=np.int32)
    for c in range(300):
        o.buffer_farend(far[0, c * 160:(c + 1) * 160])
        o.process(near[0, c * 160:(c + 1) * 160], None, 40)
        dd[c], q[c] = o.delay_estimate()
    o.close()
    _extra["quality0"] = (np.ascontiguousarray(far[0]), np.ascontiguousarray(near[0]), dd, q)
    vec = np.load(os.path.join(ROOT, "tests", "golden", "vectors.npz"))
    adv = np.load(os.path.join(ROOT, "tests", "golden", "fft_adversarial.npz"))
    _extra["fft"] = ({k: np.ascontiguousarray(vec[k]) for k in ("fft_fwd_in", "fft_fwd_out", "fft_inv_in", "fft_inv_out", "fft_inv_scale")},
                     {k: np.ascontiguousarray(adv[k]) for k in ("fwd_in", "fwd_out", "inv_in", "inv_out", "inv_scale")})
    return _extra


def primitives_ok(lib):
    """the known answers of tests/test_device_primitives_emulated.py, condensed"""
    import oracle_lib as ol

    O = ol.oracle()
    O.oracle_norm_w32.argtypes, O.oracle_norm_u32.argtypes, O.oracle_norm_w16.argtypes = [C.c_int32], [C.c_uint32], [C.c_int16]
    lib.emu_lcg_draws.restype = C.c_uint32
    rng = np.random.default_rng(5)
    vals = [0, 1, -1, 2, -2, 64, 65, 32767, 32768, -32768, 65535, 65536, (1 << 30), -(1 << 30), (1 << 31) - 1, -(1 << 31)]
    vals += [int(v) for v in rng.integers(-(1 << 31), 1 << 31, 400)] + [int(1 << k) for k in range(31)] + [int((1 << k) - 1) for k in range(1, 32)]
    for d in range(1, 66):
        for n in vals[:60] + [int(k) * d + e for k in rng.integers(0, (1 << 31) // d, 6) for e in (-1, 0, 1)]:
            if -(1 << 31) <= n < (1 << 31):
                want = abs(n) // d * (1 if n >= 0 else -1)
                if lib.emu_div_by_bin(C.c_int(n), d) != (n if d == 1 else want):
                    return False
    for v in vals:
        v16 = ((v + 32768) & 0xFFFF) - 32768
        if lib.emu_norm_w32(C.c_int(v)) != O.oracle_norm_w32(v) or lib.emu_norm_u32(C.c_uint(v & 0xFFFFFFFF)) != O.oracle_norm_u32(v & 0xFFFFFFFF) \
                or lib.emu_norm_w16(C.c_int(v16)) != O.oracle_norm_w16(v16):
            return False
        if v >= 0 and lib.emu_sqrt_floor(C.c_int(v)) != O.oracle_sqrt_floor(v):
            return False
    for a_, b_ in zip(vals[::2], vals[1::2]):
        if lib.emu_add_sat32(C.c_int(a_), C.c_int(b_)) != max(-(1 << 31), min((1 << 31) - 1, a_ + b_)):
            return False
    for e in [0, 1, 3, 255, 256, 65535, 1 << 20, (1 << 31) - 1, 1 << 31, (1 << 32) - 1] + [int(x) for x in rng.integers(0, 1 << 32, 200)]:
        for q in (0, 3, 15, 27):
            z = 32 - e.bit_length()
            want = 896 if e == 0 else ((896 + ((31 - z) << 8) + (((e << z) & 0x7FFFFFFF) >> 23) - (q << 8) + 32768) & 0xFFFF) - 32768
            if lib.emu_log_energy_q8(C.c_uint(e), q) != want:
                return False
    for sd in [666, 0, 1, 0x7FFFFFFF] + [int(x) for x in rng.integers(0, 1 << 31, 50)]:
        buf, got = np.zeros(64, dtype=np.int16), np.zeros(64, dtype=np.int16)
        s0 = C.c_uint32(sd)
        O.oracle_rand_u_array(buf.ctypes.data_as(C.c_void_p), 64, C.byref(s0))
        if lib.emu_lcg_draws(C.c_uint32(sd), got.ctypes.data_as(C.c_void_p)) != s0.value or not np.array_equal(buf, got):
            return False
    return True


def fft_ok(fft_lib, vec, adv):
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    fft_lib.emu_fft128_inverse.restype = C.c_int
    for fin, fout, iin, iout, isc in ((vec["fft_fwd_in"], vec["fft_fwd_out"], vec["fft_inv_in"], vec["fft_inv_out"], vec["fft_inv_scale"]),
                                      (adv["fwd_in"], adv["fwd_out"], adv["inv_in"], adv["inv_out"], adv["inv_scale"])):
        got = np.zeros_like(fout)
        for i in range(len(fin)):
            fft_lib.emu_fft128_forward(vp(fin[i]), vp(got[i]), 1)
        if not np.array_equal(got, fout):
            return False
        for i in range(0, len(fin) - 1, 2):
            pair, g2 = np.ascontiguousarray(fin[i:i + 2]), np.zeros((2, 130), dtype=np.int16)
            fft_lib.emu_fft128_forward(vp(pair), vp(g2), 2)
            if not np.array_equal(g2, fout[i:i + 2]):
                return False
        got = np.zeros_like(iout)
        for i in range(len(iin)):
            if fft_lib.emu_fft128_inverse(vp(iin[i]), vp(got[i])) != int(isc[i]):
                return False
        if not np.array_equal(got, iout):
            return False
    return True


def run_cases(lib_path, names):
    """child process body: exit code 0 = everything equal, 1 = a difference"""
    import oracle_lib as ol
    import test_call_path_emulated_warp as t
    from webrtc_aecm_b200 import api

    lib = C.CDLL(lib_path)
    x0 = expected_extra()
    if not fft_ok(C.CDLL(lib_path.replace("libcall_emu_mutant", "libfft_emu_mutant")), *x0["fft"]) or not primitives_ok(lib):
        os._exit(1)
    lib.emu_state_bytes.restype = C.c_size_t
    lib.emu_run_stream.restype = C.c_int
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    for name in names:
        case, far, near, clean, ms, want, ret = expected()[name]
        inst = far.shape[0] - 1
        out, acc, _ = t.emu_run(lib, case, far, near, clean, ms, inst)
        if not np.array_equal(out, want[inst]) or (acc != 0) != (ret[inst] != 0):
            os._exit(1)
    for name in ("8k_far_underrun_ms_mix", "16k_stress_clean_cng_off_mode4", "16k_80_sample_calls"):  # planes at odd int16 positions: the 16-bit movers
        case, far, near, clean, ms, want, ret = expected()[name]
        out, acc, _ = t.emu_run(lib, case, far, near, clean, ms, 0, misalign=True)
        if not np.array_equal(out, want[0]):
            os._exit(1)
    x = expected_extra()
    for cfg, case, far, near, ms, want, trace in x["dly"]:
        inst = cfg["instances"][0]
        got = np.zeros(far.shape[1] // case["spc"], dtype=np.int32)
        out, acc, _ = t.emu_run(lib, case, far, near, None, ms, inst, dly=cfg["dly"], soft_reset=cfg["soft_reset"], delay_trace=got)
        if not np.array_equal(out, want[inst]) or not np.array_equal(got, trace[inst]) or acc:
            os._exit(1)
    for sc, ev, want, want_acc, rcs in x["ctl"]:
        evs = np.ascontiguousarray(np.array([(c, k, a, b, 0, 0) for c, k, a, b in ev], dtype=np.int32).reshape(-1, 6))
        got_ev = np.full(len(ev), 99, dtype=np.int32)
        out = np.zeros_like(sc["near"])
        ms16 = np.ascontiguousarray(sc["ms"].astype(np.int16))
        spc_calls = np.ascontiguousarray(sc["n"].astype(np.int32))
        ops = np.ascontiguousarray(sc["far_ops"])
        cl = np.ascontiguousarray(sc["cln"]) if sc["clean"].all() else None
        far, near = np.ascontiguousarray(sc["far"]), np.ascontiguousarray(sc["near"])
        nbar = C.c_long(0)
        acc = lib.emu_run_stream(sc["fs"], sc["cng"], sc["echo_mode"], vp(far), vp(near), vp(cl), vp(out), len(sc["n"]), 0, vp(ms16), vp(ops), None,
                                 C.byref(nbar), None, None, vp(spc_calls), vp(evs) if ev else None, len(ev), vp(got_ev) if ev else None, None)
        if not np.array_equal(out, want) or acc != want_acc or got_ev.tolist() != rcs:
            os._exit(1)
    for case, far, near, inst, calls in x["final"]:
        out, acc, state = t.emu_run(lib, case, far, near, None, None, inst, want_state=True)
        o = ol.OracleInstance(case["fs"], case["cng"], case["echo_mode"])
        spc = case["spc"]
        for c in range(calls):
            o.buffer_farend(far[inst, c * spc:(c + 1) * spc])
            o.process(near[inst, c * spc:(c + 1) * spc], None, 40)
        bad = ol.compare_states(api.InstanceState.from_buffer_copy(state.tobytes()), o.state())
        o.close()
        if bad:
            os._exit(1)
    far, near, want_d, want_q = x["quality0"]
    out = np.zeros_like(near)
    got_d, got_q = np.zeros(300, dtype=np.int32), np.zeros(300, dtype=np.float32)
    nbar = C.c_long(0)
    lib.emu_run_stream(16000, 1, 3, vp(far), vp(near), None, vp(out), 300, 160, None, None, None, C.byref(nbar), None, vp(got_d), None, None, 0, None, vp(got_q))
    if not (np.array_equal(got_d, want_d) and np.array_equal(got_q, want_q)):
        os._exit(1)
    for fs, spc, calls, far, near, ev, want, want_rc, want_d, want_q in x["opt"]:
        out = np.zeros_like(near)
        got_rc = np.full(len(ev), 99, dtype=np.int32)
        got_d, got_q = np.zeros(calls, dtype=np.int32), np.zeros(calls, dtype=np.float32)
        dly = np.array([1, 0, 0, -1, -1, 0, 0], dtype=np.int32)
        nbar = C.c_long(0)
        acc = lib.emu_run_stream(fs, 1, 3, vp(far), vp(near), None, vp(out), calls, spc, None, None, None, C.byref(nbar), vp(dly), vp(got_d), None, vp(ev),
                                 len(ev), vp(got_rc), vp(got_q))
        if acc or not (np.array_equal(out, want) and np.array_equal(got_rc, want_rc) and np.array_equal(got_d, want_d) and np.array_equal(got_q, want_q)):
            os._exit(1)
    os._exit(0)


def one(job):
    idx, fname, k, a, b, rep, kind = job
    t0 = time.time()
    with tempfile.TemporaryDirectory() as d:
        csrc = os.path.join(d, "webrtc_aecm_b200", "csrc")
        shutil.copytree(os.path.join(ROOT, "webrtc_aecm_b200", "csrc"), csrc, ignore=shutil.ignore_patterns("build", "*.o"))
        shutil.copytree(os.path.join(ROOT, "tests", "host_emu"), os.path.join(d, "tests", "host_emu"), ignore=shutil.ignore_patterns("*.so"))
        shutil.copytree(os.path.join(ROOT, "include"), os.path.join(d, "include"))
        path = os.path.join(csrc, fname)
        lines = open(path).read().split("\n")
        before = lines[k]
        lines[k] = before[:a] + rep + before[b:]
        open(path, "w").write("\n".join(lines))
        lib = os.path.join(d, "libcall_emu_mutant.so")
        r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(d, "include"), "-o", lib,
                            os.path.join(d, "tests", "host_emu", "call_emu.cpp")], capture_output=True, text=True)
        if r.returncode == 0:
            r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", os.path.join(d, "libfft_emu_mutant.so"),
                                os.path.join(d, "tests", "host_emu", "fft_emu.cpp"), "-lpthread"], capture_output=True, text=True)
        desc = f"{fname}:{k + 1} [{kind}] `{before.strip()[:90]}` -> `...{lines[k][max(0, a - 12):a + len(rep) + 12].strip()}...`"
        if r.returncode != 0:
            return idx, "does not compile", desc, time.time() - t0
        pid = os.fork()
        if pid == 0:
            try:
                run_cases(lib, CASE_ORDER)
            finally:
                os._exit(3)
        deadline = time.time() + 240
        while True:
            done, status = os.waitpid(pid, os.WNOHANG)
            if done:
                break
            if time.time() > deadline:
                os.kill(pid, 9)
                os.waitpid(pid, 0)
                return idx, "killed (hang)", desc, time.time() - t0
            time.sleep(0.05)
        if os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0:
            return idx, "SURVIVED", desc, time.time() - t0
        return idx, "killed" if os.WIFEXITED(status) else "killed (crash)", desc, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mutants", type=int, default=240)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    rnd = random.Random(a.seed)
    sites = []
    for f in FILES:
        _, c = candidates(os.path.join(ROOT, "webrtc_aecm_b200", "csrc", f))
        sites += [(f, *x) for x in c]
    rnd.shuffle(sites)
    jobs = [(i, *s) for i, s in enumerate(sites[:a.mutants])]
    expected()  # before forking the pool
    expected_extra()
    t0 = time.time()
    with mp.Pool(a.jobs) as pool:
        res = sorted(pool.imap_unordered(one, jobs))
    n = {}
    for _, verdict, _, _ in res:
        n[verdict] = n.get(verdict, 0) + 1
    valid = len(res) - n.get("does not compile", 0)
    killed = valid - n.get("SURVIVED", 0)
    out = ["# Mutation score of the parity cases on the device code (tools/mutation_score.py)", "",
           f"{len(sites)} mutation sites in {', '.join(FILES)}; {len(res)} drawn at random (seed {a.seed}); {n.get('does not compile', 0)} do not compile.",
           f"Of the {valid} valid mutants **{killed} are killed** ({100.0 * killed / max(valid, 1):.1f} %) by {len(CASE_ORDER)} case families, 8 delay-estimator configurations (delay trace "
           f"per call), 32 random API scripts (half of them with control calls in mid-stream), 2 scripts with random option changes (float quality per call), the final records of 3 streams field by field, "
           f"the FFT known answers of the reference and the scalar-primitive known answers, all on the emulated warp "
           f"(one instance each; {n.get('killed (hang)', 0)} by hanging, {n.get('killed (crash)', 0)} by crashing), {n.get('SURVIVED', 0)} survive:", ""]
    out += [f"* {d}" for _, v, d, _ in res if v == "SURVIVED"]
    out += ["", f"({time.time() - t0:.0f} s on {a.jobs} processes)", ""]
    open(os.path.join(ROOT, "profiles", "mutation_score.md"), "w").write("\n".join(out))
    print("\n".join(out))


if __name__ == "__main__":
    main()
