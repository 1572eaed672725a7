"""Sanitizer builds on the CPU.  (1) The checker itself must be free of undefined behaviour: the oracle built with -fsanitize=undefined (aborting on the first
report) runs the case matrix, the delay-estimator configurations and a batch of randomised API scripts.  (The reference has one
known UB spot on these inputs -- the wide NLMS shift, aecm_core.cc:911 -- which the oracle restates with the masked shift
count x86 executes)  (2) The DEVICE code, compiled for the host against tests/host_emu/warp_emu.h with
the same sanitizer (array bounds of every record / scratch array, shift counts, signed overflow, alignment), runs the case matrix
and the delay-estimator configurations on the emulated warp."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import oracle_lib as ol
ol.ORACLE_SO = %(so)r
import test_fuzz_api as fz
from cases import CASES, DLY_CONFIGS, dly_inputs, make_inputs, run_kwargs
n = 0
for seed, fs in fz.SCRIPTS[:12]:
    fz.run_oracle(fz.make_script(seed, 300, fs)); n += 1
for c in CASES:
    if c["seconds"] > 12:
        continue
    far, near, clean, ms = make_inputs(c)
    ol.run_streams(far, near, impl="oracle", clean=clean, ms=ms, **run_kwargs(c)); n += 1
for cfg in DLY_CONFIGS[:8]:
    fs, spc, far, near, ms, kw = dly_inputs(cfg)
    ol.run_streams(far, near, impl="oracle", ms=ms, dly=cfg["dly"], soft_reset=cfg["soft_reset"], **kw); n += 1
print("ok", n)
"""


def test_oracle_is_ubsan_clean(tmp_path):
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    so = str(tmp_path / "libaecm_oracle_ubsan.so")
    r = subprocess.run(["gcc", "-O1", "-g", "-fwrapv", "-ffp-contract=off", "-fsanitize=undefined", "-fno-sanitize-recover=undefined",
                        "-fPIC", "-shared", "-std=gnu99", "-o", so, os.path.join(ROOT, "oracle", "aecm_oracle.c"), "-lpthread", "-ldl", "-lm"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no UBSan runtime for this compiler: " + r.stderr[:200])
    run = subprocess.run([sys.executable, "-c", CODE % dict(root=ROOT, tests=os.path.join(ROOT, "tests"), so=so)],
                         capture_output=True, text=True, timeout=900)
    assert run.returncode == 0 and "runtime error" not in run.stderr, run.stderr[-2000:]
    assert run.stdout.strip().startswith("ok")


EMU_CODE = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import ctypes as C
import numpy as np
import oracle_lib as ol
import test_call_path_emulated_warp as t
from cases import CASES, CASE_BY_NAME, DLY_CONFIGS, dly_inputs, make_inputs, run_kwargs
lib = C.CDLL(%(so)r); lib.emu_state_bytes.restype = C.c_size_t; lib.emu_run_stream.restype = C.c_int
n = 0
for c in CASES:
    if c["seconds"] > 12:
        continue
    far, near, clean, ms = make_inputs(c)
    want, ret, _ = ol.run_streams(far, near, impl="oracle", clean=clean, ms=ms, **run_kwargs(c))
    inst = far.shape[0] - 1
    out, acc, _ = t.emu_run(lib, c, far, near, clean, ms, inst)
    assert np.array_equal(out, want[inst]), c["name"]
    n += 1
for cfg in DLY_CONFIGS[:8]:
    fs, spc, far, near, ms, kw = dly_inputs(cfg)
    c = dict(CASE_BY_NAME["c5_delay_sweep_0_120ms"], fs=fs, spc=spc)
    t.emu_run(lib, c, far, near, None, ms, cfg["instances"][0], dly=cfg["dly"], soft_reset=cfg["soft_reset"])
    n += 1
print("ok", n)
"""


def test_device_code_is_ubsan_clean_on_the_emulated_warp(tmp_path):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    so = str(tmp_path / "libcall_emu_ubsan.so")
    # C++20: left shifts of negative values are defined (two's complement), as they are on the device
    r = subprocess.run(["g++", "-O1", "-g", "-ffp-contract=off", "-std=c++20", "-fsanitize=undefined", "-fno-sanitize-recover=undefined",
                        "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", so,
                        os.path.join(ROOT, "tests", "host_emu", "call_emu.cpp")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no UBSan runtime for this compiler: " + r.stderr[:200])
    run = subprocess.run([sys.executable, "-c", EMU_CODE % dict(root=ROOT, tests=os.path.join(ROOT, "tests"), so=so)],
                         capture_output=True, text=True, timeout=1200)
    assert run.returncode == 0 and "runtime error" not in run.stderr, run.stderr[-2000:]
    assert run.stdout.strip().startswith("ok")
