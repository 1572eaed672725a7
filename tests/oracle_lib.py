"""ctypes access to the CPU checkers under oracle/ (test infrastructure only).

* ``oracle()``  -> oracle/libaecm_oracle.so, our plain-C restatement
* ``ref_path()``-> oracle/_ref/libaecm_ref.so, the unmodified reference compiled by oracle/Makefile
                   (present in the build container; travels to the GPU box as a built artefact)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libaecm_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libaecm_ref.so")


class RunConfig(C.Structure):
    _fields_ = [("n_instances", C.c_int32), ("fs", C.c_int32), ("samples_per_call", C.c_int32),
                ("n_calls", C.c_int32), ("cng_mode", C.c_int32), ("echo_mode", C.c_int32),
                ("use_clean", C.c_int32), ("n_threads", C.c_int32), ("skip_far_every", C.c_int32)]


_lib = None


def build():
    """(Re)build the checkers: the oracle always, the reference library when /root/reference exists."""
    subprocess.check_call(["make", "-s", "-f", os.path.join(ORACLE_DIR, "Makefile")], cwd=ROOT)


def oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        lib = C.CDLL(ORACLE_SO)
        lib.oracle_run_streams.restype = C.c_int
        lib.oracle_run_streams.argtypes = [C.POINTER(RunConfig), C.c_int, C.c_char_p] + [C.c_void_p] * 6 + [
            C.POINTER(C.c_double)]
        lib.oracle_aecm_create.restype = C.c_void_p
        for name in ("oracle_aecm_free", "oracle_aecm_init", "oracle_aecm_set_config", "oracle_aecm_buffer_farend",
                     "oracle_aecm_get_buffer_farend_error", "oracle_aecm_process", "oracle_aecm_init_echo_path",
                     "oracle_aecm_get_echo_path"):
            getattr(lib, name).restype = C.c_int32
        lib.oracle_aecm_free.restype = None
        lib.oracle_aecm_free.argtypes = [C.c_void_p]
        lib.oracle_aecm_init.argtypes = [C.c_void_p, C.c_int32]
        lib.oracle_aecm_set_config.argtypes = [C.c_void_p, C.c_int16, C.c_int16]
        lib.oracle_aecm_buffer_farend.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_aecm_get_buffer_farend_error.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_aecm_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int16]
        lib.oracle_aecm_init_echo_path.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_aecm_get_echo_path.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_aecm_echo_path_size_bytes.restype = C.c_size_t
        lib.oracle_sqrt_floor.restype = C.c_int32
        lib.oracle_sqrt_floor.argtypes = [C.c_int32]
        lib.oracle_real_fft128_inverse.restype = C.c_int
        _lib = lib
    return _lib


def ref_path():
    return REF_SO if os.path.exists(REF_SO) else None


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run_streams(far, near, *, fs=16000, samples_per_call=None, impl="oracle", clean=None, ms=None, cng_mode=1,
                echo_mode=3, threads=None, skip_far_every=0):
    """Run [n, T] int16 planes through the oracle ('oracle') or the reference library ('ref').

    Returns (out[n, T] int16, ret_or[n] int32, seconds)."""
    far = np.ascontiguousarray(far, dtype=np.int16)
    near = np.ascontiguousarray(near, dtype=np.int16)
    n, total = far.shape
    spc = samples_per_call or (160 if fs == 16000 else 80)
    assert total % spc == 0
    cfg = RunConfig(n, fs, spc, total // spc, cng_mode, echo_mode, 0 if clean is None else 1,
                    threads or min(os.cpu_count() or 1, max(n, 1)), skip_far_every)
    if clean is not None:
        clean = np.ascontiguousarray(clean, dtype=np.int16)
    if ms is not None:
        ms = np.ascontiguousarray(np.broadcast_to(np.asarray(ms, dtype=np.int16), (n,)))
    out = np.zeros_like(near)
    ret = np.zeros(n, dtype=np.int32)
    secs = C.c_double(0)
    path = None
    if impl == "ref":
        path = ref_path()
        if path is None:
            raise FileNotFoundError(REF_SO)
        path = path.encode()
    rc = oracle().oracle_run_streams(C.byref(cfg), 1 if impl == "ref" else 0, path, _ptr(far), _ptr(near), _ptr(clean),
                                     _ptr(ms), _ptr(out), _ptr(ret), C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"oracle_run_streams rc={rc}")
    return out, ret, secs.value
