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
REF_SHIM_SO = os.path.join(ORACLE_DIR, "_ref", "libaecm_ref_shim.so")  # the same objects + oracle/ref_shim.cc (SURVEY 8-f4 knobs)


class RunConfig(C.Structure):
    _fields_ = [("n_instances", C.c_int32), ("fs", C.c_int32), ("samples_per_call", C.c_int32),
                ("n_calls", C.c_int32), ("cng_mode", C.c_int32), ("echo_mode", C.c_int32),
                ("use_clean", C.c_int32), ("n_threads", C.c_int32), ("skip_far_every", C.c_int32),
                ("ms_per_call", C.c_int32), ("extra_far_every", C.c_int32),
                ("dly_options", C.c_int32), ("dly_robust", C.c_int32), ("dly_allowed_offset", C.c_int32),
                ("dly_max_lookahead", C.c_int32), ("dly_lookahead", C.c_int32), ("soft_reset_every", C.c_int32),
                ("soft_reset_shift", C.c_int32)]


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
        lib.oracle_run_streams_ex.restype = C.c_int
        lib.oracle_run_streams_ex.argtypes = [C.POINTER(RunConfig), C.c_int, C.c_char_p] + [C.c_void_p] * 7 + [
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
        lib.oracle_pool_create.restype = C.c_void_p
        lib.oracle_pool_create.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.oracle_pool_run.restype = C.c_int
        lib.oracle_pool_run.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_int, C.c_int,
                                                                           C.POINTER(C.c_double)]
        lib.oracle_pool_run_tiled.restype = C.c_int
        lib.oracle_pool_run_tiled.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                              C.POINTER(C.c_double)]
        lib.oracle_pool_free.restype = None
        lib.oracle_pool_free.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def ref_path():
    return REF_SO if os.path.exists(REF_SO) else None


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ref_shim_path():
    return REF_SHIM_SO if os.path.exists(REF_SHIM_SO) else None


def run_streams(far, near, *, fs=16000, samples_per_call=None, impl="oracle", clean=None, ms=None, cng_mode=1,
                echo_mode=3, threads=None, skip_far_every=0, extra_far_every=0, dly=None, soft_reset=None,
                delay_trace=False):
    """Run [n, T] int16 planes through the oracle ('oracle') or the reference library ('ref').

    dly: None or dict(robust=, allowed_offset=, max_lookahead=, lookahead=) -- the delay estimator's knobs,
    applied after Init/set_config; soft_reset: None or (every_k_calls, shift).  With impl='ref' these (and delay_trace)
    go through the reference shim library.
    Returns (out[n, T] int16, ret_or[n] int32, seconds), plus delay_trace[n, n_calls] int32 when delay_trace=True."""
    far = np.ascontiguousarray(far, dtype=np.int16)
    near = np.ascontiguousarray(near, dtype=np.int16)
    n, total = far.shape
    spc = samples_per_call or (160 if fs == 16000 else 80)
    assert total % spc == 0
    per_call = ms is not None and np.ndim(ms) == 2  # [n, n_calls]: one msInSndCardBuf per call
    d = dly or {}
    sr = soft_reset or (0, 0)
    cfg = RunConfig(n, fs, spc, total // spc, cng_mode, echo_mode, 0 if clean is None else 1,
                    threads or min(os.cpu_count() or 1, max(n, 1)), skip_far_every, 1 if per_call else 0, extra_far_every,
                    1 if dly else 0, d.get("robust", 0), d.get("allowed_offset", 0), d.get("max_lookahead", -1),
                    d.get("lookahead", -1), sr[0], sr[1])
    needs_shim = bool(dly) or sr[0] > 0 or delay_trace
    if clean is not None:
        clean = np.ascontiguousarray(clean, dtype=np.int16)
    if per_call:
        ms = np.ascontiguousarray(ms, dtype=np.int16)
        assert ms.shape == (n, total // spc)
    elif ms is not None:
        ms = np.ascontiguousarray(np.broadcast_to(np.asarray(ms, dtype=np.int16), (n,)))
    out = np.zeros_like(near)
    ret = np.zeros(n, dtype=np.int32)
    secs = C.c_double(0)
    path = None
    if impl == "ref":
        path = ref_shim_path() if needs_shim else ref_path()
        if path is None:
            raise FileNotFoundError(REF_SHIM_SO if needs_shim else REF_SO)
        path = path.encode()
    trace = np.zeros((n, total // spc), dtype=np.int32) if delay_trace else None
    rc = oracle().oracle_run_streams_ex(C.byref(cfg), 1 if impl == "ref" else 0, path, _ptr(far), _ptr(near), _ptr(clean),
                                        _ptr(ms), _ptr(out), _ptr(ret), _ptr(trace), C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"oracle_run_streams rc={rc}")
    if delay_trace:
        return out, ret, secs.value, trace
    return out, ret, secs.value


# ---------------------------------------------------------------------------------------------------
# single-instance stepping + state mirror (stage-level parity / debugging)
# ---------------------------------------------------------------------------------------------------
class Fifo(C.Structure):
    _fields_ = [("cap", C.c_int32), ("rd", C.c_int32), ("cnt", C.c_int32)]


class OracleState(C.Structure):
    """Mirror of ``OracleAecm`` in oracle/aecm_oracle.h."""
    _fields_ = [
        ("fs", C.c_int32), ("init_flag", C.c_int16), ("buf_size_start", C.c_int16), ("counter", C.c_int16),
        ("sum", C.c_int16), ("first_val", C.c_int16), ("check_buf_size_ctr", C.c_int16),
        ("ms_in_snd_card_buf", C.c_int16), ("filt_delay", C.c_int16), ("last_delay_diff", C.c_int16),
        ("echo_mode", C.c_int16), ("time_for_delay_change", C.c_int32), ("ec_startup", C.c_int32),
        ("check_buff_size", C.c_int32), ("delay_change", C.c_int32), ("wrapper_known_delay", C.c_int32),
        ("farend_old", (C.c_int16 * 80) * 2), ("far_fifo", Fifo), ("far_ring", C.c_int16 * 4000),
        ("frm_far_fifo", Fifo), ("frm_near_fifo", Fifo), ("frm_clean_fifo", Fifo), ("frm_out_fifo", Fifo),
        ("frm_far", C.c_int16 * 144), ("frm_near", C.c_int16 * 144), ("frm_clean", C.c_int16 * 144),
        ("frm_out", C.c_int16 * 144),
        ("mult", C.c_int16), ("cng_mode", C.c_int16), ("seed", C.c_uint32), ("tot_count", C.c_uint32),
        ("first_vad", C.c_int32), ("q_clean", C.c_int16), ("q_clean_old", C.c_int16), ("q_noisy", C.c_int16),
        ("q_noisy_old", C.c_int16), ("x_buf", C.c_int16 * 128), ("d_noisy", C.c_int16 * 128),
        ("d_clean", C.c_int16 * 128), ("out_ola", C.c_int16 * 64), ("ch_stored", C.c_int16 * 65),
        ("ch_adapt16", C.c_int16 * 65), ("ch_adapt32", C.c_int32 * 65), ("echo_filt", C.c_int32 * 65),
        ("near_filt", C.c_int16 * 65), ("noise_est", C.c_int32 * 65), ("noise_low_ctr", C.c_int32 * 65),
        ("noise_high_ctr", C.c_int32 * 65), ("noise_est_ctr", C.c_int16), ("near_log", C.c_int16 * 64),
        ("echo_adapt_log", C.c_int16 * 64), ("echo_stored_log", C.c_int16 * 64), ("far_log", C.c_int16),
        ("mse_adapt_old", C.c_int32), ("mse_stored_old", C.c_int32), ("mse_threshold", C.c_int32),
        ("far_e_min", C.c_int16), ("far_e_max", C.c_int16), ("far_e_maxmin", C.c_int16), ("far_e_vad", C.c_int16),
        ("far_e_mse", C.c_int16), ("current_vad", C.c_int32), ("vad_update_count", C.c_int16),
        ("startup_state", C.c_int16), ("mse_channel_count", C.c_int16), ("sup_gain", C.c_int16),
        ("sup_gain_old", C.c_int16), ("sup_a", C.c_int16), ("sup_d", C.c_int16), ("sup_diff_ab", C.c_int16),
        ("sup_diff_bd", C.c_int16), ("far_hist", C.c_uint16 * 6500), ("far_hist_pos", C.c_int32),
        ("far_hist_q", C.c_int32 * 100), ("thr_far", C.c_int32 * 65), ("thr_near", C.c_int32 * 65),
        ("thr_far_init", C.c_int32), ("thr_near_init", C.c_int32), ("bin_far", C.c_uint32 * 100),
        ("bin_far_head", C.c_int32), ("mean_bit_counts", C.c_int32 * 101), ("minimum_probability", C.c_int32),
        ("last_delay_probability", C.c_int32), ("last_delay", C.c_int32),
        ("robust_enabled", C.c_int32), ("allowed_offset", C.c_int32), ("near_history_size", C.c_int32), ("lookahead", C.c_int32),
        ("bin_near_hist", C.c_uint32 * 16), ("histogram", C.c_float * 101), ("last_delay_histogram", C.c_float),
        ("candidate_hits", C.c_int32), ("last_candidate_delay", C.c_int32), ("compare_delay", C.c_int32)]


class OracleInstance:
    """One oracle instance driven call by call (the reference protocol of main.cc:111-144)."""

    def __init__(self, fs=16000, cng_mode=1, echo_mode=3):
        self.lib = oracle()
        assert self.lib.oracle_aecm_sizeof() == C.sizeof(OracleState), "OracleState mirror out of date"
        self.h = self.lib.oracle_aecm_create()
        self.rc_init = self.lib.oracle_aecm_init(self.h, fs)
        self.rc_cfg = self.lib.oracle_aecm_set_config(self.h, cng_mode, echo_mode)

    def buffer_farend(self, far):
        far = np.ascontiguousarray(far, dtype=np.int16)
        return self.lib.oracle_aecm_buffer_farend(self.h, _ptr(far), far.size)

    def process(self, near, clean=None, ms=40):
        near = np.ascontiguousarray(near, dtype=np.int16)
        out = np.zeros_like(near)
        if clean is not None:
            clean = np.ascontiguousarray(clean, dtype=np.int16)
        rc = self.lib.oracle_aecm_process(self.h, _ptr(near), _ptr(clean), _ptr(out), near.size, int(ms))
        return rc, out

    def state(self) -> OracleState:
        return OracleState.from_address(self.h)

    def set_delay_estimator_options(self, robust=0, allowed_offset=0, max_lookahead=-1, lookahead=-1):
        f = self.lib.oracle_aecm_set_delay_estimator_options
        f.argtypes = [C.c_void_p] + [C.c_int] * 4
        return f(self.h, robust, allowed_offset, max_lookahead, lookahead)

    def delay_estimate(self):
        """(WebRtc_last_delay, WebRtc_last_delay_quality) of the instance."""
        q = self.lib.oracle_aecm_last_delay_quality
        q.restype, q.argtypes = C.c_float, [C.c_void_p]
        d = self.lib.oracle_aecm_last_delay
        d.restype, d.argtypes = C.c_int32, [C.c_void_p]
        return int(d(self.h)), float(q(self.h))

    def close(self):
        if self.h:
            self.lib.oracle_aecm_free(self.h)
            self.h = None


def compare_states(dev, orc: OracleState, use_clean=False):
    """Field-by-field comparison of a device ``InstanceState`` with an ``OracleState``.
    Returns a list of human-readable mismatch descriptions (empty = identical live state)."""
    bad = []

    def arr(x, n=None):
        a = np.ctypeslib.as_array(x)
        return a if n is None else a[:n]

    def chk(name, d, o):
        d = np.asarray(d).astype(np.int64).ravel()
        o = np.asarray(o).astype(np.int64).ravel()
        if d.shape != o.shape or (d != o).any():
            idx = np.nonzero(d != o)[0][:4] if d.shape == o.shape else []
            bad.append(f"{name}: dev{d[idx].tolist() if len(idx) else d[:4].tolist()} "
                       f"oracle{o[idx].tolist() if len(idx) else o[:4].tolist()} at {list(idx)}")

    S = dev.s
    for f in ("ch_adapt32", "echo_filt", "noise_est", "ch_stored", "ch_adapt16", "near_filt", "noise_low_ctr",
              "noise_high_ctr"):
        chk(f, arr(getattr(dev, f), 65), arr(getattr(orc, f)))
    chk("x_old", arr(dev.x_old), arr(orc.x_buf)[:64])
    chk("d_noisy_old", arr(dev.d_noisy_old), arr(orc.d_noisy)[:64])
    if use_clean:
        chk("d_clean_old", arr(dev.d_clean_old), arr(orc.d_clean)[:64])
    chk("out_ola", arr(dev.out_ola), arr(orc.out_ola))
    chk("thr_far", arr(dev.thr_far), arr(orc.thr_far)[12:44])
    chk("thr_near", arr(dev.thr_near), arr(orc.thr_near)[12:44])
    dh, oh = S.bin_far_head, orc.bin_far_head
    chk("bin_far(by delay)", np.roll(arr(dev.bin_far, 100), -dh), np.roll(arr(orc.bin_far), -oh))
    chk("mean_bit_counts", arr(dev.mean_bit_counts, 101), arr(orc.mean_bit_counts))  # incl. the dummy element [100]
    k = (S.log_head + np.arange(20)) & 31
    for f in ("near_log", "echo_adapt_log", "echo_stored_log"):
        chk(f, arr(getattr(dev, f))[k], arr(getattr(orc, f))[:20])
    # FIFOs: compare live contents
    def live(store, rd, cnt, cap):
        return np.asarray(store)[(rd + np.arange(cnt)) % cap]
    chk("frm_in_cnt", S.frm_in_cnt, orc.frm_far_fifo.cnt)
    chk("frm_far live", live(arr(dev.frm_far), S.frm_in_rd, S.frm_in_cnt, 144),
        live(arr(orc.frm_far), orc.frm_far_fifo.rd, orc.frm_far_fifo.cnt, 144))
    chk("frm_near live", live(arr(dev.frm_near), S.frm_in_rd, S.frm_in_cnt, 144),
        live(arr(orc.frm_near), orc.frm_near_fifo.rd, orc.frm_near_fifo.cnt, 144))
    chk("frm_out_cnt", S.frm_out_cnt, orc.frm_out_fifo.cnt)
    chk("frm_out live", live(arr(dev.frm_out), S.frm_out_rd, S.frm_out_cnt, 144),
        live(arr(orc.frm_out), orc.frm_out_fifo.rd, orc.frm_out_fifo.cnt, 144))
    chk("farend_old", arr(dev.farend_old), arr(orc.farend_old))
    chk("far_cnt", S.far_cnt, orc.far_fifo.cnt)
    chk("far_rd", S.far_rd, orc.far_fifo.rd)
    for f in ("mult", "cng_mode", "seed", "tot_count", "first_vad", "q_clean", "q_clean_old", "q_noisy", "q_noisy_old",
              "noise_est_ctr", "far_log", "mse_adapt_old", "mse_stored_old", "mse_threshold", "far_e_min", "far_e_max",
              "far_e_maxmin", "far_e_vad", "far_e_mse", "current_vad", "vad_update_count", "startup_state",
              "mse_channel_count", "sup_gain", "sup_gain_old", "sup_a", "sup_d", "sup_diff_ab", "sup_diff_bd",
              "far_hist_pos", "thr_far_init", "thr_near_init", "minimum_probability", "last_delay_probability",
              "last_delay", "fs", "init_flag", "echo_mode", "buf_size_start", "counter", "sum", "first_val",
              "check_buf_size_ctr", "ms_in_snd_card_buf", "ec_startup", "check_buff_size", "filt_delay",
              "last_delay_diff", "time_for_delay_change", "delay_change", "wrapper_known_delay"):
        chk(f, getattr(S, f), getattr(orc, f))
    return bad


class CpuPool:
    """Persistent multi-threaded CPU instances (oracle or reference library) for benchmarking."""

    def __init__(self, n, fs=16000, impl="ref", cng_mode=1, echo_mode=3, threads=None):
        self.lib = oracle()
        self.n, self.fs, self.impl = n, fs, impl
        self.threads = threads or (os.cpu_count() or 1)
        path = None
        if impl == "ref":
            if ref_path() is None:
                raise FileNotFoundError(REF_SO)
            path = ref_path().encode()
        self.h = self.lib.oracle_pool_create(1 if impl == "ref" else 0, path, n, fs, cng_mode, echo_mode, self.threads)
        if not self.h:
            raise RuntimeError("oracle_pool_create failed")

    def run(self, far, near, clean=None, samples_per_call=160, ms=40):
        """far/near: [n, T] int16 C-contiguous; returns (out, or_of_return_codes, seconds)."""
        assert far.flags["C_CONTIGUOUS"] and near.flags["C_CONTIGUOUS"] and far.shape[0] == self.n
        out = np.empty_like(near)
        secs = C.c_double(0)
        acc = self.lib.oracle_pool_run(self.h, _ptr(far), _ptr(near), _ptr(clean), _ptr(out), far.shape[1],
                                       samples_per_call, far.shape[1] // samples_per_call, ms, C.byref(secs))
        return out, acc, secs.value

    def run_tiled(self, far, near, samples_per_call=160, ms=40):
        """Throughput runs: far/near hold R <= n distinct streams [R, T]; instance i is fed row i % R and its output is
        dropped.  Returns (or_of_return_codes, seconds)."""
        assert far.flags["C_CONTIGUOUS"] and near.flags["C_CONTIGUOUS"] and far.shape == near.shape
        secs = C.c_double(0)
        acc = self.lib.oracle_pool_run_tiled(self.h, _ptr(far), _ptr(near), far.shape[0], far.shape[1], samples_per_call,
                                             far.shape[1] // samples_per_call, ms, C.byref(secs))
        return acc, secs.value

    def close(self):
        if self.h:
            self.lib.oracle_pool_free(self.h)
            self.h = None
