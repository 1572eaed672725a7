"""ctypes mirror of include/aecm_b200.h and include/echo_control_mobile.h.

Names and argument meaning follow the C ABI one to one (which in turn follows the reference's
``aecm/echo_control_mobile.h``); see those headers for the contract.  No computation happens here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

AECM_UNSPECIFIED_ERROR = 12000
AECM_UNSUPPORTED_FUNCTION_ERROR = 12001
AECM_UNINITIALIZED_ERROR = 12002
AECM_NULL_POINTER_ERROR = 12003
AECM_BAD_PARAMETER_ERROR = 12004
AECM_BAD_PARAMETER_WARNING = 12100

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path() -> str:
    """The in-tree build; AECM_B200_LIB overrides it (kernel-variant experiments, same ABI)."""
    return os.environ.get("AECM_B200_LIB") or os.path.join(_HERE, "libaecm_b200.so")


class AecmConfig(C.Structure):
    """``AecmConfig`` of echo_control_mobile.h:32-35."""
    _fields_ = [("cngMode", C.c_int16), ("echoMode", C.c_int16)]


class StreamArgs(C.Structure):
    """``AecmB200StreamArgs`` of aecm_b200.h."""
    _fields_ = [("farend", C.c_void_p), ("nearendNoisy", C.c_void_p), ("nearendClean", C.c_void_p), ("out", C.c_void_p),
                ("rowStride", C.c_int64), ("first", C.c_int32), ("count", C.c_int32), ("nrOfSamples", C.c_int32),
                ("nCalls", C.c_int32), ("msInSndCardBuf", C.c_void_p), ("msUniform", C.c_int16),
                ("farendMask", C.c_void_p), ("status", C.c_void_p), ("instanceIndex", C.c_void_p)]


class DelayEstimatorOptions(C.Structure):
    """``AecmB200DelayEstimatorOptions`` of aecm_b200.h."""
    _fields_ = [("robustValidation", C.c_int32), ("allowedOffset", C.c_int32), ("maxLookahead", C.c_int32), ("lookahead", C.c_int32)]


class Scalars(C.Structure):
    """``AecmScalars`` of csrc/aecm_state.h (introspection only)."""
    _fields_ = [(n, C.c_int32) for n in (
        "mult", "cng_mode")] + [("seed", C.c_uint32), ("tot_count", C.c_uint32)] + [(n, C.c_int32) for n in (
            "first_vad", "q_clean", "q_clean_old", "q_noisy", "q_noisy_old", "noise_est_ctr", "far_log",
            "mse_adapt_old", "mse_stored_old", "mse_threshold", "far_e_min", "far_e_max", "far_e_maxmin", "far_e_vad",
            "far_e_mse", "current_vad", "vad_update_count", "startup_state", "mse_channel_count", "sup_gain",
            "sup_gain_old", "sup_a", "sup_d", "sup_diff_ab", "sup_diff_bd", "far_hist_pos", "log_head", "thr_far_init",
            "thr_near_init", "bin_far_head", "minimum_probability", "last_delay_probability", "last_delay",
            "frm_in_rd", "frm_in_cnt", "frm_clean_rd", "frm_clean_cnt", "frm_out_rd", "frm_out_cnt", "fs", "init_flag",
            "echo_mode", "buf_size_start", "counter", "sum", "first_val", "check_buf_size_ctr", "ms_in_snd_card_buf",
            "ec_startup", "check_buff_size", "far_rd", "far_cnt", "filt_delay", "last_delay_diff",
            "time_for_delay_change", "delay_change", "wrapper_known_delay")] + [("pad_", C.c_int32 * 3)]


class InstanceState(C.Structure):
    """``AecmInstanceState`` of csrc/aecm_state.h (introspection only)."""
    _fields_ = [("s", Scalars), ("farend_old", (C.c_int16 * 80) * 2), ("frm_far", C.c_int16 * 144),
                ("frm_near", C.c_int16 * 144), ("frm_out", C.c_int16 * 144), ("x_old", C.c_int16 * 64),
                ("d_noisy_old", C.c_int16 * 64), ("out_ola", C.c_int16 * 64), ("thr_far", C.c_int32 * 32),
                ("thr_near", C.c_int32 * 32), ("bin_far", C.c_uint32 * 104), ("mean_bit_counts", C.c_int32 * 104),
                ("near_log", C.c_int16 * 32), ("echo_adapt_log", C.c_int16 * 32), ("echo_stored_log", C.c_int16 * 32),
                ("ch_adapt32", C.c_int32 * 68), ("echo_filt", C.c_int32 * 68), ("noise_est", C.c_int32 * 68),
                ("ch_stored", C.c_int16 * 72), ("ch_adapt16", C.c_int16 * 72), ("near_filt", C.c_int16 * 72),
                ("noise_low_ctr", C.c_uint8 * 80), ("noise_high_ctr", C.c_uint8 * 80), ("d_clean_old", C.c_int16 * 64),
                ("frm_clean", C.c_int16 * 144)]


_lib = None


def lib():
    """Load libaecm_b200.so (built in-tree by ``__graft_entry__.build()`` / csrc/Makefile). No fallback."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the AECM engine has no CPU implementation)")
        L = C.CDLL(path)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        sig = {
            "AecmB200_Create": (vp, [i32, i32]), "AecmB200_Free": (None, [vp]),
            "AecmB200_NumInstances": (i32, [vp]), "AecmB200_Device": (i32, [vp]),
            "AecmB200_LastError": (C.c_char_p, [vp]),
            "AecmB200_Reset": (i32, [vp, i32, i32]), "AecmB200_Init": (i32, [vp, i32, i32, i32]),
            "AecmB200_SetConfig": (i32, [vp, i32, i32, AecmConfig, vp]),
            "AecmB200_InitEchoPath": (i32, [vp, i32, vp, C.c_size_t]),
            "AecmB200_GetEchoPath": (i32, [vp, i32, vp, C.c_size_t]),
            "AecmB200_ProcessStream": (i32, [vp, C.POINTER(StreamArgs)]),
            "AecmB200_ProcessStreamDevice": (i32, [vp, C.POINTER(StreamArgs)]),
            "AecmB200_SubmitStream": (i32, [vp, C.POINTER(StreamArgs), i32]),
            "AecmB200_WaitStream": (i32, [vp, i32]),
            "AecmB200_GroupCreate": (vp, [i32, i32, vp]), "AecmB200_GroupFree": (None, [vp]),
            "AecmB200_GroupNumShards": (i32, [vp]),
            "AecmB200_GroupShard": (vp, [vp, i32, C.POINTER(i32), C.POINTER(i32)]),
            "AecmB200_GroupInit": (i32, [vp, i32]), "AecmB200_GroupSetConfig": (i32, [vp, AecmConfig, vp]),
            "AecmB200_GroupProcessStream": (i32, [vp, C.POINTER(StreamArgs)]),
            "AecmB200_GroupSubmitStream": (i32, [vp, C.POINTER(StreamArgs), i32]), "AecmB200_GroupWaitStream": (i32, [vp, i32]),
            "AecmB200_BufferFarend": (i32, [vp, vp, C.c_size_t, vp]),
            "AecmB200_Process": (i32, [vp, vp, vp, vp, C.c_size_t, vp, C.c_int16, vp]),
            "AecmB200_Synchronize": (i32, [vp]), "AecmB200_CudaStream": (vp, [vp]),
            "AecmB200_AdoptCudaStream": (i32, [vp, vp]),
            "AecmB200_HostAlloc": (vp, [C.c_size_t]), "AecmB200_HostFree": (None, [vp]),
            "AecmB200_SelfTestSqrtFloor": (C.c_int64, [C.c_uint32, C.c_uint32]), "AecmB200_NumStreamSlots": (i32, []),
            "AecmB200_SetDelayEstimatorOptions": (i32, [vp, i32, i32, C.POINTER(DelayEstimatorOptions), vp]),
            "AecmB200_SoftResetDelayEstimator": (i32, [vp, i32, i32, i32, vp]),
            "AecmB200_GetDelayEstimate": (i32, [vp, i32, i32, vp, vp]),
            "AecmB200_SelfTestFftForward": (i32, [vp, vp, i32, i32]), "AecmB200_SelfTestFftInverse": (i32, [vp, vp, vp, i32]),
            "AecmB200_ScalarApiStats": (None, [vp, C.POINTER(i64), C.POINTER(i64)]),
            "AecmB200_StateBytes": (C.c_size_t, []), "AecmB200_ReadState": (i32, [vp, i32, vp, C.c_size_t]),
            "AecmB200_KernelLaunchCount": (i64, []), "AecmB200_DeviceCount": (i32, []),
            # the reference's scalar API
            "WebRtcAecm_Create": (vp, []), "WebRtcAecm_Free": (None, [vp]), "WebRtcAecm_Init": (i32, [vp, i32]),
            "WebRtcAecm_BufferFarend": (i32, [vp, vp, C.c_size_t]),
            "WebRtcAecm_GetBufferFarendError": (i32, [vp, vp, C.c_size_t]),
            "WebRtcAecm_Process": (i32, [vp, vp, vp, vp, C.c_size_t, C.c_int16]),
            "WebRtcAecm_set_config": (i32, [vp, AecmConfig]),
            "WebRtcAecm_InitEchoPath": (i32, [vp, vp, C.c_size_t]),
            "WebRtcAecm_GetEchoPath": (i32, [vp, vp, C.c_size_t]),
            "WebRtcAecm_echo_path_size_bytes": (C.c_size_t, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        assert L.AecmB200_StateBytes() == C.sizeof(InstanceState), "InstanceState mirror out of date"
        _lib = L
    return _lib


def scalar_api():
    """The loaded library, for calling the ten WebRtcAecm_* symbols directly."""
    return lib()


def _host_ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def _dev_ptr(t):
    """Device pointer of a torch tensor / int / None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(int(t))


class AecmBatch:
    """N AECM instances on one GPU; methods map 1:1 onto the ``AecmB200_*`` C entry points."""

    def __init__(self, n_instances: int, device: int = 0):
        self._lib = lib()
        self._h = self._lib.AecmB200_Create(int(n_instances), int(device))
        if not self._h:
            raise RuntimeError("AecmB200_Create failed (no CUDA device? the engine has no CPU path)")
        self.n = int(n_instances)
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.AecmB200_Free(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} -> {rc} ({self._lib.AecmB200_LastError(self._h).decode()})")

    def reset(self, first=0, count=None):
        self._check(self._lib.AecmB200_Reset(self._h, first, self.n - first if count is None else count), "Reset")

    def init(self, samp_freq: int, first=0, count=None) -> int:
        return self._lib.AecmB200_Init(self._h, first, self.n - first if count is None else count, int(samp_freq))

    def set_config(self, cng_mode=1, echo_mode=3, first=0, count=None):
        count = self.n - first if count is None else count
        status = np.zeros(count, dtype=np.int32)
        rc = self._lib.AecmB200_SetConfig(self._h, first, count, AecmConfig(cng_mode, echo_mode), _host_ptr(status))
        return rc, status

    def init_echo_path(self, instance: int, path) -> int:
        path = np.ascontiguousarray(path, dtype=np.int16)
        return self._lib.AecmB200_InitEchoPath(self._h, instance, _host_ptr(path), path.nbytes)

    def get_echo_path(self, instance: int):
        path = np.zeros(65, dtype=np.int16)
        rc = self._lib.AecmB200_GetEchoPath(self._h, instance, _host_ptr(path), path.nbytes)
        return rc, path

    def process_stream(self, far, near, clean=None, *, samples_per_call=160, ms=None, ms_uniform=40, far_mask=None,
                       first=0, count=None, n_calls=None):
        """Host planes [count, T] int16 -> (rc, out[count, T], status[count])."""
        ref = near if near is not None else far
        count = ref.shape[0] if count is None else count
        total = ref.shape[1]
        n_calls = total // samples_per_call if n_calls is None else n_calls
        out = np.zeros_like(near) if near is not None else None
        status = np.zeros(count, dtype=np.int32)
        if ms is not None:
            ms = np.ascontiguousarray(np.broadcast_to(np.asarray(ms, dtype=np.int16), (count,)))
        if far_mask is not None:
            far_mask = np.ascontiguousarray(far_mask, dtype=np.uint8)
            assert far_mask.shape == (n_calls, count)
        a = StreamArgs(_host_ptr(far), _host_ptr(near), _host_ptr(clean), _host_ptr(out), total, first, count,
                       samples_per_call, n_calls, _host_ptr(ms), int(ms_uniform), _host_ptr(far_mask), _host_ptr(status))
        rc = self._lib.AecmB200_ProcessStream(self._h, C.byref(a))
        return rc, out, status

    def process_stream_device(self, d_far, d_near, d_clean, d_out, *, row_stride, samples_per_call, n_calls, count=None,
                              first=0, d_ms=None, ms_uniform=40, d_far_mask=None, d_status=None, sample_offset=0):
        """Device planes (torch tensors or raw pointers); asynchronous on the batch stream."""
        count = self.n - first if count is None else count
        off = int(sample_offset) * 2

        def p(t):
            q = _dev_ptr(t)
            return None if q is None else C.c_void_p(q.value + off)

        a = StreamArgs(p(d_far), p(d_near), p(d_clean), p(d_out), int(row_stride), first, count, samples_per_call, n_calls,
                       _dev_ptr(d_ms), int(ms_uniform), _dev_ptr(d_far_mask), _dev_ptr(d_status))
        return self._lib.AecmB200_ProcessStreamDevice(self._h, C.byref(a))

    def buffer_farend(self, far):
        status = np.zeros(self.n, dtype=np.int32)
        rc = self._lib.AecmB200_BufferFarend(self._h, _host_ptr(far), far.shape[1], _host_ptr(status))
        return rc, status

    def process(self, near, clean=None, ms=None, ms_uniform=40):
        out = np.zeros_like(near)
        status = np.zeros(self.n, dtype=np.int32)
        if ms is not None:
            ms = np.ascontiguousarray(np.broadcast_to(np.asarray(ms, dtype=np.int16), (self.n,)))
        rc = self._lib.AecmB200_Process(self._h, _host_ptr(near), _host_ptr(clean), _host_ptr(out), near.shape[1],
                                        _host_ptr(ms), int(ms_uniform), _host_ptr(status))
        return rc, out, status

    def set_delay_estimator_options(self, robust=0, allowed_offset=0, max_lookahead=-1, lookahead=-1, first=0, count=None):
        count = self.n - first if count is None else count
        status = np.zeros(count, dtype=np.int32)
        o = DelayEstimatorOptions(robust, allowed_offset, max_lookahead, lookahead)
        rc = self._lib.AecmB200_SetDelayEstimatorOptions(self._h, first, count, C.byref(o), _host_ptr(status))
        return rc, status

    def soft_reset_delay_estimator(self, delay_shift, first=0, count=None):
        count = self.n - first if count is None else count
        applied = np.zeros(count, dtype=np.int32)
        rc = self._lib.AecmB200_SoftResetDelayEstimator(self._h, first, count, int(delay_shift), _host_ptr(applied))
        return rc, applied

    def delay_estimate(self, first=0, count=None):
        count = self.n - first if count is None else count
        delay, quality = np.zeros(count, dtype=np.int32), np.zeros(count, dtype=np.float32)
        rc = self._lib.AecmB200_GetDelayEstimate(self._h, first, count, _host_ptr(delay), _host_ptr(quality))
        return rc, delay, quality

    def synchronize(self):
        self._check(self._lib.AecmB200_Synchronize(self._h), "Synchronize")

    def cuda_stream(self) -> int:
        return int(self._lib.AecmB200_CudaStream(self._h) or 0)

    def adopt_cuda_stream(self, stream_ptr: int):
        self._check(self._lib.AecmB200_AdoptCudaStream(self._h, C.c_void_p(int(stream_ptr))), "AdoptCudaStream")

    def read_state(self, instance: int) -> InstanceState:
        st = InstanceState()
        self._check(self._lib.AecmB200_ReadState(self._h, instance, C.byref(st), C.sizeof(st)), "ReadState")
        return st
