"""Cases that only make sense on the host model (tests/host_emu/host_model.c); run by tests/test_host_orchestration_model.py
in a subprocess with AECM_B200_LIB pointing at the model library.  Not collected by the normal test run (no test_ prefix)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from webrtc_aecm_b200 import AecmBatch, api
from webrtc_aecm_b200.synth import synth_streams

L = api.lib()
M = C.CDLL(api.library_path())  # the same library, for the model's own hooks
assert b"hostmodel" in api.library_path().encode(), "these cases are for the host model only"
L.AecmB200_HostAlloc.restype = C.c_void_p
L.AecmB200_HostAlloc.argtypes = [C.c_size_t]
L.AecmB200_HostFree.argtypes = [C.c_void_p]
M.host_model_alloc_touching.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
M.host_model_free_touching.argtypes = [C.c_void_p]
M.host_model_pending_ops.restype = C.c_long


def pinned(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.AecmB200_HostAlloc(n)
    assert p
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (n,)).view(dtype).reshape(shape)
    return a, p


def planes_by_call(x, n, calls, spc):
    return np.ascontiguousarray(x.reshape(n, calls, spc).transpose(1, 0, 2))


@pytest.mark.parametrize("wait_order", ["oldest_first", "newest_first"])
def test_pinned_planes_every_slot_in_flight(wait_order):
    """Pinned host planes are the truly asynchronous path: nothing of a submission has run when Submit returns (the model
    executes lazily).  All slots are filled, then waited for oldest-first or NEWEST-first -- the latter drains the youngest
    slot's stream first, so its kernel is only correct if the chain of kernel-order events pulls the older slots' work in
    before it."""
    fs, spc, n, rounds = 16000, 160, 24, 10
    S = L.AecmB200_NumStreamSlots()
    calls = rounds * S
    far, near = synth_streams(n, calls * spc, fs, seed=4242, delay_ms=np.arange(n) * 4 % 90)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle")
    hfar, pf = pinned((calls, n, spc), np.int16)
    hnear, pn = pinned((calls, n, spc), np.int16)
    hout, po = pinned((calls, n, spc), np.int16)
    hst, ps = pinned((S, n), np.int32)
    hfar[:] = planes_by_call(far, n, calls, spc)
    hnear[:] = planes_by_call(near, n, calls, spc)
    hout[:] = 0
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    for r in range(rounds):
        for s in range(S):
            k = r * S + s
            a = api.StreamArgs(hfar[k].ctypes.data, hnear[k].ctypes.data, None, hout[k].ctypes.data, spc, 0, n, spc, 1, None, 40, None,
                               hst[s].ctypes.data)
            assert L.AecmB200_SubmitStream(b._h, C.byref(a), s) == 0
        if os.environ.get("FAKE_CUDA_ORDER", "lazy") == "lazy":
            assert M.host_model_pending_ops() > 0 and not hout[r * S:(r + 1) * S].any()  # really nothing has run yet
        for s in (range(S) if wait_order == "oldest_first" else reversed(range(S))):
            assert L.AecmB200_WaitStream(b._h, s) == 0 and not hst[s].any()
    np.testing.assert_array_equal(hout.transpose(1, 0, 2).reshape(n, calls * spc), want)
    b.close()
    for p in (pf, pn, po, ps):
        L.AecmB200_HostFree(p)


@pytest.mark.parametrize("layout", ["one_allocation", "two_touching_allocations"])
def test_far_plane_followed_by_near_plane(layout):
    """near == far + plane: one DMA when both lie in one pinned allocation; when they are two allocations that merely touch,
    the runtime refuses the single copy (cudaErrorInvalidValue, modelled) and the library has to fall back to two."""
    fs, spc, n, calls = 16000, 160, 8, 30
    far, near = synth_streams(n, calls * spc, fs, seed=77, delay_ms=20)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle")
    plane = n * spc * 2
    if layout == "one_allocation":
        base = L.AecmB200_HostAlloc(2 * plane)
        pa, pb = base, base + plane
    else:
        a_, b_ = C.c_void_p(), C.c_void_p()
        assert M.host_model_alloc_touching(plane, C.byref(a_), C.byref(b_)) == 0
        pa, pb = a_.value, b_.value
        assert pb == pa + plane
    hfar = np.ctypeslib.as_array(C.cast(pa, C.POINTER(C.c_int16)), (n, spc))
    hnear = np.ctypeslib.as_array(C.cast(pb, C.POINTER(C.c_int16)), (n, spc))
    out = np.zeros((n, calls * spc), dtype=np.int16)
    o = np.zeros((n, spc), dtype=np.int16)
    st = np.zeros(n, dtype=np.int32)
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    for c in range(calls):
        hfar[:] = far[:, c * spc:(c + 1) * spc]
        hnear[:] = near[:, c * spc:(c + 1) * spc]
        a = api.StreamArgs(pa, pb, None, o.ctypes.data, spc, 0, n, spc, 1, None, 40, None, st.ctypes.data)
        assert L.AecmB200_ProcessStream(b._h, C.byref(a)) == 0 and not st.any()
        out[:, c * spc:(c + 1) * spc] = o
    np.testing.assert_array_equal(out, want)
    b.close()
    if layout == "one_allocation":
        L.AecmB200_HostFree(pa)
    else:
        M.host_model_free_touching(pa)


def test_a_failed_submission_leaves_the_slot_free_and_the_stream_intact():
    """cudaMalloc fails inside AecmB200_SubmitStream (staging planes are grown on demand): the call reports the failure, the
    slot is not left busy, WaitStream on it is a no-op that writes no stale status, and the next submission works."""
    fs, spc, n, calls = 16000, 160, 6, 20
    far, near = synth_streams(n, calls * spc, fs, seed=5, delay_ms=30)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle")
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    out = np.zeros_like(near)
    st = np.full(n, -7, dtype=np.int32)
    for c in range(calls):
        f = np.ascontiguousarray(far[:, c * spc:(c + 1) * spc])
        x = np.ascontiguousarray(near[:, c * spc:(c + 1) * spc])
        o = np.zeros_like(x)
        a = api.StreamArgs(f.ctypes.data, x.ctypes.data, None, o.ctypes.data, spc, 0, n, spc, 1, None, 40, None, st.ctypes.data)
        if c == 0:  # the very first submission has to allocate its staging planes: make that fail
            M.host_model_fail_malloc_in(1)
            assert L.AecmB200_SubmitStream(b._h, C.byref(a), 1) == api.AECM_UNSPECIFIED_ERROR
            assert L.AecmB200_WaitStream(b._h, 1) == 0 and (st == -7).all()
        assert L.AecmB200_SubmitStream(b._h, C.byref(a), 1) == 0
        assert L.AecmB200_WaitStream(b._h, 1) == 0 and not st.any()
        st[:] = -7
        out[:, c * spc:(c + 1) * spc] = o
    np.testing.assert_array_equal(out, want)
    b.close()


def test_the_callers_cuda_device_is_put_back():
    """Every entry point switches to the batch's device for its own CUDA calls and restores the caller's on every return path."""
    assert M.cudaSetDevice(1) == 0
    b = AecmBatch(4, 0)
    assert M.host_model_current_device() == 1
    assert b.init(16000) == 0 and M.host_model_current_device() == 1
    rc, st = b.set_config(1, 7)  # refused per instance, still a normal return
    assert rc == 0 and (st == api.AECM_BAD_PARAMETER_ERROR).all() and M.host_model_current_device() == 1
    x = np.zeros((4, 160), dtype=np.int16)
    assert b.buffer_farend(x)[0] == 0 and b.process(x)[0] == 0 and M.host_model_current_device() == 1
    assert L.AecmB200_SubmitStream(b._h, None, 0) == api.AECM_NULL_POINTER_ERROR and M.host_model_current_device() == 1
    b.close()
    assert M.host_model_current_device() == 1
    assert M.cudaSetDevice(0) == 0


def test_main_stream_calls_see_the_state_left_by_a_submission_in_flight():
    """A call that works on the batch's main stream (here AecmB200_ReadState) issued while a submission is still in flight on
    a slot stream must observe the state that submission leaves: the main stream has to wait for the slot's kernel."""
    fs, spc, n, calls = 16000, 160, 4, 24
    far, near = synth_streams(n, (calls + 1) * spc, fs, seed=11, delay_ms=10)
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    rc, _, st = b.process_stream(np.ascontiguousarray(far[:, :calls * spc]), np.ascontiguousarray(near[:, :calls * spc]))
    assert rc == 0 and not st.any()
    before = b.read_state(2).s.tot_count
    assert before > 0
    hfar, pf = pinned((n, spc), np.int16)
    hnear, pn = pinned((n, spc), np.int16)
    hout, po = pinned((n, spc), np.int16)
    hfar[:] = far[:, calls * spc:]
    hnear[:] = near[:, calls * spc:]
    a = api.StreamArgs(hfar.ctypes.data, hnear.ctypes.data, None, hout.ctypes.data, spc, 0, n, spc, 1, None, 40, None, None)
    assert L.AecmB200_SubmitStream(b._h, C.byref(a), 2) == 0
    during = b.read_state(2).s.tot_count  # not waited for yet
    assert during in (before + 2, before + 3), (before, during)
    assert L.AecmB200_WaitStream(b._h, 2) == 0
    assert b.read_state(2).s.tot_count == during
    b.close()
    for p in (pf, pn, po):
        L.AecmB200_HostFree(p)


def test_create_survives_an_allocation_failure_at_every_point():
    """AecmB200_Create makes about twenty device allocations; whichever of them fails, it must return NULL after releasing what it
    already had (the sanitizer build of this test checks the releasing), and a later Create must work."""
    L.AecmB200_Create.restype = C.c_void_p
    L.AecmB200_Free.argtypes = [C.c_void_p]
    failed = 0
    for k in range(1, 40):
        M.host_model_fail_malloc_in(k)
        h = L.AecmB200_Create(5, 0)
        if h:
            L.AecmB200_Free(h)
            break
        failed += 1
    M.host_model_fail_malloc_in(0)
    assert 10 <= failed < 39, failed
    b = AecmBatch(5, 0)
    assert b.init(16000) == 0
    b.close()


def test_group_create_fails_cleanly_when_a_shard_cannot_be_built():
    L.AecmB200_GroupCreate.restype = C.c_void_p
    L.AecmB200_GroupFree.argtypes = [C.c_void_p]
    devices = (C.c_int32 * 2)(0, 1)
    M.host_model_fail_malloc_in(25)  # inside the second shard's Create
    assert not L.AecmB200_GroupCreate(10, 2, devices)
    M.host_model_fail_malloc_in(0)
    assert not L.AecmB200_GroupCreate(1, 2, devices)  # more shards than instances
    bad = (C.c_int32 * 2)(0, 7)
    assert not L.AecmB200_GroupCreate(10, 2, bad)  # no such device
    g = L.AecmB200_GroupCreate(10, 2, devices)
    assert g
    L.AecmB200_GroupFree(g)


def test_device_plane_calls_interleaved_with_host_plane_submissions():
    """AecmB200_ProcessStreamDevice works on the batch's main stream, the host-plane calls on the slot streams; alternating them
    call by call crosses the hand-over in both directions every time.  (On the model any host array is "device" memory.)"""
    fs, spc, n, calls = 16000, 160, 5, 40
    far, near = synth_streams(n, calls * spc, fs, seed=321, delay_ms=np.arange(n) * 11)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle")
    b = AecmBatch(n, 0)
    assert L.AecmB200_NumInstances(b._h) == n and L.AecmB200_Device(b._h) == 0 and b.cuda_stream()
    L.AecmB200_LastError.restype = C.c_char_p
    assert L.AecmB200_LastError(b._h) == b"" and L.AecmB200_LastError(None) == b"null batch"
    assert b.init(fs) == 0
    out = np.zeros_like(near)
    st = np.zeros(n, dtype=np.int32)
    c = 0
    while c < calls:
        k = 1 + (c % 3 == 2 and c + 2 <= calls)  # sometimes two fused calls
        f = np.ascontiguousarray(far[:, c * spc:(c + k) * spc])
        x = np.ascontiguousarray(near[:, c * spc:(c + k) * spc])
        o = np.zeros_like(x)
        a = api.StreamArgs(f.ctypes.data, x.ctypes.data, None, o.ctypes.data, k * spc, 0, n, spc, k, None, 40, None, st.ctypes.data)
        if (c // 2) % 2:
            assert L.AecmB200_ProcessStreamDevice(b._h, C.byref(a)) == 0
            b.synchronize()
        else:
            slot = c % L.AecmB200_NumStreamSlots()
            assert L.AecmB200_SubmitStream(b._h, C.byref(a), slot) == 0
            assert L.AecmB200_WaitStream(b._h, slot) == 0
        assert not st.any()
        out[:, c * spc:(c + k) * spc] = o
        c += k
    np.testing.assert_array_equal(out, want)
    rc, d, q = b.delay_estimate()
    assert rc == 0 and len(d) == n and (d >= -2).all() and (q >= 0).all()
    b.close()


def test_trace_mode_runs(monkeypatch, capfd):
    monkeypatch.setenv("AECM_B200_TRACE", "1")
    b = AecmBatch(3, 0)
    assert b.init(16000) == 0
    x = np.zeros((3, 160), dtype=np.int16)
    rc, out, st = b.process_stream(x, x)
    assert rc == 0 and not st.any()
    b.close()
    assert "aecm_b200 trace slot" in capfd.readouterr().err


def test_scalar_create_fails_cleanly_when_the_pool_cannot_be_built():
    S = api.scalar_api()
    S.WebRtcAecm_Create.restype = C.c_void_p
    S.WebRtcAecm_Free.argtypes = [C.c_void_p]
    first = S.WebRtcAecm_Create()  # makes sure a pool exists, then fill it so that the next Create needs a new pool
    assert first
    rest = [S.WebRtcAecm_Create() for _ in range(255)]
    assert all(rest)
    M.host_model_fail_malloc_in(3)
    assert not S.WebRtcAecm_Create()
    M.host_model_fail_malloc_in(0)
    h = S.WebRtcAecm_Create()
    assert h
    for p in [first, h, *rest]:
        S.WebRtcAecm_Free(p)


def test_adopted_stream_launch_counter_and_selftests():
    """AecmB200_AdoptCudaStream: the batch moves to a stream the caller owns (here one made by the modelled runtime), which it
    must neither destroy nor outlive; plus the launch counter and the self-test entry points (the oracle answers on the model)."""
    s = C.c_void_p()
    assert M.cudaStreamCreateWithFlags(C.byref(s), 1) == 0
    L.AecmB200_KernelLaunchCount.restype = C.c_int64
    before = L.AecmB200_KernelLaunchCount()
    b = AecmBatch(3, 0)
    b.adopt_cuda_stream(s.value)
    assert b.cuda_stream() == s.value
    assert b.init(16000) == 0
    x = np.zeros((3, 160), dtype=np.int16)
    a = api.StreamArgs(x.ctypes.data, x.ctypes.data, None, np.zeros_like(x).ctypes.data, 160, 0, 3, 160, 1, None, 40, None, None)
    assert L.AecmB200_ProcessStreamDevice(b._h, C.byref(a)) == 0
    b.synchronize()
    b.close()
    M.cudaStreamDestroy.argtypes = [C.c_void_p]
    assert M.cudaStreamDestroy(s) == 0  # still the caller's to destroy
    assert L.AecmB200_KernelLaunchCount() >= before + 2
    L.AecmB200_SelfTestSqrtFloor.restype = C.c_int64
    assert L.AecmB200_SelfTestSqrtFloor(0, 1 << 24) == 0
