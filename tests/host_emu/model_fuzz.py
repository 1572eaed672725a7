"""Model-based fuzzing of the batched C ABI on the host model (run by tests/test_host_orchestration_model.py in a subprocess
with AECM_B200_LIB = the model library; seeds from MODEL_FUZZ_SEEDS).  A random sequence of API calls -- valid ones and ones
the library has to refuse -- goes to a small batch; a Python model that owns one oracle instance per batch instance and
knows the header's contract (include/aecm_b200.h) predicts every return code, every per-row status and every output sample.
What is exercised is the host layer's bookkeeping: ranges, instanceIndex (permutations, duplicates, out of range), row
strides, fused calls, far-end masks, per-row msInSndCardBuf, pageable and pinned planes, slots in flight and waited for in
any order, Reset / Init / set_config / echo-path calls in between."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from webrtc_aecm_b200 import AecmBatch, api

L = api.lib()
assert b"hostmodel" in api.library_path().encode(), "for the host model only"
L.AecmB200_HostAlloc.restype = C.c_void_p
L.AecmB200_HostAlloc.argtypes = [C.c_size_t]
L.AecmB200_HostFree.argtypes = [C.c_void_p]
E_UNSPEC, E_UNINIT, E_NULL, E_BAD, W_BAD = 12000, 12002, 12003, 12004, 12100
N = 7
SEEDS = [int(s) for s in os.environ.get("MODEL_FUZZ_SEEDS", "1,2,3").split(",")]


class Planes:
    """Host arrays of one call, pageable (numpy) or pinned (AecmB200_HostAlloc), kept alive until the call was waited for."""

    def __init__(self, pin):
        self.pin, self.ptrs, self.keep = pin, [], []

    def array(self, shape, dtype, fill=None):
        if self.pin:
            nbytes = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
            p = L.AecmB200_HostAlloc(nbytes)
            assert p
            self.ptrs.append(p)
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (nbytes,))[:int(np.prod(shape)) * np.dtype(dtype).itemsize].view(dtype).reshape(shape)
        else:
            a = np.zeros(shape, dtype=dtype)
        if fill is not None:
            a[...] = fill
        self.keep.append(a)
        return a

    def release(self):
        for p in self.ptrs:
            L.AecmB200_HostFree(p)
        self.ptrs, self.keep = [], []


@pytest.mark.parametrize("seed", SEEDS)
def test_random_api_sequences_against_the_contract(seed):
    rng = np.random.default_rng(seed)
    b = AecmBatch(N, 0)
    model = [ol.OracleInstance.__new__(ol.OracleInstance) for _ in range(N)]

    def fresh(i):  # = AecmB200_Create / _Reset: zero-filled, un-initialised
        o = model[i]
        if getattr(o, "h", None):
            o.close()
        o.lib = ol.oracle()
        o.h = o.lib.oracle_aecm_create()

    for i in range(N):
        fresh(i)
    assert b.init(16000, first=0, count=N - 1) == 0  # the last instance starts un-initialised
    for i in range(N - 1):
        assert model[i].lib.oracle_aecm_init(model[i].h, 16000) == 0
    n_slots = L.AecmB200_NumStreamSlots()
    stats = dict(ok=0, refused=0, busy=0, rows=0, warn=0, uninit=0, max_in_flight=0)
    in_flight = {}  # slot -> (Planes, out array view, want_out, mask of rows to compare, status array, want_status)
    use_clean = bool(rng.random() < 0.4)
    audio_t = 0

    def audio(shape):
        nonlocal audio_t
        audio_t += 1
        return (rng.integers(-3000, 3000, shape) * (1 + (audio_t % 5 == 0) * 6)).astype(np.int16)

    def finish(slot):
        pl, out, want_out, cmp_rows, st, want_st = in_flight.pop(slot)
        if st is not None:
            np.testing.assert_array_equal(st, want_st, err_msg=f"status, slot {slot}")
        if out is not None:
            for r, cols in cmp_rows:
                np.testing.assert_array_equal(out[r, cols], want_out[r, cols], err_msg=f"row {r}")
        pl.release()

    def stream_call(sync):
        pin = bool(rng.random() < 0.5)
        pl = Planes(pin)
        spc = int(rng.choice([80, 160]))
        n_calls = int(rng.integers(1, 4))
        width = spc * n_calls
        stride = width + int(rng.choice([0, 0, 16, 37]))
        use_index = rng.random() < 0.4
        if use_index:
            count = int(rng.integers(0, N + 1))
            index = rng.permutation(N)[:count].astype(np.int32)
            first = 0
        else:
            first = int(rng.integers(0, N))
            count = int(rng.integers(0, N - first + 1))
            index = None
        rows = list(index) if use_index else list(range(first, first + count))
        has_far = rng.random() < 0.85
        has_near = rng.random() < 0.85 or not has_far
        far = pl.array((count, stride), np.int16, audio((count, stride))) if has_far else None
        near = pl.array((count, stride), np.int16, audio((count, stride))) if has_near else None
        clean = pl.array((count, stride), np.int16, audio((count, stride))) if has_near and use_clean else None
        out = pl.array((count, stride), np.int16, 0) if has_near else None
        ms = pl.array((count,), np.int16, rng.choice([0, 20, 40, 120, 500, -3, 600], size=count)) if rng.random() < 0.5 and count else None
        ms_uniform = int(rng.choice([0, 40, 250, 501]))
        mask = pl.array((n_calls, count), np.uint8, rng.random((n_calls, count)) < 0.8) if has_far and rng.random() < 0.4 and count else None
        status = pl.array((count,), np.int32, -9) if rng.random() < 0.8 and count else None
        idx_arr = pl.array((count,), np.int32, index) if use_index and count else (np.zeros(0, dtype=np.int32) if use_index else None)
        # sabotage, sometimes: exactly one thing is wrong and the call must be refused with the code the header names
        want_rc = 0
        spc_arg, stride_arg, first_arg, count_arg, zero_calls = spc, stride, first, count, False
        u = rng.random()
        if u < 0.05:
            spc_arg, want_rc = int(rng.choice([0, 81, 159, 320])), E_BAD
        elif u < 0.09:
            stride_arg, want_rc = width - 1, E_BAD
        elif u < 0.13 and use_index and count >= 2:
            idx_arr[count - 1] = idx_arr[0]  # a repeated instance
            want_rc = E_BAD
        elif u < 0.16 and use_index and count >= 1:
            idx_arr[int(rng.integers(0, count))] = int(rng.choice([-1, N, N + 5]))
            want_rc = E_BAD
        elif u < 0.19 and not use_index:
            first_arg, count_arg = (-1, 1) if rng.random() < 0.5 else (N - 1, 2)
            want_rc = E_BAD
        elif u < 0.21:
            zero_calls, want_rc = True, E_BAD
        ptr = lambda a: None if a is None else (a.ctypes.data if a.size else None)
        dummy = 1  # a non-NULL pointer for planes of a zero-row call (never dereferenced)
        a = api.StreamArgs(ptr(far) if count else (dummy if has_far else None), ptr(near) if count else (dummy if has_near else None), ptr(clean),
                           ptr(out) if count else (dummy if has_near else None), stride_arg, 0 if use_index else first_arg, count_arg, spc_arg,
                           0 if zero_calls else n_calls, ptr(ms), ms_uniform, ptr(mask), ptr(status),
                           (idx_arr.ctypes.data if count else dummy) if use_index else None)
        slot = int(rng.integers(0, n_slots))
        free_slots = [k for k in range(n_slots) if k not in in_flight]
        if free_slots and rng.random() < 0.7:
            slot = int(rng.choice(free_slots))
        if sync:
            if 0 in in_flight:  # ProcessStream waits for slot 0 first
                pass
            rc = L.AecmB200_ProcessStream(b._h, C.byref(a))
            if 0 in in_flight:
                finish(0)
        else:
            busy = slot in in_flight
            rc = L.AecmB200_SubmitStream(b._h, C.byref(a), slot)
            if busy:
                # a refused argument is reported before the busy check; otherwise the slot is busy
                assert rc == (want_rc or E_BAD), (rc, want_rc)
                stats["busy"] += 1
                pl.release()
                return
        assert rc == want_rc, (rc, want_rc, dict(spc=spc_arg, stride=stride_arg, count=count, use_index=use_index, u=u, sync=sync,
                                                idx=None if idx_arr is None else idx_arr.tolist(), first=first_arg, count_arg=count_arg))
        if rc:
            stats["refused"] += 1
            pl.release()
            return
        stats["ok"] += 1
        # the contract, row by row
        want_out = np.zeros((count, stride), dtype=np.int16)
        want_st = np.zeros(count, dtype=np.int32)
        cmp_rows = []
        for r, i in enumerate(rows):
            st = 0
            o = model[i]
            for c in range(n_calls):
                sl = slice(c * spc, (c + 1) * spc)
                if has_far and (mask is None or mask[c, r]):
                    rcf = o.lib.oracle_aecm_buffer_farend(o.h, far[r, sl].ctypes.data_as(C.c_void_p), spc)
                    st = st or rcf
                if has_near:
                    cl = np.ascontiguousarray(clean[r, sl]) if clean is not None else None
                    rcp, w = o.process(np.ascontiguousarray(near[r, sl]), cl, int(ms[r]) if ms is not None else ms_uniform)
                    st = st or rcp
                    if rcp in (0, W_BAD):
                        want_out[r, sl] = w
                        cmp_rows.append((r, sl))
            want_st[r] = st
            stats["rows"] += 1
            stats["warn"] += st == W_BAD
            stats["uninit"] += st == E_UNINIT
        rec = (pl, out, want_out, cmp_rows, status, want_st)
        if sync:
            in_flight[-1] = rec
            finish(-1)
        else:
            in_flight[slot] = rec
            stats["max_in_flight"] = max(stats["max_in_flight"], len(in_flight))

    for step in range(220):
        u = rng.random()
        if u < 0.10:
            first, count = int(rng.integers(0, N)), int(rng.integers(0, N + 1))
            fs = int(rng.choice([8000, 16000, 16000, 44100]))
            rc = b.init(fs, first=first, count=count)
            ok_range = first + count <= N
            if fs == 44100:
                assert rc == E_BAD
            elif not ok_range:
                assert rc == E_BAD
            else:
                assert rc == 0
                for i in range(first, first + count):
                    assert model[i].lib.oracle_aecm_init(model[i].h, fs) == 0
        elif u < 0.16:
            first = int(rng.integers(0, N))
            count = int(rng.integers(0, N - first + 1))
            cng, mode = int(rng.choice([0, 1, 1, 2])), int(rng.choice([0, 1, 2, 3, 4, 5]))
            rc, st = b.set_config(cng, mode, first=first, count=count)
            assert rc == 0
            want = [model[i].lib.oracle_aecm_set_config(model[i].h, cng, mode) for i in range(first, first + count)]
            assert st.tolist() == want
        elif u < 0.19:
            first = int(rng.integers(0, N))
            count = int(rng.integers(0, N - first + 1))
            b.reset(first=first, count=count)  # raises on a non-zero return code
            for i in range(first, first + count):
                fresh(i)
        elif u < 0.24:
            i = int(rng.integers(0, N))
            path = rng.integers(0, 1 << 14, 65).astype(np.int16)
            assert b.init_echo_path(i, path) == model[i].lib.oracle_aecm_init_echo_path(model[i].h, path.ctypes.data_as(C.c_void_p), 130)
        elif u < 0.30:
            i = int(rng.integers(0, N))
            rc, got = b.get_echo_path(i)
            want = np.zeros(65, dtype=np.int16)
            assert rc == model[i].lib.oracle_aecm_get_echo_path(model[i].h, want.ctypes.data_as(C.c_void_p), 130)
            if rc == 0:
                np.testing.assert_array_equal(got, want)
        elif u < 0.40:
            slot = int(rng.integers(0, n_slots))
            assert L.AecmB200_WaitStream(b._h, slot) == 0
            if slot in in_flight:
                finish(slot)
        elif u < 0.70:
            stream_call(sync=False)
        else:
            stream_call(sync=True)
    for slot in sorted(in_flight, reverse=bool(seed % 2)):
        assert L.AecmB200_WaitStream(b._h, slot) == 0
        finish(slot)
    b.close()
    for o in model:
        o.close()
    print("fuzz stats", seed, stats)
    assert stats["ok"] >= 60 and stats["refused"] >= 5 and stats["rows"] >= 150 and stats["max_in_flight"] >= 3, stats
