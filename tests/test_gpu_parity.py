"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, must be
bit-exact with the CPU oracle, with the committed golden fixtures generated from the unmodified reference,
and with the reference library itself when oracle/_ref travelled along."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from cases import CASE_BY_NAME, CASES, far_mask, make_inputs, run_kwargs
from webrtc_aecm_b200.shard import stream_checksum
from webrtc_aecm_b200.synth import synth_streams

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STREAMS = json.load(open(os.path.join(GOLD, "streams.json")))["cases"]
VEC = np.load(os.path.join(GOLD, "vectors.npz"))


def gpu_run(case, far, near, clean, ms, n_calls_per_launch=None):
    """Drive the batch like the reference's caller (main.cc:111-144). Returns (out, status_or)."""
    from webrtc_aecm_b200 import AecmBatch

    n, total = far.shape
    spc = case["spc"]
    calls = total // spc
    b = AecmBatch(n, 0)
    assert b.init(case["fs"]) == 0
    rc, st = b.set_config(case["cng"], case["echo_mode"])
    assert rc == 0 and not st.any()
    mask = far_mask(case, calls) if case["skip_far_every"] else None
    out = np.zeros_like(near)
    acc = np.zeros(n, dtype=np.int32)
    if (ms is not None and np.ndim(ms) == 2) or case["extra_far_every"]:
        # msInSndCardBuf changes from call to call and/or BufferFarend runs twice before a Process: one launch per API call
        k = case["extra_far_every"]
        for c in range(calls):
            sl = slice(c * spc, (c + 1) * spc)
            f = np.ascontiguousarray(far[:, sl])
            for _ in range(0 if (mask is not None and not mask[c, 0]) else (2 if k and c % k == k - 1 else 1)):
                rc, status = b.buffer_farend(f)
                assert rc == 0
                acc |= status
            rc, o, status = b.process(np.ascontiguousarray(near[:, sl]), None if clean is None else np.ascontiguousarray(clean[:, sl]),
                                      ms=None if ms is None else (ms[:, c] if np.ndim(ms) == 2 else ms), ms_uniform=40)
            assert rc == 0
            out[:, sl] = o
            acc |= status
        b.close()
        return out, acc
    step = n_calls_per_launch or calls
    for c0 in range(0, calls, step):
        c1 = min(calls, c0 + step)
        sl = slice(c0 * spc, c1 * spc)
        rc, o, status = b.process_stream(np.ascontiguousarray(far[:, sl]), np.ascontiguousarray(near[:, sl]),
                                         None if clean is None else np.ascontiguousarray(clean[:, sl]),
                                         samples_per_call=spc, ms=ms, ms_uniform=40,
                                         far_mask=None if mask is None else np.ascontiguousarray(mask[c0:c1]))
        assert rc == 0
        out[:, sl] = o
        acc |= status
    b.close()
    return out, acc


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_matches_oracle_and_golden(case):
    far, near, clean, ms = make_inputs(case)
    out, acc = gpu_run(case, far, near, clean, ms)
    want, ret, _ = ol.run_streams(far, near, impl="oracle", clean=clean, ms=ms, **run_kwargs(case))
    bad = np.nonzero((out != want).any(axis=1))[0]
    assert bad.size == 0, f"instances {bad[:8]} differ, first sample {np.argmax(out[bad[0]] != want[bad[0]])}"
    # the oracle runner ORs all return codes; the device reports the first non-zero one -> compare non-zero-ness
    # and the warning value
    assert ((acc != 0) == (ret != 0)).all()
    assert set(acc[acc != 0].tolist()) <= {12100}
    g = STREAMS[case["name"]]
    assert stream_checksum(out) == g["sha256"]  # == the unmodified reference's output (tools/make_golden.py)
    key = "out__" + case["name"]
    if key in VEC.files:
        np.testing.assert_array_equal(out, VEC[key])


@pytest.mark.skipif(ol.ref_path() is None, reason="oracle/_ref/libaecm_ref.so did not travel")
@pytest.mark.parametrize("name", ["c2_16k_default_10s", "8k_stress", "16k_clean_path", "16k_ms_mix"])
def test_cuda_matches_reference_library(name):
    case = CASE_BY_NAME[name]
    far, near, clean, ms = make_inputs(case)
    out, _ = gpu_run(case, far, near, clean, ms)
    ref, _, _ = ol.run_streams(far, near, fs=case["fs"], samples_per_call=case["spc"], impl="ref", clean=clean, ms=ms,
                               cng_mode=case["cng"], echo_mode=case["echo_mode"])
    np.testing.assert_array_equal(out, ref)


@pytest.mark.parametrize("chunk", [1, 2, 7])
def test_fused_calls_equal_call_by_call(chunk):
    """nCalls > 1 (state held on-chip across calls) must not change a single bit."""
    case = CASE_BY_NAME["16k_far_underrun"]
    far, near, clean, ms = make_inputs(case)
    whole, _ = gpu_run(case, far, near, clean, ms)
    stepped, _ = gpu_run(case, far, near, clean, ms, n_calls_per_launch=chunk)
    np.testing.assert_array_equal(whole, stepped)


def test_state_matches_oracle_field_by_field():
    """Stage-level parity: after every call the device record equals the oracle's state."""
    from webrtc_aecm_b200 import AecmBatch

    fs, spc, n, calls = 16000, 160, 3, 120
    far, near = synth_streams(n, calls * spc, fs, seed=77, delay_ms=[0, 40, 100])
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    orcs = [ol.OracleInstance(fs) for _ in range(n)]
    for c in range(calls):
        sl = slice(c * spc, (c + 1) * spc)
        f, nr = np.ascontiguousarray(far[:, sl]), np.ascontiguousarray(near[:, sl])
        rc, out, status = b.process_stream(f, nr, samples_per_call=spc)
        assert rc == 0
        for i in range(n):
            orcs[i].buffer_farend(f[i])
            orc_rc, o = orcs[i].process(nr[i], None, 40)
            assert orc_rc == status[i]
            np.testing.assert_array_equal(o, out[i])
            if c % 10 == 9 or c < 12:
                assert ol.compare_states(b.read_state(i), orcs[i].state()) == []
    b.close()


def test_full_batch_4096_properties():
    """BASELINE configs[1] size, run far enough (460 calls = 1135 blocks, fused 46 per launch) that every instance of the
    FULL batch is in steady state (startupState == 2, channel store/reset decisions running).  64 distinct streams are
    scattered over the 4096 instances: every instance must equal the oracle -- and the unmodified reference library when
    it travelled along -- on its stream, bit for bit, wherever it sits in the batch; and every instance differs from
    pass-through once the canceller is active."""
    from webrtc_aecm_b200 import AecmBatch

    fs, spc, n, calls, uniq, fuse = 16000, 160, 4096, 460, 64, 46
    far_u, near_u = synth_streams(uniq, calls * spc, fs, seed=4096, delay_ms=np.arange(uniq) % 100)
    perm = (np.arange(n) * 37) % uniq  # scatter the 64 distinct streams over the batch
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    out = np.zeros((n, calls * spc), dtype=np.int16)
    for c0 in range(0, calls, fuse):
        sl = slice(c0 * spc, (c0 + fuse) * spc)
        rc, o, status = b.process_stream(np.ascontiguousarray(far_u[perm, sl]), np.ascontiguousarray(near_u[perm, sl]),
                                         samples_per_call=spc)
        assert rc == 0 and not status.any()
        out[:, sl] = o
    for k in (0, n // 2, n - 1):
        st = b.read_state(k).s
        assert st.startup_state == 2 and st.tot_count == 1135  # 460 calls, the first 6 of them start-up pass-through: 454 x 2.5 blocks
    want, _, _ = ol.run_streams(far_u, near_u, fs=fs, impl="oracle")
    np.testing.assert_array_equal(out, want[perm])
    if ol.ref_path():
        ref, _, _ = ol.run_streams(far_u, near_u, fs=fs, impl="ref")
        np.testing.assert_array_equal(out[:uniq * 37:37], ref[perm[:uniq * 37:37]])  # 64 sampled instances, all 64 streams
        np.testing.assert_array_equal(want, ref)
    assert (out[:, 8 * spc:] != near_u[perm][:, 8 * spc:]).any(axis=1).all()
    b.close()


def test_c4_262144_instances_on_one_gpu():
    """BASELINE configs[3] at its stated size on one GPU (the persistent kernel walks 64 rows per warp): 262144 instances,
    128 distinct streams dealt over them; EVERY instance must reproduce the oracle's output for its stream."""
    from webrtc_aecm_b200 import AecmBatch

    fs, spc, n, calls, uniq = 16000, 160, 262144, 24, 128
    far_u, near_u = synth_streams(uniq, calls * spc, fs, seed=262144, delay_ms=np.arange(uniq) % 110)
    idx = (np.arange(n) * 61) % uniq
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    rc, out, status = b.process_stream(np.ascontiguousarray(far_u[idx]), np.ascontiguousarray(near_u[idx]), samples_per_call=spc)
    assert rc == 0 and not status.any()
    want, _, _ = ol.run_streams(far_u, near_u, fs=fs, impl="oracle")
    bad = np.nonzero((out != want[idx]).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} instances differ, first {bad[:8]}"
    b.close()


@pytest.mark.parametrize("fs", [16000, 8000])
def test_two_minute_soak_matches_oracle(fs):
    """120 s of audio per instance (12 000 API calls, fused 100 per launch) against the oracle, output and final state."""
    from cases import soak_inputs
    from webrtc_aecm_b200 import AecmBatch

    far, near = soak_inputs(fs)
    n, total = far.shape
    spc = fs // 100
    ms = ((np.arange(n) * 37) % 300).astype(np.int16)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle", ms=ms, echo_mode=3)
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    out = np.zeros_like(near)
    step = 100 * spc
    for t0 in range(0, total, step):
        sl = slice(t0, min(total, t0 + step))
        rc, o, status = b.process_stream(np.ascontiguousarray(far[:, sl]), np.ascontiguousarray(near[:, sl]),
                                         samples_per_call=spc, ms=ms)
        assert rc == 0 and not status.any()
        out[:, sl] = o
    bad = np.nonzero((out != want).any(axis=1))[0]
    assert bad.size == 0, f"instances {bad[:8]} differ, first sample {np.argmax(out[bad[0]] != want[bad[0]])}"
    # final state of one instance, field by field
    o = ol.OracleInstance(fs, 1, 3)
    for c in range(total // spc):
        o.buffer_farend(far[3, c * spc:(c + 1) * spc])
        o.process(near[3, c * spc:(c + 1) * spc], None, int(ms[3]))
    assert ol.compare_states(b.read_state(3), o.state()) == []
    o.close()
    b.close()


def test_delay_sweep_32768_instances_sampled():
    """BASELINE configs[4] size: per-instance far-end delay 0-120 ms across 32768 instances; every 257th
    instance is checked against the oracle, all others through stream-duplication equality."""
    from webrtc_aecm_b200 import AecmBatch

    fs, spc, n, calls, uniq = 16000, 160, 32768, 40, 121
    far_u, near_u = synth_streams(uniq, calls * spc, fs, seed=32768, delay_ms=np.arange(uniq))
    idx = np.arange(n) % uniq
    far, near = np.ascontiguousarray(far_u[idx]), np.ascontiguousarray(near_u[idx])
    ms = (np.arange(n) % 3 * 40).astype(np.int16)  # 0 / 40 / 80 ms sound-card delay mixes wrapper trajectories
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    rc, out, status = b.process_stream(far, near, samples_per_call=spc, ms=ms)
    assert rc == 0 and not status.any()
    pick = np.arange(0, n, 257)
    want, _, _ = ol.run_streams(far[pick], near[pick], fs=fs, impl="oracle", ms=ms[pick])
    np.testing.assert_array_equal(out[pick], want)
    # instances with the same (stream, ms) pair must agree bit for bit: period lcm(121, 3) = 363
    np.testing.assert_array_equal(out[: n - 363], out[363:])
    b.close()


def test_65536_instances_8khz_consistency():
    """BASELINE configs[2] size (8 kHz): duplicated-stream equality across the whole batch plus an oracle sample."""
    from webrtc_aecm_b200 import AecmBatch

    fs, spc, n, calls, uniq = 8000, 80, 65536, 30, 32
    far_u, near_u = synth_streams(uniq, calls * spc, fs, seed=65536, delay_ms=40)
    far, near = np.ascontiguousarray(np.tile(far_u, (n // uniq, 1))), np.ascontiguousarray(np.tile(near_u, (n // uniq, 1)))
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    rc, out, status = b.process_stream(far, near, samples_per_call=spc)
    assert rc == 0 and not status.any()
    want, _, _ = ol.run_streams(far_u, near_u, fs=fs, impl="oracle")
    np.testing.assert_array_equal(out[:uniq], want)
    assert (out.reshape(n // uniq, uniq, -1) == out[None, :uniq]).all()
    b.close()


def test_ragged_ranges_and_partial_batches():
    """first/count sub-ranges: processing only a slice must leave the other instances untouched."""
    from webrtc_aecm_b200 import AecmBatch

    fs, spc, n, calls = 16000, 160, 10, 30
    far, near = synth_streams(n, calls * spc, fs, seed=5)
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    before = bytes(b.read_state(9))
    rc, out, status = b.process_stream(np.ascontiguousarray(far[3:8]), np.ascontiguousarray(near[3:8]),
                                       samples_per_call=spc, first=3, count=5)
    assert rc == 0 and not status.any()
    assert bytes(b.read_state(9)) == before
    want, _, _ = ol.run_streams(far[3:8], near[3:8], fs=fs, impl="oracle")
    np.testing.assert_array_equal(out, want)
    # count == 0 is a no-op; out-of-range is rejected
    assert b.process_stream(far[:0], near[:0], samples_per_call=spc, first=0, count=0)[0] == 0
    assert b.process_stream(far, near, samples_per_call=spc, first=5, count=10)[0] == 12004
    b.close()


def test_device_floor_sqrt_is_exact_on_all_2_to_31_arguments():
    """The magnitude step uses one hardware square-root estimate plus two integer corrections; swept exhaustively on
    the device against the reference's bit-by-bit construction (signal_processing_library.cc:76-105)."""
    from webrtc_aecm_b200 import api

    L = api.lib()
    assert L.AecmB200_SelfTestSqrtFloor(0, 1 << 30) == 0
    assert L.AecmB200_SelfTestSqrtFloor(1 << 30, 1 << 30) == 0


def test_device_fft_known_answers():
    """The kernel's own warp FFT (AecmB200_SelfTestFft*: the register-resident forward128 / inverse128 with the production
    tables, PRMT packing and redux peaks) on the GPU against (a) the committed golden vectors produced by the reference's
    WebRtcSpl_RealForwardFFT / WebRtcSpl_RealInverseFFT, returned scale included, and (b) a larger adversarial set
    (impulses, full scale, per-stage peaks straddling 13573 / 27146) with answers from the same reference functions
    (tests/golden/fft_adversarial.npz, written by tools/make_golden.py)."""
    import ctypes as C

    from webrtc_aecm_b200 import api

    L = api.lib()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    fin, fout = np.ascontiguousarray(VEC["fft_fwd_in"]), VEC["fft_fwd_out"]
    iin, iout, iscale = np.ascontiguousarray(VEC["fft_inv_in"]), VEC["fft_inv_out"], VEC["fft_inv_scale"]
    for two in (0, 1):
        got = np.zeros_like(fout)
        assert L.AecmB200_SelfTestFftForward(vp(fin), vp(got), len(fin), two) == 0
        np.testing.assert_array_equal(got, fout)
    got, sc = np.zeros_like(iout), np.zeros(len(iin), dtype=np.int32)
    assert L.AecmB200_SelfTestFftInverse(vp(iin), vp(got), vp(sc), len(iin)) == 0
    np.testing.assert_array_equal(sc, iscale)
    np.testing.assert_array_equal(got, iout)

    A = np.load(os.path.join(GOLD, "fft_adversarial.npz"))  # reference answers for tests/cases.py:fft_adversarial_vectors
    fwd, inv = np.ascontiguousarray(A["fwd_in"]), np.ascontiguousarray(A["inv_in"])
    for two in (0, 1):
        got = np.zeros_like(A["fwd_out"])
        assert L.AecmB200_SelfTestFftForward(vp(fwd), vp(got), len(fwd), two) == 0
        np.testing.assert_array_equal(got, A["fwd_out"])
    got, sc = np.zeros_like(A["inv_out"]), np.zeros(len(inv), dtype=np.int32)
    assert L.AecmB200_SelfTestFftInverse(vp(inv), vp(got), vp(sc), len(inv)) == 0
    np.testing.assert_array_equal(sc, A["inv_scale"])
    np.testing.assert_array_equal(got, A["inv_out"])


@pytest.mark.parametrize("adjacent,all_slots", [(False, False), (True, False), (True, True)],
                         ids=["separate_planes", "near_follows_far", "all_slots_in_flight"])
def test_pipelined_submit_wait_equals_synchronous(adjacent, all_slots):
    """AecmB200_SubmitStream / WaitStream (copies of later calls overlap the kernels of earlier ones) must give exactly
    the outputs of the synchronous call sequence, with two calls in flight or with every slot in use.  With `adjacent`
    every call's near plane starts where its far plane ends, which the library turns into a single host-to-device
    transfer."""
    import ctypes as C

    from webrtc_aecm_b200 import AecmBatch, api

    fs, spc, n, calls = 16000, 160, 256, 50
    far, near = synth_streams(n, calls * spc, fs, seed=909, delay_ms=np.arange(n) % 90)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle")
    # [call][instance][samples] so every call's plane is dense
    hfar = np.ascontiguousarray(far.reshape(n, calls, spc).transpose(1, 0, 2))
    hnear = np.ascontiguousarray(near.reshape(n, calls, spc).transpose(1, 0, 2))
    if adjacent:
        hin = np.ascontiguousarray(np.stack([hfar, hnear], axis=1))  # [call][far|near][instance][samples]
        hfar, hnear = hin[:, 0], hin[:, 1]
        assert hnear[3].ctypes.data == hfar[3].ctypes.data + hfar[3].nbytes
    hout = np.zeros_like(near).reshape(calls, n, spc)
    L = api.lib()
    n_slots = L.AecmB200_NumStreamSlots()
    assert n_slots >= 2
    S = n_slots if all_slots else 2
    status = [np.zeros(n, dtype=np.int32) for _ in range(S)]
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0

    def submit(k):
        a = api.StreamArgs(hfar[k].ctypes.data, hnear[k].ctypes.data, None, hout[k].ctypes.data, spc, 0, n, spc, 1, None, 40,
                           None, status[k % S].ctypes.data)
        assert L.AecmB200_SubmitStream(b._h, C.byref(a), k % S) == 0

    def wait(k):
        assert L.AecmB200_WaitStream(b._h, k % S) == 0 and not status[k % S].any()

    submit(0)
    assert L.AecmB200_SubmitStream(b._h, None, 1) == 12003  # NULL args
    a_busy = api.StreamArgs(hfar[0].ctypes.data, hnear[0].ctypes.data, None, hout[0].ctypes.data, spc, 0, n, spc, 1, None, 40, None, None)
    assert L.AecmB200_SubmitStream(b._h, C.byref(a_busy), 0) == 12004  # slot 0 is still in flight
    assert L.AecmB200_SubmitStream(b._h, C.byref(a_busy), n_slots) == 12004  # no such slot
    for k in range(1, calls):
        submit(k)
        if k - (S - 1) >= 0:
            wait(k - (S - 1))
    for k in range(calls - (S - 1), calls):
        wait(k)
    got = hout.transpose(1, 0, 2).reshape(n, calls * spc)
    np.testing.assert_array_equal(got, want)
    b.close()


def test_device_planes_at_odd_sample_offsets_and_strides():
    """AecmB200_ProcessStreamDevice with rows that start at odd int16 positions (odd row stride, offset 1): the kernel's
    32-bit sample movers must fall back to 16-bit accesses and still produce the oracle's stream."""
    import torch

    from webrtc_aecm_b200 import AecmBatch

    fs, spc, n, calls = 16000, 160, 9, 40
    far, near = synth_streams(n, calls * spc, fs, seed=4321, delay_ms=np.arange(n) * 9)
    clean = np.ascontiguousarray(near[:, ::-1] // 2)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle", clean=clean)
    stride = calls * spc + 3  # odd: every second row starts on an odd sample
    dev = torch.device("cuda", 0)

    def plane(a):
        t = torch.zeros((n, stride), dtype=torch.int16, device=dev)
        t[:, 1:1 + a.shape[1]] = torch.from_numpy(a).to(dev)
        return t

    d_far, d_near, d_clean, d_out = plane(far), plane(near), plane(clean), plane(np.zeros_like(near))
    b = AecmBatch(n, 0)
    b.adopt_cuda_stream(torch.cuda.current_stream().cuda_stream)
    assert b.init(fs) == 0
    for c0 in range(0, calls, 8):
        rc = b.process_stream_device(d_far, d_near, d_clean, d_out, row_stride=stride, samples_per_call=spc, n_calls=min(8, calls - c0),
                                     sample_offset=1 + c0 * spc)
        assert rc == 0
    b.synchronize()
    got = d_out[:, 1:1 + calls * spc].cpu().numpy()
    np.testing.assert_array_equal(got, want)
    b.close()


def test_instance_index_scatters_rows_over_the_batch():
    """AecmB200StreamArgs.instanceIndex: plane row r works on instance instanceIndex[r].  A permuted subset of a batch,
    driven through host planes, must equal the oracle per stream while the untouched instances stay untouched."""
    import ctypes as C

    from webrtc_aecm_b200 import AecmBatch, api

    fs, spc, n, calls, total = 16000, 160, 40, 60, 100
    far, near = synth_streams(n, calls * spc, fs, seed=1234, delay_ms=np.arange(n) * 2)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle")
    rng = np.random.default_rng(3)
    index = np.ascontiguousarray(rng.permutation(total)[:n].astype(np.int32))
    b = AecmBatch(total, 0)
    assert b.init(fs) == 0
    L = api.lib()
    out = np.zeros_like(near)
    status = np.ones(n, dtype=np.int32)
    for k0 in range(0, calls, 7):  # several launches, 7 (then fewer) fused calls each
        k1 = min(calls, k0 + 7)
        a = api.StreamArgs(far[:, k0 * spc:].ctypes.data, near[:, k0 * spc:].ctypes.data, None, out[:, k0 * spc:].ctypes.data,
                           calls * spc, 0, n, spc, k1 - k0, None, 40, None, status.ctypes.data, index.ctypes.data)
        assert L.AecmB200_ProcessStream(b._h, C.byref(a)) == 0 and not status.any()
    np.testing.assert_array_equal(out, want)
    touched = set(index.tolist())
    for inst in range(total):
        tot = b.read_state(inst).s.tot_count
        assert (tot > 0) == (inst in touched), inst
    bad = index.copy()
    bad[5] = total  # out of range
    a = api.StreamArgs(far.ctypes.data, near.ctypes.data, None, out.ctypes.data, calls * spc, 0, n, spc, 1, None, 40, None, None, bad.ctypes.data)
    assert L.AecmB200_ProcessStream(b._h, C.byref(a)) == 12004
    dup = index.copy()
    dup[7] = dup[2]  # two rows naming one instance would be two warps racing on one record: rejected up front
    a = api.StreamArgs(far.ctypes.data, near.ctypes.data, None, out.ctypes.data, calls * spc, 0, n, spc, 1, None, 40, None, None, dup.ctypes.data)
    assert L.AecmB200_SubmitStream(b._h, C.byref(a), 1) == 12004
    assert L.AecmB200_WaitStream(b._h, 1) == 0  # the refused submission left its slot free ...
    a = api.StreamArgs(far.ctypes.data, near.ctypes.data, None, out.ctypes.data, calls * spc, 0, n, spc, 1, None, 40, None,
                       status.ctypes.data, index.ctypes.data)
    assert L.AecmB200_SubmitStream(b._h, C.byref(a), 1) == 0  # ... so it takes the next one
    assert L.AecmB200_SubmitStream(b._h, C.byref(a), 1) == 12004  # and is busy until waited for
    assert L.AecmB200_WaitStream(b._h, 1) == 0 and not status.any()
    b.close()


def test_group_on_distinct_devices_and_callers_device_is_left_alone():
    """AecmB200_Group* with one shard per visible GPU (needs >= 2): every shard's batch sits on its own device, the sharded
    result equals the oracle, and no library call leaves the calling thread on another CUDA device than it found."""
    import ctypes as C

    import torch

    from webrtc_aecm_b200 import api

    L = api.lib()
    ndev = L.AecmB200_DeviceCount()
    if ndev < 2:
        pytest.skip("one GPU visible: distinct-device sharding needs two")
    fs, spc, calls = 16000, 160, 50
    n = 64 * ndev + 3
    far, near = synth_streams(n, calls * spc, fs, seed=8, delay_ms=np.arange(n) % 90)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle")
    home = ndev - 1
    torch.cuda.set_device(home)
    g = L.AecmB200_GroupCreate(n, ndev, None)
    assert g and torch.cuda.current_device() == home
    f0, c0 = C.c_int32(), C.c_int32()
    devices = []
    for i in range(ndev):
        h = L.AecmB200_GroupShard(g, i, C.byref(f0), C.byref(c0))
        devices.append(L.AecmB200_Device(h))
    assert devices == list(range(ndev))
    assert L.AecmB200_GroupInit(g, fs) == 0 and torch.cuda.current_device() == home
    st = np.zeros(n, dtype=np.int32)
    assert L.AecmB200_GroupSetConfig(g, api.AecmConfig(1, 3), st.ctypes.data) == 0 and not st.any()
    out = np.zeros_like(near)
    S = L.AecmB200_NumStreamSlots()
    fuse = 5
    chunks = [(np.ascontiguousarray(far[:, c * spc:(c + fuse) * spc]), np.ascontiguousarray(near[:, c * spc:(c + fuse) * spc]),
               np.zeros((n, fuse * spc), dtype=np.int16)) for c in range(0, calls, fuse)]
    for k, (f, nr, o) in enumerate(chunks):  # pipelined over the slots, every shard working on its slice
        if k >= S:
            assert L.AecmB200_GroupWaitStream(g, k % S) == 0
        a = api.StreamArgs(f.ctypes.data, nr.ctypes.data, None, o.ctypes.data, fuse * spc, 0, n, spc, fuse, None, 40, None, st.ctypes.data)
        assert L.AecmB200_GroupSubmitStream(g, C.byref(a), k % S) == 0
        assert torch.cuda.current_device() == home
    for s_ in range(S):
        assert L.AecmB200_GroupWaitStream(g, s_) == 0
    assert not st.any()
    for k, (_, _, o) in enumerate(chunks):
        out[:, k * fuse * spc:(k + 1) * fuse * spc] = o
    np.testing.assert_array_equal(out, want)
    L.AecmB200_GroupFree(g)
    assert torch.cuda.current_device() == home
    torch.cuda.set_device(0)


def test_group_sharding_over_devices_matches_oracle():
    """AecmB200_Group*: the instance range is ceil-split into shards (here: three shards, all placed on GPU 0 unless
    more GPUs are visible) and every shard processes its slice of the host planes; ragged sub-ranges included."""
    import ctypes as C

    from webrtc_aecm_b200 import api

    L = api.lib()
    ndev = max(1, L.AecmB200_DeviceCount())
    fs, spc, n, calls, shards = 16000, 160, 11, 40, 3
    far, near = synth_streams(n, calls * spc, fs, seed=1212, delay_ms=40)
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle")
    devs = (C.c_int32 * shards)(*[i % ndev for i in range(shards)])
    g = L.AecmB200_GroupCreate(n, shards, devs)
    assert g and L.AecmB200_GroupNumShards(g) == shards
    f0, c0 = C.c_int32(), C.c_int32()
    sizes = []
    for i in range(shards):
        assert L.AecmB200_GroupShard(g, i, C.byref(f0), C.byref(c0))
        sizes.append((f0.value, c0.value))
    assert sizes == [(0, 4), (4, 4), (8, 3)]
    assert L.AecmB200_GroupInit(g, fs) == 0
    st = np.zeros(n, dtype=np.int32)
    assert L.AecmB200_GroupSetConfig(g, api.AecmConfig(1, 3), st.ctypes.data) == 0 and not st.any()
    out = np.zeros_like(near)
    half = calls // 2 * spc
    for sl in (slice(0, half), slice(half, calls * spc)):  # two group calls, fused 20 x 10 ms each
        f, nr = np.ascontiguousarray(far[:, sl]), np.ascontiguousarray(near[:, sl])
        o = np.zeros_like(nr)
        a = api.StreamArgs(f.ctypes.data, nr.ctypes.data, None, o.ctypes.data, f.shape[1], 0, n, spc, f.shape[1] // spc, None,
                           40, None, st.ctypes.data)
        assert L.AecmB200_GroupProcessStream(g, C.byref(a)) == 0 and not st.any()
        out[:, sl] = o
    np.testing.assert_array_equal(out, want)
    # a ragged sub-range that straddles two shards: instances 2..9 only
    a = api.StreamArgs(f.ctypes.data, nr.ctypes.data, None, o.ctypes.data, f.shape[1], 2, 8, spc, 1, None, 40, None, None)
    assert L.AecmB200_GroupProcessStream(g, C.byref(a)) == 0
    assert L.AecmB200_GroupProcessStream(g, None) == 12003
    # the pipelined form: fresh state, one 10 ms call per submission, two slots in flight on every shard
    assert L.AecmB200_GroupInit(g, fs) == 0
    assert L.AecmB200_GroupSetConfig(g, api.AecmConfig(1, 3), st.ctypes.data) == 0
    hf = np.ascontiguousarray(far.reshape(n, calls, spc).transpose(1, 0, 2))
    hn = np.ascontiguousarray(near.reshape(n, calls, spc).transpose(1, 0, 2))
    ho = np.zeros_like(hn)
    for k in range(calls):
        a = api.StreamArgs(hf[k].ctypes.data, hn[k].ctypes.data, None, ho[k].ctypes.data, spc, 0, n, spc, 1, None, 40, None, None)
        assert L.AecmB200_GroupSubmitStream(g, C.byref(a), k % 2) == 0
        if k:
            assert L.AecmB200_GroupWaitStream(g, (k - 1) % 2) == 0
    assert L.AecmB200_GroupWaitStream(g, (calls - 1) % 2) == 0
    np.testing.assert_array_equal(ho.transpose(1, 0, 2).reshape(n, calls * spc), want)
    L.AecmB200_GroupFree(g)


DLY_GOLD = json.load(open(os.path.join(GOLD, "dly_streams.json")))["cases"]


@pytest.mark.parametrize("cfg", __import__("cases").DLY_CONFIGS, ids=[c["id"] for c in __import__("cases").DLY_CONFIGS])
def test_delay_estimator_options_match_reference(cfg):
    """SURVEY 8-f4 through the C ABI (AecmB200_SetDelayEstimatorOptions / _SoftResetDelayEstimator / _GetDelayEstimate): robust
    validation with its float histogram, near-end lookahead and soft resets on the delay-sweep inputs.  Output AND the
    delay estimate after every call must equal the oracle's, the checksums of the reference driven through its own
    setters (tests/golden/dly_streams.json), and that reference itself when oracle/_ref/libaecm_ref_shim.so travelled."""
    from webrtc_aecm_b200 import AecmBatch

    from cases import dly_inputs

    fs, spc, far, near, ms, kw = dly_inputs(cfg)
    n, total = far.shape
    calls = total // spc
    want, ret, _, wtrace = ol.run_streams(far, near, impl="oracle", ms=ms, dly=cfg["dly"], soft_reset=cfg["soft_reset"], delay_trace=True, **kw)
    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    rc, st = b.set_config(kw["cng_mode"], kw["echo_mode"])
    assert rc == 0 and not st.any()
    if cfg["dly"] is not None:
        rc, st = b.set_delay_estimator_options(**cfg["dly"])
        assert rc == 0 and not st.any()
    every, shift = cfg["soft_reset"] or (0, 0)
    out = np.zeros_like(near)
    trace = np.zeros((n, calls), dtype=np.int32)
    for c in range(calls):  # one launch per call: the delay estimate is read back after each
        if every and c % every == every - 1:
            rc, applied = b.soft_reset_delay_estimator(-shift if (c // every) & 1 else shift)
            assert rc == 0
        sl = slice(c * spc, (c + 1) * spc)
        rc, o, status = b.process_stream(np.ascontiguousarray(far[:, sl]), np.ascontiguousarray(near[:, sl]), samples_per_call=spc,
                                         ms=ms, ms_uniform=40)
        assert rc == 0 and not status.any()
        out[:, sl] = o
        rc, d, q = b.delay_estimate()
        assert rc == 0
        trace[:, c] = d
    bad = np.nonzero((out != want).any(axis=1))[0]
    assert bad.size == 0, f"instances {bad[:8]} differ, first sample {np.argmax(out[bad[0]] != want[bad[0]])}"
    np.testing.assert_array_equal(trace, wtrace)
    gold = DLY_GOLD[cfg["id"]]
    assert stream_checksum(out) == gold["sha256"] and stream_checksum(trace) == gold["delay_trace_sha256"]
    assert np.isfinite(q).all() and (q >= 0).all() and (q <= 1).all()
    if every == 0:  # WebRtc_last_delay_quality (a float) of a few instances, from oracle instances stepped alongside
        for inst in sorted({0, n // 2, n - 1}):
            o = ol.OracleInstance(fs, kw["cng_mode"], kw["echo_mode"])
            if cfg["dly"] is not None:
                assert o.set_delay_estimator_options(**cfg["dly"]) == 0
            msi = 40 if ms is None else int(np.resize(ms, n)[inst])
            for c in range(calls):
                o.buffer_farend(far[inst, c * spc:(c + 1) * spc])
                o.process(near[inst, c * spc:(c + 1) * spc], None, msi)
            od, oq = o.delay_estimate()
            assert (od, oq) == (int(d[inst]), float(q[inst])), (inst, od, oq, d[inst], q[inst])
            o.close()
    if ol.ref_shim_path():
        rout, _, _, rtrace = ol.run_streams(far, near, impl="ref", ms=ms, dly=cfg["dly"], soft_reset=cfg["soft_reset"], delay_trace=True, **kw)
        np.testing.assert_array_equal(out, rout)
        np.testing.assert_array_equal(trace, rtrace)
    b.close()


def test_delay_estimator_options_fused_calls_error_codes_and_init_semantics():
    """(a) Robust validation with many calls fused per launch equals call by call; (b) argument checking follows the reference's
    setters; (c) the options belong to the estimator object: they survive AecmB200_Init, AecmB200_Reset clears them; (d) soft
    reset returns the lookahead shift actually applied (delay_estimator.cc:506-519)."""
    import ctypes as C

    from webrtc_aecm_b200 import AecmBatch, api

    case = CASE_BY_NAME["c5_delay_sweep_0_120ms"]
    far, near, clean, ms = make_inputs(case)
    dly = dict(robust=1, allowed_offset=2, max_lookahead=6, lookahead=2)
    want, _, _ = ol.run_streams(far, near, impl="oracle", ms=ms, dly=dly, **run_kwargs(case))
    n = far.shape[0]
    b = AecmBatch(n, 0)
    rc, st = b.set_delay_estimator_options(robust=1)
    assert rc == 0 and (st == api.AECM_UNINITIALIZED_ERROR).all()
    assert b.init(16000) == 0
    rc, st = b.set_delay_estimator_options(**dly)
    assert rc == 0 and not st.any()
    rc, out, status = b.process_stream(far, near, samples_per_call=160, ms=ms)  # 600 calls in one launch
    assert rc == 0 and not status.any()
    np.testing.assert_array_equal(out, want)
    # (c) Init keeps the options (lookahead as last set), so a second run from Init reproduces the same output ...
    assert b.init(16000) == 0
    rc, out2, status = b.process_stream(far, near, samples_per_call=160, ms=ms)
    np.testing.assert_array_equal(out2, want)
    # (b)
    L = api.lib()
    for bad in (api.DelayEstimatorOptions(2, 0, -1, -1), api.DelayEstimatorOptions(0, -1, -1, -1), api.DelayEstimatorOptions(0, 0, 16, -1),
                api.DelayEstimatorOptions(0, 0, 3, 4)):
        assert L.AecmB200_SetDelayEstimatorOptions(b._h, 0, n, C.byref(bad), None) == api.AECM_BAD_PARAMETER_ERROR
    assert L.AecmB200_SetDelayEstimatorOptions(b._h, 0, n, None, None) == api.AECM_NULL_POINTER_ERROR
    assert L.AecmB200_SoftResetDelayEstimator(b._h, 0, n, 100, None) == api.AECM_BAD_PARAMETER_ERROR
    rc, st = b.set_delay_estimator_options(robust=1, allowed_offset=2, lookahead=7)  # beyond max_lookahead 6: WebRtc_set_lookahead -> -1
    assert rc == 0 and (st == -1).all()
    # (d) lookahead 2, near history 7: +1 -> 1 (applied 1); +5 -> 0 (applied 1); -9 -> 6 (applied -6)
    for shift, applied in ((1, 1), (5, 1), (-9, -6)):
        rc, got = b.soft_reset_delay_estimator(shift)
        assert rc == 0 and (got == applied).all(), (shift, got)
    # ... and Reset returns to AECM's own wiring: the default output
    b.reset()
    assert b.init(16000) == 0
    base, _, _ = ol.run_streams(far, near, impl="oracle", ms=ms, **run_kwargs(case))
    rc, out3, status = b.process_stream(far, near, samples_per_call=160, ms=ms)
    np.testing.assert_array_equal(out3, base)
    b.close()


def test_far_and_near_planes_that_touch_but_are_two_pinned_allocations():
    """A dense near plane that starts exactly where the dense far plane ends normally crosses PCIe in ONE transfer.  When the
    two planes are separate pinned allocations that merely touch (two cudaMallocHost blocks can end up back to back, as
    examples/delay_estimator_options.c found out), the runtime refuses a copy spanning both: the library must then move
    them separately instead of failing the call."""
    import torch

    from webrtc_aecm_b200 import AecmBatch

    fs, spc, n, calls = 16000, 160, 64, 12
    plane = n * spc  # 20480 bytes = 5 pages
    far_s, near_s = synth_streams(n, calls * spc, fs, seed=321, delay_ms=30)
    want, _, _ = ol.run_streams(far_s, near_s, fs=fs, impl="oracle")
    raw = np.zeros(2 * plane + 4096, dtype=np.int16)
    off = (-raw.ctypes.data % 4096) // 2
    far, near = raw[off:off + plane].reshape(n, spc), raw[off + plane:off + 2 * plane].reshape(n, spc)
    assert near.ctypes.data == far.ctypes.data + plane * 2
    rt = torch.cuda.cudart()
    for a in (far, near):  # two registrations of two adjacent page ranges
        assert int(rt.cudaHostRegister(a.ctypes.data, plane * 2, 0)) == 0
    try:
        b = AecmBatch(n, 0)
        assert b.init(fs) == 0
        for c in range(calls):
            far[:] = far_s[:, c * spc:(c + 1) * spc]
            near[:] = near_s[:, c * spc:(c + 1) * spc]
            rc, out, status = b.process_stream(far, near, samples_per_call=spc)
            assert rc == 0 and not status.any(), (rc, b._lib.AecmB200_LastError(b._h))
            np.testing.assert_array_equal(out, want[:, c * spc:(c + 1) * spc])
        b.close()
    finally:
        for a in (far, near):
            rt.cudaHostUnregister(a.ctypes.data)


@pytest.mark.parametrize("seed,fs", [(3, 8000), (4, 16000), (7, 16000), (10, 16000), (12, 8000), (17, 16000)])
def test_random_call_sequences_match_oracle(seed, fs):
    """The randomised API scripts of tests/test_fuzz_api.py on the GPU: 12 instances share a script's call structure (mixed
    80/160-sample calls, BufferFarend 0/1/2 times per Process, clean plane or not, echo mode, CNG) and differ in audio and in
    their per-call msInSndCardBuf; every call is its own launch through AecmB200_BufferFarend / AecmB200_Process.  Output and
    per-call status equal the oracle's, which the same scripts pin to the reference."""
    import test_fuzz_api as fz
    from webrtc_aecm_b200 import AecmBatch

    n_inst, n_calls = 12, 400
    base = fz.make_script(seed, n_calls, fs)
    total = int(base["n"].sum())
    far, near = synth_streams(n_inst, total, fs, seed=seed * 31, delay_ms=(np.arange(n_inst) * 9) % 110)
    cln = synth_streams(n_inst, total, fs, seed=seed * 31 + 1, delay_ms=5)[1]
    ms = np.stack([fz.make_script(seed * 100 + i, n_calls, fs)["ms"] for i in range(n_inst)]).astype(np.int16)
    use_clean = bool(base["clean"].all())
    want = np.zeros_like(near)
    want_rc = np.zeros((n_inst, n_calls), dtype=np.int64)
    for i in range(n_inst):
        s = dict(base, far=far[i], near=near[i], cln=cln[i], ms=ms[i].astype(np.int32))
        want[i], want_rc[i] = fz.run_oracle(s)
    b = AecmBatch(n_inst, 0)
    assert b.init(fs) == 0
    rc, st = b.set_config(base["cng"], base["echo_mode"])
    assert rc == 0 and not st.any()
    out = np.zeros_like(near)
    pos = 0
    for c in range(n_calls):
        n = int(base["n"][c])
        sl = slice(pos, pos + n)
        acc = np.zeros(n_inst, dtype=np.int64)
        for _ in range(int(base["far_ops"][c])):
            rc, status = b.buffer_farend(np.ascontiguousarray(far[:, sl]))
            assert rc == 0
            acc |= status
        rc, o, status = b.process(np.ascontiguousarray(near[:, sl]), np.ascontiguousarray(cln[:, sl]) if use_clean else None, ms=ms[:, c])
        assert rc == 0
        acc |= status
        out[:, sl] = o
        np.testing.assert_array_equal(acc, want_rc[:, c], err_msg=f"return codes of call {c}")
        pos += n
    bad = np.nonzero((out != want).any(axis=1))[0]
    assert bad.size == 0, f"instances {bad[:8]} differ, first sample {np.argmax(out[bad[0]] != want[bad[0]])}"
    b.close()


@pytest.mark.parametrize("seed,fs", [(2, 16000), (6, 8000), (14, 16000), (22, 16000)])
def test_random_control_calls_in_mid_stream_match_oracle(seed, fs):
    """Control calls in the middle of running streams, on sub-ranges of a batch: AecmB200_SetConfig (valid and refused
    values), AecmB200_Init (same or the other rate -- the range starts over while its neighbours keep running) and
    AecmB200_InitEchoPath between the 10 ms calls of the randomised scripts.  Per-instance status of every control call,
    status of every 10 ms call and all output equal the oracle's, stepped per instance with the events that hit it
    (tests/test_fuzz_api.py pins the same kind of script to the reference)."""
    import test_fuzz_api as fz
    from webrtc_aecm_b200 import AecmBatch

    n_inst, n_calls = 12, 400
    base = fz.make_script(seed, n_calls, fs)
    total = int(base["n"].sum())
    far, near = synth_streams(n_inst, total, fs, seed=seed * 37, delay_ms=(np.arange(n_inst) * 7) % 110)
    cln = synth_streams(n_inst, total, fs, seed=seed * 37 + 1, delay_ms=5)[1]
    use_clean = bool(base["clean"].all())
    rng = np.random.default_rng(seed + 9_000_000)
    events = []  # (call, kind, a, b, first, count)
    for c, kind, a, b in fz.make_control_events(seed, n_calls) + fz.make_control_events(seed + 1000, n_calls):
        first = int(rng.integers(0, n_inst))
        events.append((c, kind, a, b, first, 1 if kind == 4 else int(rng.integers(1, n_inst - first + 1))))
    events.sort(key=lambda e: e[0])
    assert {e[1] for e in events} == {2, 3, 4}
    want = np.zeros_like(near)
    want_rc = np.zeros((n_inst, n_calls), dtype=np.int64)
    want_ev = []
    for i in range(n_inst):
        mine = [(c, kind, a, b) for c, kind, a, b, first, count in events if first <= i < first + count]
        s = dict(base, far=far[i], near=near[i], cln=cln[i])
        rcs = []
        want[i], want_rc[i] = fz.run_oracle(s, mine, rcs)
        want_ev.append(rcs)
    b = AecmBatch(n_inst, 0)
    assert b.init(fs) == 0
    rc, st = b.set_config(base["cng"], base["echo_mode"])
    assert rc == 0 and not st.any()
    got_ev = [[] for _ in range(n_inst)]
    out = np.zeros_like(near)
    pos = k = 0
    for c in range(n_calls):
        while k < len(events) and events[k][0] == c:
            _, kind, a, b_, first, count = events[k]
            if kind == 2:
                rc, st = b.set_config(a, b_, first=first, count=count)
                assert rc == 0
                st = st.tolist()
            elif kind == 3:
                st = [b.init(a, first=first, count=count)] * count
            else:
                st = [b.init_echo_path(first, fz.event_path(a))]
            for j in range(count):
                got_ev[first + j].append(st[j])
            k += 1
        n = int(base["n"][c])
        sl = slice(pos, pos + n)
        acc = np.zeros(n_inst, dtype=np.int64)
        for _ in range(int(base["far_ops"][c])):
            rc, status = b.buffer_farend(np.ascontiguousarray(far[:, sl]))
            assert rc == 0
            acc |= status
        rc, o, status = b.process(np.ascontiguousarray(near[:, sl]), np.ascontiguousarray(cln[:, sl]) if use_clean else None,
                                  ms=np.full(n_inst, int(base["ms"][c]), dtype=np.int16))
        assert rc == 0
        acc |= status
        out[:, sl] = o
        np.testing.assert_array_equal(acc, want_rc[:, c], err_msg=f"return codes of call {c}")
        pos += n
    assert got_ev == want_ev
    bad = np.nonzero((out != want).any(axis=1))[0]
    assert bad.size == 0, f"instances {bad[:8]} differ, first sample {np.argmax(out[bad[0]] != want[bad[0]])}"
    b.close()
