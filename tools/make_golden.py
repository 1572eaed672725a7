#!/usr/bin/env python3
"""Generate tests/golden/ from the UNMODIFIED reference library (oracle/_ref/libaecm_ref.so).

Run in the build container (needs /root/reference to have been compiled by oracle/Makefile):
    python tools/make_golden.py
Writes
    tests/golden/streams.json   per case: sha256 of the reference's int16 output + OR of its return codes
    tests/golden/dly_streams.json     SURVEY 8-f4: output / delay-trace checksums of the reference with robust validation, lookahead
                                      and soft resets switched on through oracle/ref_shim.cc
    tests/golden/fft_adversarial.npz  a larger adversarial FFT/IFFT family with the reference's answers
    tests/golden/vectors.npz    full reference outputs for two small cases, and stage-level known-answer
                                vectors produced by the reference's own C symbols (forward/inverse real FFT
                                incl. the returned scale, WebRtcSpl_SqrtFloor, WebRtcSpl_RandUArray)
The GPU box has no /root/reference: there the tests compare against these files (and against the prebuilt
oracle/_ref when it travelled along).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402
from cases import CASE_BY_NAME, CASES, DLY_CONFIGS, dly_inputs, fft_adversarial_vectors, make_inputs, run_kwargs  # noqa: E402
from webrtc_aecm_b200.shard import stream_checksum  # noqa: E402

FULL_VECTOR_CASES = ("16k_80_sample_calls", "16k_full_scale_alternating")


def stage_vectors(ref):
    rng = np.random.default_rng(20260924)
    ref.WebRtcSpl_CreateRealFFT.restype = C.c_void_p
    ref.WebRtcSpl_RealForwardFFT.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    ref.WebRtcSpl_RealInverseFFT.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    ref.WebRtcSpl_RealInverseFFT.restype = C.c_int
    ref.WebRtcSpl_SqrtFloor.argtypes = [C.c_int32]
    ref.WebRtcSpl_SqrtFloor.restype = C.c_int32
    fft = ref.WebRtcSpl_CreateRealFFT(7)
    # forward: random at several amplitudes + adversarial rows
    fin = [rng.integers(-a, a + 1, size=128) for a in (1, 7, 100, 3000, 16384, 32767) for _ in range(6)]
    fin += [np.zeros(128), np.full(128, 32767), np.full(128, -32768), np.where(np.arange(128) % 2, 32767, -32768),
            np.eye(1, 128, 0)[0] * 32767, np.eye(1, 128, 64)[0] * -32768]
    fin = np.asarray(fin, dtype=np.int16)
    fout = np.zeros((len(fin), 130), dtype=np.int16)
    for i in range(len(fin)):
        ref.WebRtcSpl_RealForwardFFT(fft, fin[i].ctypes.data, fout[i].ctypes.data)
    # inverse: random complex bins incl. values straddling the scaling thresholds 13573 / 27146
    iin = [rng.integers(-a, a + 1, size=130) for a in (1, 50, 3000, 13573, 13574, 20000, 27146, 27147, 32767)
           for _ in range(5)]
    iin += [np.zeros(130), np.full(130, 32767), np.full(130, -32768)]
    iin = np.asarray(iin, dtype=np.int16)
    iout = np.zeros((len(iin), 128), dtype=np.int16)
    iscale = np.zeros(len(iin), dtype=np.int32)
    for i in range(len(iin)):
        buf = np.zeros(260, dtype=np.int16)  # the reference reads n+2 and may be handed a larger buffer
        buf[:130] = iin[i]
        iscale[i] = ref.WebRtcSpl_RealInverseFFT(fft, buf.ctypes.data, iout[i].ctypes.data)
    # a second, larger adversarial family (tests/cases.py), kept in its own file
    afwd, ainv = fft_adversarial_vectors()
    afout = np.zeros((len(afwd), 130), dtype=np.int16)
    for i in range(len(afwd)):
        ref.WebRtcSpl_RealForwardFFT(fft, afwd[i].ctypes.data, afout[i].ctypes.data)
    aiout = np.zeros((len(ainv), 128), dtype=np.int16)
    aiscale = np.zeros(len(ainv), dtype=np.int32)
    for i in range(len(ainv)):
        buf = np.zeros(260, dtype=np.int16)
        buf[:130] = ainv[i]
        aiscale[i] = ref.WebRtcSpl_RealInverseFFT(fft, buf.ctypes.data, aiout[i].ctypes.data)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fft_adversarial.npz"), fwd_in=afwd, fwd_out=afout, inv_in=ainv,
                        inv_out=aiout, inv_scale=aiscale)
    sq_in = np.concatenate([np.arange(0, 70), (np.arange(1, 46341, 97) ** 2), (np.arange(2, 46341, 89) ** 2) - 1,
                            (np.arange(1, 46340, 83) ** 2) + 1, [2**31 - 1, 2**30, -1, -2**31],
                            rng.integers(0, 2**31, size=500)]).astype(np.int64)
    sq_in = np.clip(sq_in, -2**31, 2**31 - 1).astype(np.int32)
    sq_out = np.array([ref.WebRtcSpl_SqrtFloor(int(v)) for v in sq_in], dtype=np.int32)
    seeds = np.array([666, 1, 0, 0x7FFFFFFF, 123456789], dtype=np.uint32)
    rnd = np.zeros((len(seeds), 64), dtype=np.int16)
    seed_after = np.zeros(len(seeds), dtype=np.uint32)
    for i, s in enumerate(seeds):
        sd = C.c_uint32(int(s))
        ref.WebRtcSpl_RandUArray(rnd[i].ctypes.data_as(C.c_void_p), C.c_int16(64), C.byref(sd))
        seed_after[i] = sd.value
    return dict(fft_fwd_in=fin, fft_fwd_out=fout, fft_inv_in=iin, fft_inv_out=iout, fft_inv_scale=iscale,
                sqrt_in=sq_in, sqrt_out=sq_out, rand_seed=seeds, rand_out=rnd, rand_seed_after=seed_after)


def main():
    if ol.ref_path() is None:
        ol.build()
    assert ol.ref_path(), "oracle/_ref/libaecm_ref.so missing: run in the build container"
    streams = {}
    vectors = {}
    for case in CASES:
        far, near, clean, ms = make_inputs(case)
        out, ret, _ = ol.run_streams(far, near, impl="ref", clean=clean, ms=ms, **run_kwargs(case))
        streams[case["name"]] = {"sha256": stream_checksum(out), "ret_or": [int(r) for r in ret],
                                 "shape": list(out.shape), "inputs_sha256": stream_checksum(np.concatenate([far, near]))}
        if case["name"] in FULL_VECTOR_CASES:
            vectors["out__" + case["name"]] = out
        print(case["name"], streams[case["name"]]["sha256"][:16], sorted(set(ret.tolist())))
    # the delay estimator's optional machinery, driven on the reference through oracle/ref_shim.cc
    dly = {}
    for cfg in DLY_CONFIGS:
        fs_, spc_, far, near, ms, kw = dly_inputs(cfg)
        out, ret, _, trace = ol.run_streams(far, near, impl="ref", ms=ms, dly=cfg["dly"], soft_reset=cfg["soft_reset"], delay_trace=True, **kw)
        dly[cfg["id"]] = {"sha256": stream_checksum(out), "delay_trace_sha256": stream_checksum(trace), "ret_or": [int(r) for r in ret],
                          "distinct_delays": int(len(np.unique(trace)))}
        print("dly", cfg["id"], dly[cfg["id"]]["sha256"][:16], dly[cfg["id"]]["distinct_delays"])
    with open(os.path.join(ROOT, "tests", "golden", "dly_streams.json"), "w") as f:
        json.dump({"generator": "tools/make_golden.py", "source": "oracle/_ref/libaecm_ref_shim.so (unmodified reference + oracle/ref_shim.cc)",
                   "inputs": "tests/cases.py DLY_CONFIGS on dly_inputs(cfg): the delay-sweep case, or the entry's own stream", "cases": dly}, f, indent=1, sort_keys=True)
    vectors.update(stage_vectors(C.CDLL(ol.ref_path())))
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    with open(os.path.join(gdir, "streams.json"), "w") as f:
        json.dump({"generator": "tools/make_golden.py", "source": "oracle/_ref/libaecm_ref.so (unmodified reference)",
                   "cases": streams}, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(gdir, "vectors.npz"), **vectors)
    print("wrote", gdir)


if __name__ == "__main__":
    main()
