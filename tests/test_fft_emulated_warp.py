"""CPU check of DEVICE code: webrtc_aecm_b200/csrc/aecm_fft.cuh (the register-resident warp FFT) is compiled for
the host against tests/host_emu/warp_emu.h (32 threads + barriers emulate a warp: __syncwarp, redux, PRMT, ...)
and must reproduce the reference's forward / inverse real FFT known answers (tests/golden/vectors.npz, produced by
the reference's own WebRtcSpl_RealForwardFFT / WebRtcSpl_RealInverseFFT), including the returned scale."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emu", "fft_emu.cpp")
LIB = os.path.join(HERE, "host_emu", "libfft_emu.so")
VEC = np.load(os.path.join(HERE, "golden", "vectors.npz"))


@pytest.fixture(scope="module")
def emu():
    deps = [SRC, os.path.join(HERE, "host_emu", "warp_emu.h"),
            os.path.join(HERE, "..", "webrtc_aecm_b200", "csrc", "aecm_fft.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC, "-lpthread"])
    lib = C.CDLL(LIB)
    lib.emu_fft128_inverse.restype = C.c_int
    return lib


def test_forward_fft_matches_reference_vectors(emu):
    fin, fout = VEC["fft_fwd_in"], VEC["fft_fwd_out"]
    got = np.zeros_like(fout)
    for i in range(len(fin)):
        emu.emu_fft128_forward(fin[i].ctypes.data_as(C.c_void_p), got[i].ctypes.data_as(C.c_void_p), 1)
    np.testing.assert_array_equal(got, fout)


def test_forward_fft_two_signals_interleaved(emu):
    fin, fout = VEC["fft_fwd_in"], VEC["fft_fwd_out"]
    for a, b in ((3, 20), (35, 0), (41, 36)):
        pair = np.ascontiguousarray(np.stack([fin[a], fin[b]]))
        got = np.zeros((2, 130), dtype=np.int16)
        emu.emu_fft128_forward(pair.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), 2)
        np.testing.assert_array_equal(got[0], fout[a])
        np.testing.assert_array_equal(got[1], fout[b])


def test_inverse_fft_and_scale_match_reference_vectors(emu):
    iin, iout, iscale = VEC["fft_inv_in"], VEC["fft_inv_out"], VEC["fft_inv_scale"]
    got = np.zeros_like(iout)
    scale = np.zeros_like(iscale)
    for i in range(len(iin)):
        scale[i] = emu.emu_fft128_inverse(iin[i].ctypes.data_as(C.c_void_p), got[i].ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(scale, iscale)
    np.testing.assert_array_equal(got, iout)


def test_adversarial_family_matches_reference_answers(emu):
    """tests/golden/fft_adversarial.npz (reference answers for tests/cases.py:fft_adversarial_vectors -- impulses, full scale,
    per-stage peaks straddling 13573 / 27146, scales 0..8): the device FFT on the emulated warp, both forward variants."""
    A = np.load(os.path.join(HERE, "golden", "fft_adversarial.npz"))
    fwd, inv = np.ascontiguousarray(A["fwd_in"]), np.ascontiguousarray(A["inv_in"])
    got = np.zeros_like(A["fwd_out"])
    for i in range(len(fwd)):
        emu.emu_fft128_forward(fwd[i].ctypes.data_as(C.c_void_p), got[i].ctypes.data_as(C.c_void_p), 1)
    np.testing.assert_array_equal(got, A["fwd_out"])
    for i in range(0, len(fwd) - 1, 2):
        pair = np.ascontiguousarray(fwd[i:i + 2])
        g2 = np.zeros((2, 130), dtype=np.int16)
        emu.emu_fft128_forward(pair.ctypes.data_as(C.c_void_p), g2.ctypes.data_as(C.c_void_p), 2)
        np.testing.assert_array_equal(g2, A["fwd_out"][i:i + 2])
    got, sc = np.zeros_like(A["inv_out"]), np.zeros(len(inv), dtype=np.int32)
    for i in range(len(inv)):
        sc[i] = emu.emu_fft128_inverse(inv[i].ctypes.data_as(C.c_void_p), got[i].ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(sc, A["inv_scale"])
    np.testing.assert_array_equal(got, A["inv_out"])
