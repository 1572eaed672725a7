"""The scalar primitives of the device code (aecm_prims.cuh), compiled for the host through tests/host_emu, against the
reference's own SPL functions where it exports them (WebRtcSpl_DivW32W16, signal_processing_library.cc:107-123), against
the oracle's restatements (pinned elsewhere) and against plain integer arithmetic."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from test_call_path_emulated_warp import emu  # noqa: F401  (fixture: builds tests/host_emu/libcall_emu.so)

EDGE = [0, 1, -1, 2, -2, 3, 64, 65, 66, 127, 128, 32767, 32768, -32768, -32769, 65535, 65536, (1 << 30) - 1, 1 << 30, -(1 << 30),
        (1 << 31) - 2, (1 << 31) - 1, -(1 << 31), -(1 << 31) + 1]


def test_division_by_bin_index_is_exact(emu):  # noqa: F811
    """div_by_bin(n, d) = multiply-high by a per-divisor constant: must equal C's truncating n / d (aecm_core.cc:904:
    WebRtcSpl_DivW32W16(tmp32no2, i + 1)) for every divisor 1..65 -- on edge values, on multiples of d and their neighbours,
    and on a random sweep; plus the sufficient condition that makes the constants exact for EVERY |n| <= 2^31."""
    rng = np.random.default_rng(65)
    ref = None
    if ol.ref_path():
        ref = C.CDLL(ol.ref_path())
        ref.WebRtcSpl_DivW32W16.restype = C.c_int32
        ref.WebRtcSpl_DivW32W16.argtypes = [C.c_int32, C.c_int16]
    emu.emu_div_by_bin.restype = C.c_int
    for d in range(1, 66):
        vals = set(EDGE)
        for k in rng.integers(0, (1 << 31) // d, 40):
            vals.update({int(k) * d, int(k) * d - 1, int(k) * d + 1, -int(k) * d, -int(k) * d + 1, -int(k) * d - 1})
        vals.update(int(v) for v in rng.integers(-(1 << 31), 1 << 31, 300))
        for n in vals:
            if not -(1 << 31) <= n < (1 << 31):
                continue
            want = abs(n) // d * (1 if n >= 0 else -1)  # C semantics: truncation toward zero
            if n == -(1 << 31) and d == 1:
                want = n
            got = emu.emu_div_by_bin(C.c_int(n), d)
            assert got == want, (n, d, got, want)
            if ref is not None:
                assert ref.WebRtcSpl_DivW32W16(n, d) == want, (n, d)
        if d >= 2:  # Granlund-Montgomery: M = ceil(2^(32+s') / d) is exact for all 0 <= n <= 2^31 iff (M*d - 2^(32+s')) * 2^31 < 2^(32+s')
            s = d.bit_length() - 1
            sp = s - 1 if d == 1 << s else s
            m = -(-(1 << (32 + sp)) // d)
            assert m < 1 << 32 and (m * d - (1 << (32 + sp))) << 31 < 1 << (32 + sp), d


def test_norms_sqrt_log_and_saturation_match_the_oracle(emu):  # noqa: F811
    O = ol.oracle()
    O.oracle_norm_w32.argtypes, O.oracle_norm_u32.argtypes, O.oracle_norm_w16.argtypes = [C.c_int32], [C.c_uint32], [C.c_int16]
    rng = np.random.default_rng(7)
    vals = [v for v in EDGE if -(1 << 31) <= v < (1 << 31)] + [int(v) for v in rng.integers(-(1 << 31), 1 << 31, 4000)]
    vals += [int(1 << k) for k in range(31)] + [int((1 << k) - 1) for k in range(1, 32)] + [-int(1 << k) for k in range(32)]
    for v in vals:
        assert emu.emu_norm_w32(C.c_int(v)) == O.oracle_norm_w32(v), v
        assert emu.emu_norm_u32(C.c_uint(v & 0xFFFFFFFF)) == O.oracle_norm_u32(v & 0xFFFFFFFF), v
        v16 = ((v + 32768) & 0xFFFF) - 32768
        assert emu.emu_norm_w16(C.c_int(v16)) == O.oracle_norm_w16(v16), v16
        if v >= 0:
            assert emu.emu_sqrt_floor(C.c_int(v)) == O.oracle_sqrt_floor(v) == int(np.floor(np.sqrt(np.float64(v)))), v
    for a, b in zip(vals[::3], vals[1::3]):
        want = max(-(1 << 31), min((1 << 31) - 1, a + b))  # WebRtcSpl_AddSatW32, spl_inl.h:70-82
        assert emu.emu_add_sat32(C.c_int(a), C.c_int(b)) == want, (a, b)
    # LogOfEnergyInQ8 (aecm_core.cc:612-628): 896 + ((31 - clz) << 8) + top 8 fraction bits - (q << 8), as int16; 896 for e == 0
    for e in [0, 1, 2, 3, 255, 256, 65535, 1 << 20, (1 << 31) - 1, 1 << 31, (1 << 32) - 1] + [int(x) for x in rng.integers(0, 1 << 32, 500)]:
        for q in (0, 3, 12, 15, 27):
            if e == 0:
                want = 896
            else:
                z = 32 - e.bit_length()
                frac = (((e << z) & 0x7FFFFFFF) >> 23)
                want = ((896 + ((31 - z) << 8) + frac - (q << 8) + 32768) & 0xFFFF) - 32768
            assert emu.emu_log_energy_q8(C.c_uint(e), q) == want, (e, q)


def test_lcg_jump_ahead_equals_the_references_generator(emu):  # noqa: F811
    """The device draws the k-th comfort-noise number directly (seed * A^k + C_k mod 2^31, a compile-time table) instead of
    stepping the generator 64 times: against WebRtcSpl_RandUArray's own outputs (tests/golden/vectors.npz, generated from the
    reference, signal_processing_library.cc:126-147) and against the oracle's stepping generator on random seeds."""
    vec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.npz"))
    emu.emu_lcg_draws.restype = C.c_uint32
    O = ol.oracle()
    rng = np.random.default_rng(11)
    seeds = [(int(s), np.asarray(w), int(a)) for s, w, a in zip(vec["rand_seed"], vec["rand_out"], vec["rand_seed_after"])]
    for s in [666, 0, 1, 0x7FFFFFFF] + [int(x) for x in rng.integers(0, 1 << 31, 200)]:
        buf = np.zeros(64, dtype=np.int16)
        sd = C.c_uint32(s)
        O.oracle_rand_u_array(buf.ctypes.data_as(C.c_void_p), 64, C.byref(sd))
        seeds.append((s, buf, sd.value))
    assert len(seeds) > 200
    for s, want, after in seeds:
        got = np.zeros(64, dtype=np.int16)
        assert emu.emu_lcg_draws(C.c_uint32(s), got.ctypes.data_as(C.c_void_p)) == after, s
        np.testing.assert_array_equal(got, want[:64])
