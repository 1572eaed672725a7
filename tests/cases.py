"""Named parity cases shared by the oracle tests, the GPU tests and tools/make_golden.py.

Each case is a dict of API-level parameters; ``make_inputs(case)`` builds the int16 planes deterministically.
The families follow SURVEY.md section 8(c): both rates, echoMode 0-4, CNG on/off, clean path, stress audio
(silence / full scale / clipping), msInSndCardBuf extremes incl. clamped values, far-end underrun, both
frame sizes at both rates, and the per-instance delay sweep of BASELINE.json configs[4]."""
from __future__ import annotations

import numpy as np

from webrtc_aecm_b200.synth import synth_streams

MS_MIX = [0, 40, 120, 500, -5, 600, 10, 250]


def _c(name, **kw):
    d = dict(name=name, fs=16000, spc=None, n=4, seconds=4.0, echo_mode=3, cng=1, clean=False, kind="echo", ms=None,
             skip_far_every=0, delay_ms="sweep20", seed=101, special=None)
    d.update(kw)
    if d["spc"] is None:
        d["spc"] = 160 if d["fs"] == 16000 else 80
    return d


CASES = [
    _c("c2_16k_default_10s", seconds=10.0, n=4),
    _c("c3_8k_default_10s", fs=8000, seconds=10.0, n=4),
    *[_c(f"16k_echomode{m}", echo_mode=m, seed=110 + m) for m in range(5)],
    *[_c(f"8k_echomode{m}", fs=8000, echo_mode=m, seed=120 + m, seconds=3.0) for m in (0, 4)],
    _c("16k_cng_off", cng=0, seed=130),
    _c("8k_cng_off", fs=8000, cng=0, seed=131),
    _c("16k_clean_path", clean=True, seed=140),
    _c("8k_clean_path", fs=8000, clean=True, seed=141),
    _c("16k_stress", kind="stress", seed=150, seconds=8.0),
    _c("8k_stress", fs=8000, kind="stress", seed=151, seconds=8.0),
    _c("16k_stress_clean_cng_off_mode4", kind="stress", clean=True, cng=0, echo_mode=4, seed=152),
    _c("16k_ms_mix", ms=MS_MIX, n=8, seed=160),
    _c("8k_ms_mix", fs=8000, ms=MS_MIX, n=8, seed=161),
    _c("16k_far_underrun", skip_far_every=7, seed=170),
    _c("8k_far_underrun_ms_mix", fs=8000, skip_far_every=2, ms=MS_MIX, n=8, seed=171),
    _c("16k_80_sample_calls", spc=80, seed=180, seconds=2.0),
    _c("8k_160_sample_calls", fs=8000, spc=160, seed=181),
    _c("c5_delay_sweep_0_120ms", n=16, delay_ms="sweep8", seconds=6.0, seed=190),
    _c("16k_silence", special="silence", seconds=2.0),
    _c("16k_far_only", special="far_only", kind="stress", seconds=3.0, seed=200),
    _c("16k_full_scale_alternating", special="fullscale", seconds=2.0),
    _c("8k_full_scale_alternating", fs=8000, special="fullscale", seconds=2.0),
    # adversarial families that drive the NLMS rescaling shift (aecm_core.cc:906-912) beyond 31 bits -- undefined
    # in C, x86 SAR semantics in the reference as built (found with a UBSan build of the reference)
    _c("16k_adv_gated_tone_dc_near", special="adv_gated", seconds=12.0, seed=316),
    _c("8k_adv_gated_tone_dc_near_mode0", fs=8000, special="adv_gated", seconds=12.0, seed=300, echo_mode=0, cng=0),
    _c("8k_adv_tone_far_whisper_near", fs=8000, special="adv_whisper", seconds=12.0, seed=303, echo_mode=0, cng=0),
    _c("16k_adv_loud_far_zero_near_mode4", special="adv_zero_near", seconds=12.0, seed=304, echo_mode=4, ms=[0, 120, 500, 40]),
]
CASE_BY_NAME = {c["name"]: c for c in CASES}


def make_inputs(case):
    """-> (far, near, clean_or_None, ms_or_None) int16 planes [n, T]."""
    fs, n = case["fs"], case["n"]
    total = int(case["seconds"] * 100) * (fs // 100)
    total -= total % case["spc"]
    if case["delay_ms"] == "sweep20":
        delay = (np.arange(n) * 20) % 121
    elif case["delay_ms"] == "sweep8":
        delay = (np.arange(n) * 8) % 121
    else:
        delay = case["delay_ms"]
    far, near = synth_streams(n, total, fs, seed=case["seed"], delay_ms=delay, kind=case["kind"])
    if case["special"] == "silence":
        far[:] = 0
        near[:] = 0
    elif case["special"] == "far_only":
        near[:] = 0
    elif case["special"] == "fullscale":
        far[:] = 32767
        far[:, ::2] = -32768
        near = far.copy()
    elif case["special"] and case["special"].startswith("adv_"):
        rng = np.random.default_rng(case["seed"])
        t = np.arange(total)
        f0 = rng.integers(100, fs // 2 - 100, size=(n, 1))
        tone = (30000 * np.sin(2 * np.pi * f0 * t[None, :] / fs)).astype(np.int16)
        if case["special"] == "adv_gated":      # loud far tone gated at 8 Hz, near = constant 1 LSB
            far = (tone * ((t // (fs // 8)) % 2).astype(np.int16)).astype(np.int16)
            near = np.full((n, total), 1, np.int16)
        elif case["special"] == "adv_whisper":  # loud far tone, near = sparse +-1 LSB clicks
            far = tone
            near = ((rng.random((n, total)) < 0.02) * rng.choice([-1, 1], size=(n, total))).astype(np.int16)
        else:                                   # full-scale far noise, near digital zero
            far = rng.integers(-32768, 32767, size=(n, total)).astype(np.int16)
            near = np.zeros((n, total), np.int16)
    clean = None
    if case["clean"]:
        clean = synth_streams(n, total, fs, seed=case["seed"] + 1000, delay_ms=10, kind="echo")[1]
    ms = None if case["ms"] is None else np.resize(np.asarray(case["ms"], dtype=np.int16), n)
    return far, near, clean, ms


def far_mask(case, n_calls):
    """[n_calls, n] uint8 mask equivalent to the runner's skip_far_every."""
    k = case["skip_far_every"]
    m = np.ones((n_calls, case["n"]), dtype=np.uint8)
    if k > 0:
        m[(np.arange(n_calls) % k) == k - 1, :] = 0
    return m
