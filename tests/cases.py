"""Named parity cases shared by the oracle tests, the GPU tests and tools/make_golden.py.

Each case is a dict of API-level parameters; ``make_inputs(case)`` builds the int16 planes deterministically.
The families follow SURVEY.md section 8(c): both rates, echoMode 0-4, CNG on/off, clean path, stress audio
(silence / full scale / clipping), msInSndCardBuf extremes incl. clamped values, far-end underrun, both
frame sizes at both rates, the per-instance delay sweep of BASELINE.json configs[4], and -- found with
tools/ref_coverage.py (gcov over the reference) -- msInSndCardBuf that changes from call to call (start-up
stability check, EstBufDelay's delay-change logic), far-end overrun (two BufferFarend per Process), a constant
far level (minimum step size) and a near end that stays a few LSB for half a minute (noise-floor counters)."""
from __future__ import annotations

import numpy as np

from webrtc_aecm_b200.synth import synth_streams

MS_MIX = [0, 40, 120, 500, -5, 600, 10, 250]


def _c(name, **kw):
    d = dict(name=name, fs=16000, spc=None, n=4, seconds=4.0, echo_mode=3, cng=1, clean=False, kind="echo", ms=None,
             skip_far_every=0, extra_far_every=0, ms_schedule=None, delay_ms="sweep20", seed=101, special=None)
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
    # families added after measuring which reference lines the cases above never reach (oracle/COVERAGE.md)
    _c("16k_ms_jumps", ms_schedule="jumps", seconds=8.0, seed=400),          # echo_control_mobile.cc:552-566
    _c("8k_ms_jumps_underrun", fs=8000, ms_schedule="jumps", skip_far_every=5, seconds=8.0, seed=401),
    _c("16k_ms_unstable_startup", ms_schedule="startup", seconds=3.0, seed=402, n=8),  # :305-337
    _c("16k_far_overrun", extra_far_every=3, seed=410, seconds=6.0),         # :543-546 (two BufferFarend per Process)
    _c("8k_far_overrun_ms_jumps", fs=8000, extra_far_every=2, ms_schedule="jumps", seed=411, seconds=6.0),
    _c("8k_constant_far_level", fs=8000, special="const_far", seconds=6.0, seed=420),   # aecm_core.cc:778-779
    _c("8k_near_few_lsb_30s", fs=8000, special="tiny_near", seconds=30.0, seed=430, n=2),  # aecm_core_c.cc:117-124
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
    elif case["special"] == "const_far":   # stationary far end: min and max trackers of the far energy meet
        rng = np.random.default_rng(case["seed"])
        far = rng.integers(-4000, 4000, size=(n, total)).astype(np.int16)
        near = (far // 3 + rng.integers(-20, 20, size=(n, total))).astype(np.int16)
    elif case["special"] == "tiny_near":   # the near end stays within a few LSB: the noise-floor estimate creeps
        rng = np.random.default_rng(case["seed"])
        near = rng.integers(-2, 3, size=(n, total)).astype(np.int16)
    ms = None if case["ms"] is None else np.resize(np.asarray(case["ms"], dtype=np.int16), n)
    if case["ms_schedule"]:
        ms = ms_schedule(case, total // case["spc"])
    return far, near, clean, ms


def ms_schedule(case, n_calls):
    """[n, n_calls] int16: msInSndCardBuf per call.  'jumps': piecewise constant with large steps both ways every
    0.3-1 s (per instance its own pattern); 'startup': values that wobble by more than the +-20 % / 8 ms tolerance during
    the first calls, then settle -- each instance settles at a different call."""
    n = case["n"]
    rng = np.random.default_rng(case["seed"] + 7)
    ms = np.zeros((n, n_calls), dtype=np.int16)
    if case["ms_schedule"] == "jumps":
        levels = np.array([0, 10, 40, 60, 120, 200, 350, 500, 20, 480, 33, 77], dtype=np.int16)
        for i in range(n):
            c = 0
            while c < n_calls:
                seg = int(rng.integers(30, 100))
                ms[i, c:c + seg] = levels[int(rng.integers(0, len(levels)))]
                c += seg
                if rng.random() < 0.5:  # a two-call spike to the top followed by a drop to zero (one-call swings of the
                    ms[i, c:c + 2] = 500  # filtered delay across both EstBufDelay thresholds)
                    ms[i, c + 2:c + 4] = 0
                    c += 4
    else:
        for i in range(n):
            settle = 2 + 7 * i  # instance 0 settles at once, the last one only after the 0.5 s give-up limit
            ms[i, :] = 40 + 10 * i
            ms[i, :settle] = rng.choice([5, 40, 150, 300, 90], size=min(settle, n_calls))[:n_calls]
    return ms


def soak_inputs(fs, n=12, seconds=120.0, seed=777):
    """Two minutes of audio per instance for the slow state machines (store/reset decisions every >= 30 qualifying
    blocks, VAD update halts after 1024 blocks, delay-change counters): alternating echo and stress material, per
    instance its own echo delay and gain pattern.  -> far, near [n, T] int16"""
    total = int(seconds * 100) * (fs // 100)
    seg = total // 6
    fars, nears = [], []
    for k in range(6):
        kind = "stress" if k % 3 == 2 else "echo"
        f, x = synth_streams(n, seg, fs, seed=seed + k, delay_ms=(np.arange(n) * 11 + 7 * k) % 121, kind=kind)
        fars.append(f)
        nears.append(x)
    far, near = np.concatenate(fars, axis=1), np.concatenate(nears, axis=1)
    pad = total - far.shape[1]
    if pad:
        far, near = np.pad(far, ((0, 0), (0, pad))), np.pad(near, ((0, 0), (0, pad)))
    return np.ascontiguousarray(far), np.ascontiguousarray(near)


def run_kwargs(case):
    """The case's API-level parameters as keyword arguments of oracle_lib.run_streams."""
    return dict(fs=case["fs"], samples_per_call=case["spc"], cng_mode=case["cng"], echo_mode=case["echo_mode"],
                skip_far_every=case["skip_far_every"], extra_far_every=case["extra_far_every"])


def far_mask(case, n_calls):
    """[n_calls, n] uint8 mask equivalent to the runner's skip_far_every."""
    k = case["skip_far_every"]
    m = np.ones((n_calls, case["n"]), dtype=np.uint8)
    if k > 0:
        m[(np.arange(n_calls) % k) == k - 1, :] = 0
    return m


def fft_adversarial_vectors(rng=None):
    """Inputs for the device FFT known-answer test beyond the committed golden set: random, impulses, full scale,
    alternating extremes, and inverse inputs whose per-stage peaks straddle the block-floating thresholds 13573 / 27146
    (complex_fft.c:386-396)."""
    rng = rng or np.random.default_rng(128)
    fwd = [rng.integers(-32768, 32768, 128), rng.integers(-2000, 2000, 128), np.zeros(128), np.full(128, 32767),
           np.full(128, -32768), np.where(np.arange(128) % 2, 32767, -32768), np.eye(1, 128, 0)[0] * 32767,
           np.eye(1, 128, 127)[0] * -32768]
    fwd += [rng.integers(-a, a + 1, 128) for a in (1, 3, 100, 8191, 16384, 23170, 32767) for _ in range(6)]
    inv = []
    for amp in (0, 1, 50, 848, 849, 1696, 1697, 3393, 3394, 6786, 6787, 13572, 13573, 13574, 27145, 27146, 27147, 32767):
        for shape in range(4):
            v = np.zeros(130, dtype=np.int64)
            if shape == 0:
                v[:] = rng.integers(-amp, amp + 1, 130)
            elif shape == 1:
                v[2 * int(rng.integers(0, 65))] = amp
            elif shape == 2:
                v[::2] = amp
            else:
                v[:] = np.where(rng.integers(0, 2, 130) > 0, amp, -amp)
            v[1] = 0  # bins 0 and 64 are real (aecm_core_c.cc:206-214 hands them over that way)
            v[129] = 0
            inv.append(v)
    inv += [np.where(np.arange(130) % 2, 0, -32768)]
    return np.array(fwd, dtype=np.int16), np.array(inv, dtype=np.int16)


# configurations of the delay estimator's optional machinery (oracle_lib.run_streams dly= / soft_reset=),
# run on the delay-sweep inputs (per-instance echo delay 0..120 ms) where the estimate actually moves
DLY_CONFIGS = [
    dict(id="statistics_only", dly=dict(robust=0), soft_reset=None, instances=(0, 9)),
    dict(id="robust", dly=dict(robust=1), soft_reset=None, instances=(1, 7, 14)),
    dict(id="robust_offset3", dly=dict(robust=1, allowed_offset=3), soft_reset=None, instances=(2, 11)),
    dict(id="lookahead5", dly=dict(robust=0, max_lookahead=5), soft_reset=None, instances=(3, 12)),
    dict(id="robust_lookahead8_at3", dly=dict(robust=1, max_lookahead=8, lookahead=3), soft_reset=None, instances=(4, 13)),
    dict(id="robust_lookahead4_softreset", dly=dict(robust=1, max_lookahead=4), soft_reset=(97, 3), instances=(5, 10)),
    dict(id="softreset_only", dly=None, soft_reset=(50, 7), instances=(6, 15)),
    dict(id="robust_softreset_back20", dly=dict(robust=1), soft_reset=(37, -20), instances=(8, 0)),
    # 45 s of steady echo: the winning histogram bin runs into its cap of 3000 (delay_estimator.cc:120-122) after ~27-40 s
    dict(id="robust_long_steady", dly=dict(robust=1, allowed_offset=1), soft_reset=None, instances=(1,),
         inputs=dict(n=3, seconds=45, seed=4711, delay_ms=(20, 60, 100))),
]


def dly_inputs(cfg):
    """(fs, spc, far, near, ms, run_kwargs) of a DLY_CONFIGS entry: the delay-sweep case unless the entry brings its own."""
    from webrtc_aecm_b200.synth import synth_streams

    if "inputs" not in cfg:
        case = CASE_BY_NAME["c5_delay_sweep_0_120ms"]
        far, near, _, ms = make_inputs(case)
        return case["fs"], case["spc"], far, near, ms, run_kwargs(case)
    i = cfg["inputs"]
    far, near = synth_streams(i["n"], i["seconds"] * 16000, 16000, seed=i["seed"], delay_ms=list(i["delay_ms"]))
    return 16000, 160, far, near, None, dict(fs=16000, samples_per_call=160, cng_mode=1, echo_mode=3, skip_far_every=0, extra_far_every=0)
