"""Deterministic, integer-only synthetic near/far int16 streams.

Every sample is a pure function of (seed, instance, sample index): a counter-based 32-bit hash drives
white noise, a burst envelope makes it speech-like (on/off segments of 0.25 s with amplitudes from a
small table, digital silence included), and the near end is a delayed, filtered copy of the far end
(an integer 3-tap echo path with per-instance gain) plus near-end noise and occasional double talk.
No floating point is involved, so host and device always see identical bits.

``kind='echo'``   the moderate family used by the throughput bench and most parity tests
``kind='stress'`` adds full-scale / clipping / exact-zero segments on both ends (exercises the
                  saturation, Q-realignment and min-statistics branches listed in SURVEY.md 8c)
"""
from __future__ import annotations

import numpy as np

_AMP_ECHO = np.array([0, 0, 30, 500, 3000, 6000, 12000, 20000], dtype=np.int64)
_AMP_STRESS = np.array([0, 0, 1, 30, 6000, 20000, 32767, 40000], dtype=np.int64)


def _hash32(x: np.ndarray) -> np.ndarray:
    """lowbias32 integer hash on uint32 arrays (wraps mod 2**32)."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _noise16(seed: int, inst: np.ndarray, t: np.ndarray, salt: int) -> np.ndarray:
    key = (np.uint32(seed & 0xFFFFFFFF) ^ (inst.astype(np.uint32) * np.uint32(0x9E3779B9))
           ^ (t.astype(np.uint32) * np.uint32(0x85EBCA6B)) ^ np.uint32(salt))
    return (_hash32(key) >> np.uint32(16)).astype(np.int64) - 32768


def synth_streams(n_instances: int, n_samples: int, fs: int = 16000, seed: int = 0xAEC0,
                  delay_ms=40, kind: str = "echo", first_instance: int = 0, first_sample: int = 0, instances=None):
    """Return (far, near) int16 arrays of shape [n_instances, n_samples].

    ``delay_ms`` is a scalar or a per-instance array: the acoustic echo delay in ms.
    ``first_instance``/``first_sample`` let shards/chunks generate their slice of one global stream;
    ``instances`` (array of global instance numbers, length n_instances) picks arbitrary rows of it instead.
    """
    with np.errstate(over="ignore"):
        if instances is not None:
            inst = np.asarray(instances, dtype=np.int64).reshape(n_instances)[:, None]
        else:
            inst = (np.arange(n_instances, dtype=np.int64) + first_instance)[:, None]
        delay = np.broadcast_to(np.asarray(delay_ms, dtype=np.int64), (n_instances,))[:, None] * (fs // 1000)
        max_d = int(delay.max()) + 4
        # generate the far end with a pre-roll so the delayed copy exists from sample 0
        t = (np.arange(-max_d, n_samples, dtype=np.int64) + first_sample)[None, :]
        seg = np.floor_divide(t, fs // 4)
        amp_tab = _AMP_STRESS if kind == "stress" else _AMP_ECHO
        sel = _hash32((inst.astype(np.uint32) * np.uint32(0x632BE5AB)) ^ (seg.astype(np.uint32) * np.uint32(0x2545F491))
                      ^ np.uint32(seed & 0xFFFFFFFF))
        amp = amp_tab[(sel >> np.uint32(7)) % np.uint32(len(amp_tab))]
        raw = _noise16(seed, inst, t, 0x1234567)
        # mild low-pass so the spectrum is not flat: y[t] = (x[t] + x[t-1]) / 2 (first column kept)
        lp = raw.copy()
        lp[:, 1:] = (raw[:, 1:] + raw[:, :-1]) >> 1
        far_full = np.clip((lp * amp) >> 15, -32768, 32767)
        far_full[t.repeat(n_instances, 0) < 0] = 0  # nothing before the global stream start

        idx = np.arange(n_samples, dtype=np.int64)[None, :] + max_d
        g = 64 + (_hash32(inst.astype(np.uint32) * np.uint32(0x27D4EB2F) + np.uint32(seed & 0xFFFF)) % np.uint32(448)).astype(np.int64)
        rows = np.arange(n_instances)[:, None]
        e0 = far_full[rows, idx - delay]
        e1 = far_full[rows, idx - delay - 1]
        e3 = far_full[rows, idx - delay - 3]
        echo = (g * e0 + (g >> 1) * e1 - (g >> 2) * e3) >> 8  # Q8 gains, 0.25 .. 2.0 main tap

        tn = t[:, max_d:]
        nseg = np.floor_divide(tn, fs // 2)
        nsel = _hash32((inst.astype(np.uint32) * np.uint32(0x9E3779B1)) ^ (nseg.astype(np.uint32) * np.uint32(0x7F4A7C15))
                       ^ np.uint32((seed >> 3) & 0xFFFFFFFF))
        if kind == "stress":
            ntab = np.array([0, 0, 0, 1, 3, 300, 32767, 50000], dtype=np.int64)
        else:
            ntab = np.array([0, 3, 30, 100, 300, 300, 4000, 100], dtype=np.int64)  # one double-talk level
        namp = ntab[(nsel >> np.uint32(9)) % np.uint32(len(ntab))]
        nnoise = (_noise16(seed ^ 0x5bd1e995, inst, tn, 0x7654321) * namp) >> 15
        if kind == "stress":
            # exact digital silence on the near end for some segments (echo muted too)
            mute = ((nsel >> np.uint32(3)) % np.uint32(5)) == 0
            echo = np.where(mute, 0, echo)
            nnoise = np.where(mute, 0, nnoise)
        near = np.clip(echo + nnoise, -32768, 32767)
        far = far_full[:, max_d:]
        return far.astype(np.int16), near.astype(np.int16)


# ---------------------------------------------------------------------------------------------------
# the same generator on a torch device (bit-identical to synth_streams; used by bench.py to fill HBM
# with large batches without a host round trip).  All arithmetic is int64 with explicit 32-bit masks.
# ---------------------------------------------------------------------------------------------------
def _t_hash32(x):
    M = 0xFFFFFFFF
    x = x & M
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & M
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & M
    x = x ^ (x >> 16)
    return x


def synth_streams_torch(n_instances: int, n_samples: int, fs: int = 16000, seed: int = 0xAEC0, delay_ms=40,
                        kind: str = "echo", first_instance: int = 0, first_sample: int = 0, device="cuda", instances=None):
    """torch twin of :func:`synth_streams`; returns (far, near) int16 tensors on ``device``."""
    import torch

    M = 0xFFFFFFFF
    dev = torch.device(device)
    i64 = torch.int64
    if instances is not None:
        inst = torch.as_tensor(instances, dtype=i64, device=dev).reshape(n_instances)[:, None]
    else:
        inst = (torch.arange(n_instances, dtype=i64, device=dev) + first_instance)[:, None]
    delay = torch.as_tensor(delay_ms, dtype=i64, device=dev).expand(n_instances).clone()[:, None] * (fs // 1000)
    max_d = int(delay.max().item()) + 4
    t = (torch.arange(-max_d, n_samples, dtype=i64, device=dev) + first_sample)[None, :]
    tu = t & M  # uint32 view of possibly negative sample indices
    instu = inst & M
    seg = torch.div(t, fs // 4, rounding_mode="floor")
    amp_tab = torch.as_tensor((_AMP_STRESS if kind == "stress" else _AMP_ECHO), dtype=i64, device=dev)
    sel = _t_hash32(((instu * 0x632BE5AB) & M) ^ (((seg & M) * 0x2545F491) & M) ^ (seed & M))
    amp = amp_tab[(sel >> 7) % amp_tab.numel()]

    def noise16(sd, tt, salt):
        key = (sd & M) ^ ((instu * 0x9E3779B9) & M) ^ (((tt & M) * 0x85EBCA6B) & M) ^ salt
        return (_t_hash32(key) >> 16) - 32768

    raw = noise16(seed, tu, 0x1234567)
    lp = raw.clone()
    lp[:, 1:] = (raw[:, 1:] + raw[:, :-1]) >> 1
    far_full = torch.clamp((lp * amp) >> 15, -32768, 32767)
    far_full = torch.where(t.expand_as(far_full) < 0, torch.zeros_like(far_full), far_full)

    idx = torch.arange(n_samples, dtype=i64, device=dev)[None, :] + max_d
    g = 64 + (_t_hash32(((instu * 0x27D4EB2F) & M) + (seed & 0xFFFF)) % 448)
    e0 = torch.gather(far_full, 1, idx - delay)
    e1 = torch.gather(far_full, 1, idx - delay - 1)
    e3 = torch.gather(far_full, 1, idx - delay - 3)
    echo = (g * e0 + (g >> 1) * e1 - (g >> 2) * e3) >> 8

    tn = t[:, max_d:]
    nseg = torch.div(tn, fs // 2, rounding_mode="floor")
    nsel = _t_hash32(((instu * 0x9E3779B1) & M) ^ (((nseg & M) * 0x7F4A7C15) & M) ^ ((seed >> 3) & M))
    ntab = torch.as_tensor([0, 0, 0, 1, 3, 300, 32767, 50000] if kind == "stress" else [0, 3, 30, 100, 300, 300, 4000, 100],
                           dtype=i64, device=dev)
    namp = ntab[(nsel >> 9) % ntab.numel()]
    nnoise = (noise16(seed ^ 0x5bd1e995, tn, 0x7654321) * namp) >> 15
    if kind == "stress":
        mute = ((nsel >> 3) % 5) == 0
        echo = torch.where(mute, torch.zeros_like(echo), echo)
        nnoise = torch.where(mute, torch.zeros_like(nnoise), nnoise)
    near = torch.clamp(echo + nnoise, -32768, 32767)
    far = far_full[:, max_d:]
    return far.to(torch.int16).contiguous(), near.to(torch.int16).contiguous()
