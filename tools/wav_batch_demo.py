#!/usr/bin/env python3
"""Throughput demo for aecm_wav_batch: N synthetic 16 kHz pairs of ragged length (8-20 s), one run.  Prints the tool's
own summary line.   usage: python tools/wav_batch_demo.py [pairs=512] [extra tool args...]"""
import os
import subprocess
import sys
import tempfile
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from webrtc_aecm_b200.synth import synth_streams  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
fs = 16000
rng = np.random.default_rng(1)
with tempfile.TemporaryDirectory() as d:
    far, near = synth_streams(n, fs * 20, fs, seed=99, delay_ms=np.arange(n) % 100)
    args = []
    for i in range(n):
        m = int(rng.integers(8 * fs, 20 * fs))
        for name, x in (("far", far[i, :m]), ("near", near[i, :m])):
            with wave.open(f"{d}/{name}{i}.wav", "wb") as w:
                w.setnchannels(1); w.setsampwidth(2); w.setframerate(fs)
                w.writeframes(np.ascontiguousarray(x, dtype="<i2").tobytes())
        args += [f"{d}/far{i}.wav", f"{d}/near{i}.wav"]
    lst = f"{d}/pairs.txt"
    with open(lst, "w") as f:
        for i in range(n):
            f.write(f"{d}/far{i}.wav {d}/near{i}.wav\n")
    r = subprocess.run([os.path.join(ROOT, "webrtc_aecm_b200", "aecm_wav_batch"), *sys.argv[2:], "--list", lst], capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr.strip())
