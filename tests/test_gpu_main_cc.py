"""BASELINE configs[0]: the reference's own CLI (main.cc, unmodified) on one 16 kHz mono WAV pair.  The same
main.cc is linked once against the reference library and once against libaecm_b200.so (through the drop-in
header include/aecm/echo_control_mobile.h); the two output WAV files must be byte-identical."""
import os
import shutil
import subprocess
import wave

import numpy as np
import pytest

from webrtc_aecm_b200.synth import synth_streams

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "aecm_run_ref")
OUR_BIN = os.path.join(ROOT, "oracle", "_ref", "aecm_run_b200")


def write_wav(path, samples, fs):
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(fs)
        w.writeframes(np.ascontiguousarray(samples, dtype="<i2").tobytes())


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(OUR_BIN)), reason="CLI binaries not built (need /root/reference at build time)")
@pytest.mark.parametrize("fs", [16000, 8000])
def test_main_cc_wav_pair_is_byte_identical(tmp_path, fs):
    far, near = synth_streams(1, fs * 5, fs, seed=4242, delay_ms=60)
    outs = {}
    for tag, binary in (("ref", REF_BIN), ("b200", OUR_BIN)):
        d = tmp_path / tag
        d.mkdir()
        write_wav(str(d / "far.wav"), far[0], fs)
        write_wav(str(d / "near.wav"), near[0], fs)
        res = subprocess.run([binary, str(d / "far.wav"), str(d / "near.wav")], stdin=subprocess.DEVNULL,
                             capture_output=True, text=True, timeout=300)  # main.cc:193 getchar() sees EOF
        assert res.returncode == 0, res.stderr
        produced = [f for f in os.listdir(d) if f not in ("far.wav", "near.wav")]
        assert len(produced) == 1, produced  # main.cc writes <near>_out.wav
        with wave.open(str(d / produced[0]), "rb") as w:
            outs[tag] = w.readframes(w.getnframes())
    assert len(outs["ref"]) == len(outs["b200"]) > 0
    assert outs["ref"] == outs["b200"]
    assert outs["ref"] != np.ascontiguousarray(near[0], dtype="<i2").tobytes()  # the canceller did act


def test_c_examples_run_on_the_gpu(tmp_path):
    """examples/*.c built against include/ and run on the device: the throughput loop of INTEGRATION.md and the delay
    estimator's optional machinery (its parity is tested in test_gpu_parity.py, here only that the plain-C
    program works end to end)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "webrtc_aecm_b200")
    outs = {}
    for name, args in (("batch_stream", ["512", "60"]), ("delay_estimator_options", [])):
        exe = str(tmp_path / name)
        r = subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "examples", name + ".c"), "-L", lib_dir,
                            "-laecm_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        run = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)
        assert run.returncode == 0, (run.stdout, run.stderr)
        outs[name] = run.stdout
    lines = [l for l in outs["delay_estimator_options"].splitlines() if l.startswith("instance")]
    assert len(lines) == 8
    delays = [int(l.split("delay estimate")[1].split()[0]) for l in lines]
    assert all(-2 <= d < 100 for d in delays), delays  # WebRtc_last_delay: -2 (no estimate yet) or a history index
    assert "soft reset by 2 blocks" in outs["delay_estimator_options"]
