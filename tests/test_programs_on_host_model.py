"""Whole programs on the host model, under AddressSanitizer + UBSan, without a GPU: the C examples (examples/*.c) and the
reference's own CLI -- main.cc, unmodified, compiled from where it lies -- linked against the product's host layer
(csrc/echo_control_mobile.c + csrc/aecm_batch.c, unchanged) on the modelled CUDA runtime (tests/host_emu/host_model.c, lazy
execution, the oracle behind the launches).  main.cc's output WAV must be byte-identical to what the same main.cc writes when
linked against the reference library (oracle/_ref/aecm_run_ref).  The GPU suite runs the same programs over the shipped
library (tests/test_gpu_main_cc.py); here the point is the host code's memory safety on real call patterns."""
import os
import subprocess

import numpy as np
import pytest

from test_wav_batch_io import REF_BIN, write_wav
from webrtc_aecm_b200.synth import synth_streams

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "webrtc_aecm_b200", "csrc")
REF = "/root/reference"
SAN = ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-O1", "-g"]
INC = ["-I", os.path.join(HERE, "host_emu", "fake_cuda"), "-I", os.path.join(ROOT, "include"), "-I", CSRC]
ENV = dict(os.environ, ASAN_OPTIONS="detect_leaks=1", FAKE_CUDA_DEVICES="2")


@pytest.fixture(scope="module")
def host_objects(tmp_path_factory):
    """the host layer + the model + the oracle, compiled once as C objects"""
    d = tmp_path_factory.mktemp("hostobjs")
    objs = []
    for src in (os.path.join(CSRC, "aecm_batch.c"), os.path.join(CSRC, "echo_control_mobile.c"), os.path.join(HERE, "host_emu", "host_model.c"),
                os.path.join(ROOT, "oracle", "aecm_oracle.c")):
        o = str(d / (os.path.basename(src) + ".o"))
        r = subprocess.run(["gcc", *SAN, "-std=gnu11", *INC, "-c", src, "-o", o], capture_output=True, text=True)
        if r.returncode != 0 and "asan" in r.stderr.lower() and "cannot find" in r.stderr.lower():
            pytest.skip("AddressSanitizer runtime not available")
        assert r.returncode == 0, r.stderr
        objs.append(o)
    return objs


def clean(res):
    assert "Sanitizer" not in res.stderr and "runtime error:" not in res.stderr, res.stderr[-3000:]


@pytest.mark.parametrize("name,args", [("batch_stream", ["64", "40"]), ("delay_estimator_options", [])])
def test_c_examples(host_objects, tmp_path, name, args):
    exe = str(tmp_path / name)
    r = subprocess.run(["gcc", *SAN, "-std=gnu11", *INC, os.path.join(ROOT, "examples", name + ".c"), *host_objects, "-o", exe, "-lpthread", "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600, env=ENV)
    assert res.returncode == 0, (res.stdout, res.stderr[-2000:])
    clean(res)


@pytest.mark.skipif(not (os.path.exists(os.path.join(REF, "main.cc")) and os.path.exists(REF_BIN)), reason="needs /root/reference (main.cc) and oracle/_ref/aecm_run_ref")
@pytest.mark.parametrize("fs", [16000, 8000])
def test_unmodified_main_cc_over_the_host_layer(host_objects, tmp_path, fs):
    exe = str(tmp_path / "aecm_run_hostmodel")
    r = subprocess.run(["g++", *SAN, "-std=c++11", "-w", "-I", os.path.join(ROOT, "include"), "-I", REF, os.path.join(REF, "main.cc"), *host_objects, "-o", exe,
                        "-lpthread", "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    far, near = synth_streams(1, fs * 4, fs, seed=99, delay_ms=50)
    outs = {}
    for tag, binary in (("ref", REF_BIN), ("model", exe)):
        d = tmp_path / tag
        d.mkdir()
        write_wav(d / "far.wav", far[0], fs)
        write_wav(d / "near.wav", near[0], fs)
        res = subprocess.run([binary, str(d / "far.wav"), str(d / "near.wav")], stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=600,
                             env=dict(ENV, ASAN_OPTIONS="detect_leaks=0"))  # main.cc itself never frees its sample buffers on some paths
        assert res.returncode == 0, res.stderr[-2000:]
        clean(res)
        outs[tag] = open(d / "near_out.wav", "rb").read()
    assert outs["ref"] == outs["model"] and outs["ref"][44:] != np.ascontiguousarray(near[0], dtype="<i2").tobytes()
