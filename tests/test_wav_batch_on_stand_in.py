"""aecm_wav_batch's host logic on the CPU, under sanitizers.

csrc/aecm_wav_batch.c (RIFF reader/writer, reader and writer threads, sharding over devices, lockstep chunks with instances
dropping out as their files end, the pipelined Submit/Wait loop, per-pair failure handling) is compiled unchanged against
tests/host_emu/fake_batch.c -- the CPU stand-in for the device batch, on the oracle -- once with AddressSanitizer +
UBSan and once with ThreadSanitizer, and must write every <near>_out.wav byte-identical to the reference CLI's
(oracle/_ref/aecm_run_ref: unmodified main.cc + reference library), the scenarios of tests/test_gpu_wav_batch.py.  The GPU
suite runs the shipped binary over the real batch; this one is about the tool's own memory and thread safety."""
import os
import subprocess

import numpy as np
import pytest

from test_gpu_wav_batch import make_pairs, reference_outputs
from test_wav_batch_io import REF_BIN, write_wav

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/aecm_run_ref not built (needs /root/reference)")

SAN = {"asan_ubsan": (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"], "ERROR: AddressSanitizer", "runtime error:"),
       "tsan": (["-fsanitize=thread"], "WARNING: ThreadSanitizer", "WARNING: ThreadSanitizer")}


# what lies below the tool: the stand-in batch (fake_batch.c), or the product's own csrc/aecm_batch.c on the modelled CUDA runtime
# (host_model.c, lazy execution) -- then the batch layer's staging, slots and events are under the sanitizer as well
BELOW = {"stand_in": [os.path.join(HERE, "host_emu", "fake_batch.c")],
         "host_model": [os.path.join(ROOT, "webrtc_aecm_b200", "csrc", "aecm_batch.c"), os.path.join(HERE, "host_emu", "host_model.c")]}


@pytest.fixture(scope="module", params=[(s_, b_) for s_ in sorted(SAN) for b_ in sorted(BELOW)], ids=lambda p: f"{p[0]}-{p[1]}")
def tool(request, tmp_path_factory):
    san, below = request.param
    flags, *marks = SAN[san]
    exe = str(tmp_path_factory.mktemp("wavtool") / f"aecm_wav_batch_{san}_{below}")
    cmd = ["gcc", *flags, "-O1", "-g", "-std=gnu11", "-I", os.path.join(HERE, "host_emu", "fake_cuda"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "webrtc_aecm_b200", "csrc"), os.path.join(ROOT, "webrtc_aecm_b200", "csrc", "aecm_wav_batch.c"), *BELOW[below],
           os.path.join(ROOT, "oracle", "aecm_oracle.c"), "-o", exe, "-lpthread", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "san" in r.stderr.lower() and "cannot find" in r.stderr.lower():
        pytest.skip("sanitizer runtime not available")
    assert r.returncode == 0, r.stderr
    return exe, marks


def run(tool, *args):
    exe, marks = tool
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", TSAN_OPTIONS="halt_on_error=0")
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=900, env=env)
    for m in marks:
        assert m not in r.stderr, r.stderr[-4000:]
    return r


@pytest.mark.parametrize("extra", [[], ["--devices", "0,0,0", "--chunk-calls", "37"]], ids=["one_shard", "three_shards_odd_chunks"])
def test_outputs_byte_identical_to_the_reference_cli(tool, tmp_path, extra):
    pairs = make_pairs(tmp_path, 9, seed=31)
    odd = tmp_path / "odd"  # an unsupported rate rides along: written back unchanged by both
    odd.mkdir()
    x = (np.arange(22050) * 13 % 5000 - 2500).astype(np.int16)
    write_wav(odd / "far.wav", x, 22050)
    write_wav(odd / "near.wav", x[::-1], 22050)
    pairs.append(odd)
    want = reference_outputs(pairs, tmp_path)
    args = []
    for d in pairs:
        args += [d / "far.wav", d / "near.wav"]
    res = run(tool, *extra, *args)
    assert res.returncode == 0, res.stderr + res.stdout
    for d in pairs:
        assert open(d / "near_out.wav", "rb").read() == want[d.name], d.name
    assert "10 pairs written" in res.stdout


def test_one_unreadable_pair_costs_only_its_own_output(tool, tmp_path):
    pairs = make_pairs(tmp_path, 4, seed=9)
    want = reference_outputs(pairs, tmp_path)
    (tmp_path / "junk_far.wav").write_bytes(b"this is not a RIFF file")
    (tmp_path / "junk_near.wav").write_bytes(b"RIFF\x00\x00\x00\x00nope")
    args = []
    for d in pairs[:2]:
        args += [d / "far.wav", d / "near.wav"]
    args += [tmp_path / "junk_far.wav", tmp_path / "junk_near.wav"]
    for d in pairs[2:]:
        args += [d / "far.wav", d / "near.wav"]
    res = run(tool, *args)
    assert res.returncode == 1, (res.stdout, res.stderr)
    assert "1 pair(s) unreadable" in res.stderr and "4 pairs written" in res.stdout
    for d in pairs:
        assert open(d / "near_out.wav", "rb").read() == want[d.name], d.name


@pytest.mark.parametrize("fail_in", [3, 24, 27, 30])
def test_a_shard_that_loses_an_allocation_costs_only_its_own_pairs(tool, tmp_path, fail_in):
    """Two shards on two (modelled) devices; the fail_in-th device or pinned allocation of the process fails -- while a batch is
    created, while a shard sets up its staging, or inside a submission.  The tool must say which shard lost its pairs, exit 1,
    still write every other pair byte-identical to the reference CLI, and release everything (leak check on under ASan)."""
    exe, _ = tool
    if "host_model" not in exe:
        pytest.skip("allocation failures are injected by the modelled CUDA runtime")
    pairs = make_pairs(tmp_path, 6, seed=13)
    want = reference_outputs(pairs, tmp_path)
    args = []
    for d in pairs:
        args += [d / "far.wav", d / "near.wav"]
    os.environ["FAKE_CUDA_FAIL_MALLOC_IN"] = str(fail_in)
    try:
        res = run(tool, "--devices", "0,1", *args)
    finally:
        del os.environ["FAKE_CUDA_FAIL_MALLOC_IN"]
    assert res.returncode == 1 and "get no output" in res.stderr, (res.stdout, res.stderr[-2000:])
    written = 0
    for d in pairs:
        if (d / "near_out.wav").exists():
            assert open(d / "near_out.wav", "rb").read() == want[d.name], d.name
            written += 1
    assert 0 < written < len(pairs)
