"""The scalar-API layer's host logic under ThreadSanitizer, without a GPU.

webrtc_aecm_b200/csrc/echo_control_mobile.c (handle pools, the request combiner with its leader election and hand-over, the
deferred far end, pool teardown) is compiled unchanged with -fsanitize=thread against tests/host_emu/fake_batch.c, a CPU
stand-in for the device batch built on the oracle, and driven by tests/host_emu/scalar_tsan_main.c from 8 threads: every
handle follows a script (underruns, overruns, live set_config, echo-path import/export, both frame sizes, clean plane or
not), the expected output of every call comes from the oracle stepped directly.  The GPU suite drives the same layer over
the real batch (tests/test_gpu_scalar_api.py); what it cannot see -- a data race that happens not to corrupt anything on
the day -- is what this test is for."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LAYER = os.path.join(ROOT, "webrtc_aecm_b200", "csrc", "echo_control_mobile.c")


def build(tmp_path, layer_src, name, backend="stand_in"):
    """backend "stand_in": the layer over tests/host_emu/fake_batch.c; "host_model": over the product's own csrc/aecm_batch.c on
    the modelled CUDA runtime (tests/host_emu/host_model.c), so that the batch layer is under the sanitizer as well."""
    exe = str(tmp_path / name)
    below = [os.path.join(HERE, "host_emu", "fake_batch.c")] if backend == "stand_in" else [
        os.path.join(ROOT, "webrtc_aecm_b200", "csrc", "aecm_batch.c"), os.path.join(HERE, "host_emu", "host_model.c")]
    cmd = ["gcc", "-fsanitize=thread", "-O1", "-g", "-std=gnu11", "-I", os.path.join(HERE, "host_emu", "fake_cuda"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "webrtc_aecm_b200", "csrc"), os.path.join(HERE, "host_emu", "scalar_tsan_main.c"),
           *below, layer_src, os.path.join(ROOT, "oracle", "aecm_oracle.c"), "-o", exe, "-lpthread", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "tsan" in r.stderr.lower():
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr.strip().splitlines()[-1])
    assert r.returncode == 0, r.stderr
    return exe


def run(exe, *args):
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0")
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, env=env, timeout=900)
    return r, r.stderr.count("WARNING: ThreadSanitizer")


@pytest.mark.parametrize("backend", ["stand_in", "host_model"])
def test_scalar_layer_is_race_free_and_exact_under_threads(tmp_path, backend):
    exe = build(tmp_path, LAYER, "scalar_tsan", backend)
    # 8 threads x 40 handles = 320 handles: two pools; 50 calls each; then 30 rounds of create/use/free churn per thread
    r, warnings = run(exe, 8, 40, 50, 30)
    assert r.returncode == 0 and warnings == 0, r.stdout[-2000:] + r.stderr[-4000:]
    m = re.search(r"(\d+) rounds, (\d+) requests", r.stdout)
    assert m and int(m.group(2)) > 1.5 * int(m.group(1)), "the combiner did not combine: " + r.stdout
    if backend == "stand_in":
        m = re.search(r"device batches created (\d+), freed (\d+)", r.stdout)
        assert m and int(m.group(1)) >= 2 and int(m.group(2)) == int(m.group(1)) - 1, "pool teardown: all but the last pool are released: " + r.stdout


def test_the_harness_notices_a_missing_lock(tmp_path):
    """Negative control: the same layer with the batch lock taken out of the combiner's round is reported."""
    src = open(LAYER).read()
    i = src.index("static void run_round(")
    j = src.index("/* Queue the request and wait for its result", i)
    body = src[i:j].replace("pthread_mutex_lock(&p->exec);", "").replace("pthread_mutex_unlock(&p->exec);", "")
    assert body != src[i:j]
    mutated = tmp_path / "echo_control_mobile_nolock.c"
    mutated.write_text(src[:i] + body + src[j:])
    exe = build(tmp_path, str(mutated), "scalar_tsan_nolock")
    r, warnings = run(exe, 8, 6, 80, 0)
    assert warnings > 0 or r.returncode != 0
