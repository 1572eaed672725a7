"""The product's host layer on a model of everything below it -- no GPU.

webrtc_aecm_b200/csrc/aecm_batch.c and csrc/echo_control_mobile.c are compiled UNCHANGED against tests/host_emu/host_model.c:
a model of the CUDA runtime in which streams are in-order queues executed as LATE as the API allows (plus an eager and a
random order), "device" memory is poisoned host memory, pinned and pageable host memory behave differently as they do on the
real runtime, and every kernel launch steps the CPU oracle.  The arithmetic is not under test here; the plumbing is: staging
planes, the four stream slots and their events, kernel order across slots, main-stream / slot-stream hand-over, index
checks, status delivery, the group API, the scalar combiner -- under an execution order that a real GPU would only pick on a
bad day.  A subset of the GPU suite's own tests runs against the model library (in a subprocess: the library is chosen at
import), followed by cases that need the model's hooks (tests/host_emu/model_cases.py) and by model-based fuzzing of the
batched ABI against the header's contract (tests/host_emu/model_fuzz.py).  Three mutants of aecm_batch.c with a
dependency removed must FAIL: the model is not vacuous.

TEST INFRASTRUCTURE: the model library is built into a temporary directory, announces itself on stderr when loaded, and is
not reachable from the product (which loads webrtc_aecm_b200/libaecm_b200.so and fails without a CUDA device)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "webrtc_aecm_b200", "csrc")

GPU_SUITE_SUBSET = ("fused_calls_equal or ragged or pipelined or instance_index or group_sharding or error_codes or scalar_api or "
                    "echo_path_round or deferred or live_config or pools_come or random_control")


def build(out_dir, batch_src=None, name="libaecm_b200_hostmodel.so", extra=()):
    lib = os.path.join(str(out_dir), name)
    cmd = ["gcc", "-O1", "-g", "-std=gnu11", "-fPIC", "-shared", *extra, "-I", os.path.join(HERE, "host_emu", "fake_cuda"), "-I", CSRC,
           "-I", os.path.join(ROOT, "include"), batch_src or os.path.join(CSRC, "aecm_batch.c"), os.path.join(CSRC, "echo_control_mobile.c"),
           os.path.join(HERE, "host_emu", "host_model.c"), os.path.join(ROOT, "oracle", "aecm_oracle.c"), "-o", lib, "-lpthread", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return lib


def run_pytest(lib, order, *args, seeds=()):
    env = dict(os.environ, AECM_B200_LIB=lib, FAKE_CUDA_ORDER=order, FAKE_CUDA_DEVICES="2", MODEL_FUZZ_SEEDS=",".join(map(str, seeds)) or "1,2,3,4,5")
    return subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", *args], capture_output=True, text=True,
                          env=env, cwd=ROOT, timeout=1500)


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    return build(tmp_path_factory.mktemp("hostmodel"))


@pytest.mark.parametrize("order", ["lazy", "eager", "random:1"])
def test_gpu_suite_subset_and_model_cases_pass_on_the_host_model(model, order):
    if order == "lazy":  # the GPU suite's own tests under the latest legal order (and, below, in random order under the sanitizers); the other
        # orders run the model's own cases and the fuzzer -- eager execution is what the real GPU does to those tests every round anyway
        r = run_pytest(model, order, "tests/test_gpu_parity.py", "tests/test_gpu_scalar_api.py", "-m", "gpu", "-k", GPU_SUITE_SUBSET)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1500:]
    r = run_pytest(model, order, "tests/host_emu/model_cases.py")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1500:]
    # model-based fuzzing of the batched ABI: random call sequences, valid and refused, against the header's contract
    r = run_pytest(model, order, "tests/host_emu/model_fuzz.py", seeds=range(1, 41))
    assert r.returncode == 0 and "40 passed" in r.stdout, r.stdout[-3000:] + r.stderr[-1500:]


def test_the_same_under_address_and_undefined_behaviour_sanitizers(tmp_path):
    """aecm_batch.c + echo_control_mobile.c + the model, built with -fsanitize=address,undefined and loaded into the test
    process's Python through LD_PRELOAD of the sanitizer runtime: staging-plane growth, slot reuse, index bitmaps, status copies
    and teardown run without a report (leak checking is off: the interpreter's own allocations would drown it)."""
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan.so not found")
    lib = build(tmp_path, name="libaecm_b200_hostmodel_asan.so", extra=("-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"))
    env = dict(os.environ, AECM_B200_LIB=lib, FAKE_CUDA_ORDER="random:3", FAKE_CUDA_DEVICES="2", LD_PRELOAD=libasan,
               ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    env["MODEL_FUZZ_SEEDS"] = ",".join(map(str, range(41, 81)))
    for args in (["tests/test_gpu_parity.py", "tests/test_gpu_scalar_api.py", "-m", "gpu", "-k", GPU_SUITE_SUBSET], ["tests/host_emu/model_cases.py"],
                 ["tests/host_emu/model_fuzz.py"]):
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", *args], capture_output=True, text=True, env=env,
                           cwd=ROOT, timeout=1500)
        assert r.returncode == 0 and "AddressSanitizer" not in r.stderr and "runtime error:" not in r.stderr, r.stdout[-2000:] + r.stderr[-3000:]


def test_every_allocation_failure_path_releases_what_it_held(tmp_path):
    """tests/host_emu/alloc_failure_main.c: each device / pinned allocation of AecmB200_Create, AecmB200_GroupCreate,
    WebRtcAecm_Create and of SubmitStream's on-demand staging is failed once in turn, under AddressSanitizer with leak checking."""
    exe = str(tmp_path / "alloc_failure")
    cmd = ["gcc", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-O1", "-g", "-std=gnu11", "-I", os.path.join(HERE, "host_emu", "fake_cuda"),
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, os.path.join(HERE, "host_emu", "alloc_failure_main.c"), os.path.join(CSRC, "aecm_batch.c"),
           os.path.join(CSRC, "echo_control_mobile.c"), os.path.join(HERE, "host_emu", "host_model.c"), os.path.join(ROOT, "oracle", "aecm_oracle.c"),
           "-o", exe, "-lpthread", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "asan" in r.stderr.lower() and "cannot find" in r.stderr.lower():
        pytest.skip("AddressSanitizer runtime not available")
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"), timeout=600)
    assert r.returncode == 0 and "Sanitizer" not in r.stderr and "runtime error:" not in r.stderr, r.stdout + r.stderr[-3000:]
    assert "allocation failures survived" in r.stdout


MUTANTS = {
    # slot streams no longer wait for work queued earlier on the batch's main stream (Init, set_config ...)
    "main_stream_handover": ("    if (b->main_dirty) {", "    if (0 && b->main_dirty) {", "lazy",
                             ["tests/test_gpu_parity.py", "-m", "gpu", "-k", "pipelined or ragged"]),
    # a submission's kernel no longer waits for the previous submission's kernel on another slot's stream
    "kernel_order_across_slots": ("    if (b->last_kernel) CU(b, cudaStreamWaitEvent(sl->copy, b->last_kernel, 0));", "    (void)0;", "lazy",
                                  ["tests/host_emu/model_cases.py", "-k", "newest_first"]),
    # main-stream work (here: a state read-back) no longer waits for the kernels of earlier submissions
    "main_stream_follows_slots": ("        CU(b, cudaStreamWaitEvent(b->stream, b->last_kernel, 0));", "        (void)0;", "lazy",
                                  ["tests/host_emu/model_cases.py", "-k", "main_stream_calls_see"]),
    # not a dependency but a check: a repeated instanceIndex is no longer refused (two rows would race on one record)
    "duplicate_index_accepted": ("rc = AECM_BAD_PARAMETER_ERROR; /* duplicate */", "rc = 0;", "eager", ["tests/host_emu/model_fuzz.py"]),
}


@pytest.mark.parametrize("name", sorted(MUTANTS))
def test_a_removed_dependency_is_caught(tmp_path, name):
    old, new, order, args = MUTANTS[name]
    src = open(os.path.join(CSRC, "aecm_batch.c")).read()
    assert src.count(old) == 1, "mutation site moved: update MUTANTS"
    mutated = tmp_path / "aecm_batch_mutant.c"
    mutated.write_text(src.replace(old, new))
    lib = build(tmp_path, str(mutated), name="libaecm_b200_hostmodel_mutant.so")
    r = run_pytest(lib, order, *args)
    assert r.returncode != 0 and ("failed" in r.stdout or "Aborted" in r.stderr or "error" in r.stdout.lower()), r.stdout[-2000:]
