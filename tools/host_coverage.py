#!/usr/bin/env python
"""gcov of the product's HOST layer under the CPU tests that drive it through the host model.

webrtc_aecm_b200/csrc/aecm_batch.c and echo_control_mobile.c are built with --coverage against tests/host_emu/host_model.c
(the modelled CUDA runtime with the oracle behind the launches: TEST INFRASTRUCTURE), the subset of the GPU suite that
tests/test_host_orchestration_model.py runs on the model, the model's own cases, the API fuzzer and the allocation-failure
driver are executed against it, and gcov reports which lines of the two files ran.   ->  profiles/host_code_coverage.md

    python tools/host_coverage.py
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
CSRC = os.path.join(ROOT, "webrtc_aecm_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "host_emu")
INC = ["-I", os.path.join(EMU, "fake_cuda"), "-I", CSRC, "-I", os.path.join(ROOT, "include")]
BELOW = [os.path.join(CSRC, "aecm_batch.c"), os.path.join(EMU, "host_model.c"), os.path.join(ROOT, "oracle", "aecm_oracle.c")]


def sh(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        sys.exit(f"{' '.join(map(str, cmd))}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
    return r


def gcov_report(workdir, source):
    """(executed, total, [(line number, text)] not executed) for `source` from the .gcda files in workdir."""
    base = os.path.basename(source)
    gcda = [f for f in os.listdir(workdir) if f.endswith(".gcda") and os.path.splitext(base)[0] in f]
    assert gcda, (base, os.listdir(workdir))
    sh(["gcov", "-o", workdir, *gcda], cwd=workdir)
    ran = total = 0
    missed = []
    for line in open(os.path.join(workdir, base + ".gcov")):
        m = re.match(r"\s*([^:]+):\s*(\d+):(.*)", line)
        if not m or m.group(1).strip() == "-":
            continue
        total += 1
        if m.group(1).strip().startswith(("#####", "=====")):
            missed.append((int(m.group(2)), m.group(3).rstrip()))
        else:
            ran += 1
    return ran, total, missed


def reason(src, text):
    """why a line cannot be reached by these tests"""
    t = text.strip()
    if src == "aecm_batch.c":
        if "AecmB200_Free(b)" in t or t == "return NULL;" or "GroupFree(g)" in t:
            return "<- event creation / first memset / calloc failing: the model only injects allocation failures"
        if "fail(b," in t:
            return "<- a copy failing with an error other than the refused two-allocation DMA"
    if src == "echo_control_mobile.c":
        if "keep" in t:
            return "<- two threads inside ONE instance at the same time: outside the reference's contract (not thread safe per instance)"
        return "<- AecmB200_Reset failing on a healthy batch"
    return ""


def main():
    import test_host_orchestration_model as t

    out = ["# Host-code coverage of the CPU tests that run the host layer on the model of the CUDA runtime (tools/host_coverage.py)", "",
           "gcov, line coverage; the device code has its own report (profiles/device_code_coverage.md).  What stays unexecuted is listed", "with the reason.", ""]
    with tempfile.TemporaryDirectory() as d:
        lib = os.path.join(d, "libaecm_b200_hostmodel_cov.so")
        sh(["gcc", "-O0", "-g", "--coverage", "-std=gnu11", "-fPIC", "-shared", *INC, os.path.join(CSRC, "echo_control_mobile.c"), *BELOW, "-o", lib,
            "-lpthread", "-lm"], cwd=d)
        env = dict(os.environ, AECM_B200_LIB=lib, FAKE_CUDA_DEVICES="2", MODEL_FUZZ_SEEDS=",".join(map(str, range(1, 21))))
        for order in ("lazy", "random:1"):
            e = dict(env, FAKE_CUDA_ORDER=order)
            sh([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "tests/test_gpu_parity.py", "tests/test_gpu_scalar_api.py", "-m", "gpu", "-k",
                t.GPU_SUITE_SUBSET + " or device_fft or delay_estimator_options"], cwd=ROOT, env=e)
            sh([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "tests/host_emu/model_cases.py", "tests/host_emu/model_fuzz.py"], cwd=ROOT, env=e)
        exe = os.path.join(d, "alloc_failure")
        sh(["gcc", "-O0", "-g", "--coverage", "-std=gnu11", *INC, os.path.join(EMU, "alloc_failure_main.c"), os.path.join(CSRC, "echo_control_mobile.c"), *BELOW,
            "-o", exe, "-lpthread", "-lm"], cwd=d)
        sh([exe], cwd=d)
        for src in ("aecm_batch.c", "echo_control_mobile.c"):
            # two builds contributed (.so and the driver): gcov merges nothing across object names, so report the better of the two per line
            reports = []
            for tag in ("libaecm_b200_hostmodel_cov.so-", "alloc_failure-"):
                sub = os.path.join(d, "gc_" + tag.strip("-"))
                os.makedirs(sub, exist_ok=True)
                for f in os.listdir(d):
                    if f.startswith(tag + os.path.splitext(src)[0]) and f.endswith((".gcda", ".gcno")):
                        os.replace(os.path.join(d, f), os.path.join(sub, f))
                reports.append(gcov_report(sub, os.path.join(CSRC, src)))
            total = reports[0][1]
            missed = sorted(set(reports[0][2]) & set(reports[1][2]))
            out += [f"## `csrc/{src}`: {total - len(missed)} of {total} lines ({100.0 * (total - len(missed)) / total:.1f} %)", ""]
            out += [f"    {n:4d}: {text.strip()[:72]:<74}{reason(src, text)}" for n, text in missed] + [""]
    open(os.path.join(ROOT, "profiles", "host_code_coverage.md"), "w").write("\n".join(out))
    print("\n".join(out))


if __name__ == "__main__":
    main()
