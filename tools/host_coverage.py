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


def wav_tool_coverage():
    """csrc/aecm_wav_batch.c over csrc/aecm_batch.c on the model: the scenarios of tests/test_wav_batch_on_stand_in.py and
    tests/test_wav_batch_io.py plus the command-line corners."""
    import pathlib

    import numpy as np
    from test_gpu_wav_batch import make_pairs
    from test_wav_batch_io import write_wav

    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "aecm_wav_batch_cov")
        sh(["gcc", "-O0", "-g", "--coverage", "-std=gnu11", *INC, os.path.join(CSRC, "aecm_wav_batch.c"), *BELOW, "-o", exe, "-lpthread", "-lm"], cwd=d)
        root = pathlib.Path(d) / "audio"
        root.mkdir()
        pairs = make_pairs(root, 7, seed=31)
        odd = root / "odd"
        odd.mkdir()
        x = (np.arange(22050) * 13 % 5000 - 2500).astype(np.int16)
        write_wav(odd / "far.wav", x, 22050)
        write_wav(odd / "near.wav", x[::-1], 22050)
        for width in (1, 3, 4):  # other sample widths, stereo
            write_wav(root / f"w{width}_far.wav", x[:16000], 16000, width)
            write_wav(root / f"w{width}_near.wav", x[:16000], 16000, width)
        write_wav(root / "st_far.wav", np.repeat(x[:16000], 2), 16000, 2, channels=2)
        write_wav(root / "st_near.wav", np.repeat(x[:16000], 2), 16000, 2, channels=2)
        (root / "junk_far.wav").write_bytes(b"this is not a RIFF file")
        (root / "junk_near.wav").write_bytes(b"RIFF\x00\x00\x00\x00nope")
        args = [str(p / f) for p in pairs + [odd] for f in ("far.wav", "near.wav")]
        extra = [str(root / f"{n}_{k}.wav") for n in ("w1", "w3", "w4", "st") for k in ("far", "near")]
        lst = root / "more.txt"
        lst.write_text(f"{root / 'w3_far.wav'} {root / 'w3_near.wav'} {root / 'w3_listed_out.wav'}\n# a comment\n\n")
        runs = [[*args], ["--devices", "0,0,0", "--chunk-calls", "37", "-q", *args], ["--gpus", "2", "--io-threads", "2", *args, *extra],
                ["--echo-mode", "3", "--ms", "20", "--cng", "0", "--suffix", "_x", *args[:4]], ["--passthrough", *args[:4], *extra],
                [*args[:4], str(root / "junk_far.wav"), str(root / "junk_near.wav")], ["--list", str(lst), *args[:2]], ["--list", str(root / "missing.txt")],
                [str(root / "nope_far.wav"), str(root / "nope_near.wav")], [], ["-h"], ["--bogus"], ["--echo-mode", "9", *args[:2]], ["--gpus", "0", *args[:2]],
                ["--devices", ",", *args[:2]], [args[0]]]
        # damaged and mismatched files
        good = open(pairs[0] / "near.wav", "rb").read()
        (root / "trunc_near.wav").write_bytes(good[:44 + 1000])            # the header promises more samples than the file holds
        (root / "nodata_near.wav").write_bytes(good[:36] + b"LIST" + (4).to_bytes(4, "little") + b"abcd")  # no data chunk
        write_wav(root / "short_far.wav", x[:8000], 16000)                 # a far file shorter than its near file
        write_wav(root / "long_near.wav", x[:16000], 16000)
        lst2 = root / "bad_out.txt"
        lst2.write_text(f"{pairs[1] / 'far.wav'} {pairs[1] / 'near.wav'} {root / 'no_such_dir' / 'out.wav'}\n")
        runs += [[str(pairs[0] / "far.wav"), str(root / "trunc_near.wav")], [str(pairs[0] / "far.wav"), str(root / "nodata_near.wav")],
                 [str(root / "st_far.wav"), str(root / "w4_near.wav")], [str(root / "junk_far.wav"), str(root / "w4_near.wav")],
                 [str(root / "short_far.wav"), str(root / "long_near.wav")], ["--list", str(lst2)]]
        for r in runs:
            subprocess.run([exe, *r], capture_output=True, text=True, cwd=d, timeout=600, env=dict(os.environ, FAKE_CUDA_DEVICES="2"))
        for k in (3, *range(22, 34)):  # a device or pinned allocation fails while a shard is being set up / while it runs: its pairs get no output, the tool says so
            subprocess.run([exe, "--devices", "0,1", *args], capture_output=True, text=True, cwd=d, timeout=600,
                           env=dict(os.environ, FAKE_CUDA_DEVICES="2", FAKE_CUDA_FAIL_MALLOC_IN=str(k)))
        ran, total, missed = gcov_report(d, os.path.join(CSRC, "aecm_wav_batch.c"))
    return [f"## `csrc/aecm_wav_batch.c`: {ran} of {total} lines ({100.0 * ran / total:.1f} %)", ""] + [
        f"    {n:4d}: {text.strip()[:72]:<74}<- an I/O error in the middle of a file that opened and sized correctly" for n, text in missed] + [""]


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
    out += wav_tool_coverage()
    open(os.path.join(ROOT, "profiles", "host_code_coverage.md"), "w").write("\n".join(out))
    print("\n".join(out))


if __name__ == "__main__":
    main()
