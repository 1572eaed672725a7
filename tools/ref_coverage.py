#!/usr/bin/env python3
"""Which lines of the reference's hot path do the parity cases execute?

Builds a gcov-instrumented copy of the UNMODIFIED reference library in a scratch directory (sources read where they
lie under /root/reference; nothing is copied into the repo), drives every case of tests/cases.py through it with the
oracle's own runner, and reports per file the executed-line share and the lines never reached.  Writes
oracle/COVERAGE.md.  CPU only; needs /root/reference (so: this container, not the GPU box).

    python tools/ref_coverage.py [case-name ...]"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("AECM_REFERENCE", "/root/reference")
SRC_C = ["complex_fft.c", "real_fft.c", "ring_buffer.c"]
SRC_CC = ["aecm_core.cc", "aecm_core_c.cc", "delay_estimator.cc", "delay_estimator_wrapper.cc", "echo_control_mobile.cc",
          "signal_processing_library.cc"]


# why a line is never reached: (file, first line, last line) -> reason
WHY = [
    ("aecm_core.cc", 186, 230, "allocation-failure unwinding in WebRtcAecm_CreateCore"),
    ("aecm_core.cc", 364, 392, "InitCore argument / delay-estimator init failures; the wrapper validates the rate first"),
    ("aecm_core.cc", 477, 486, "WebRtcAecm_Control (nlpFlag, fixedDelay) has no caller; FreeCore(NULL)"),
    ("aecm_core.cc", 545, 550, "ProcessBlock failure propagation (delay-estimator handle errors only)"),
    ("aecm_core.cc", 779, 779, "mu = MU_MIN needs currentVAD with farEnergyMin >= farEnergyMax after start-up: the VAD can only be "
                               "set while max - min > FAR_ENERGY_DIFF once start-up is over, and max overtakes min within a block of "
                               "any level step that sets it earlier (analysed in DESIGN.md); the device code keeps the branch"),
    ("aecm_core.cc", 1087, 1087, "negative farBuf read position: the far history has no negative delays (lookahead 0)"),
    ("aecm_core_c.cc", 473, 494, "delay-estimator errors and the fixedDelay override (never set)"),
    ("aecm_core_c.cc", 608, 608, "hnl < 0 after 1 - x with x in [0, 2^14]: arithmetically impossible"),
    ("aecm_core_c.cc", 658, 658, "hnl > ONE_Q14 clamp: hnl <= ONE_Q14 by construction"),
    ("aecm_core_c.cc", 689, 698, "nlpFlag == 0 branch: nlpFlag is 1 forever (aecm_core.cc:399)"),
    ("echo_control_mobile.cc", 95, 157, "allocation failures, Free(NULL), NULL instance, bad rate: argument checks (exercised against "
                                        "the product by tests/test_gpu_scalar_api.py, not by the stream runner)"),
    ("echo_control_mobile.cc", 187, 269, "argument / state checks of Init, BufferFarend, Process (see tests/test_gpu_scalar_api.py)"),
    ("echo_control_mobile.cc", 396, 427, "core failure propagation; set_config argument checks (tests/test_gpu_scalar_api.py)"),
    ("echo_control_mobile.cc", 481, 531, "InitEchoPath / GetEchoPath / echo_path_size_bytes: covered by tests/test_oracle.py and "
                                         "tests/test_gpu_scalar_api.py, not called by the stream runner"),
    ("echo_control_mobile.cc", 548, 548, "first line of a two-line statement (gcov attributes it to the second line)"),
    ("ring_buffer.c", 56, 106, "allocation failures and NULL-argument checks"),
    ("ring_buffer.c", 128, 128, "ReadBuffer without a data_ptr: every AECM caller passes one"),
    ("ring_buffer.c", 146, 227, "NULL-argument checks; reading more than is buffered (callers check availability first)"),
    ("delay_estimator.cc", 0, 10 ** 6, "allocation failures / NULL checks / history-size mismatch returns, soft reset by 0; "
                                       "WebRtc_binary_last_delay_quality is not called by the stream runner: it is pinned by "
                                       "tests/test_oracle.py::test_last_delay_quality_oracle_equals_reference, which drives the reference call "
                                       "by call, and on the product by AecmB200_GetDelayEstimate in tests/test_gpu_parity.py"),
    ("delay_estimator_wrapper.cc", 0, 10 ** 6, "float spectrum entry points (AECM uses the fixed-point ones), WebRtc_set_history_size (AECM's "
                                               "far history is a fixed 100 rows), getters, allocation failures and argument checks"),
    ("signal_processing_library.cc", 0, 10 ** 6, "helpers AECM never calls (other SPL functions kept in the file)"),
    ("complex_fft.c", 0, 10 ** 6, "mode-0 (low-accuracy) branches: AECM always passes mode 1; allocation-free error returns"),
    ("real_fft.c", 0, 10 ** 6, "order-out-of-range returns"),
]


def why(f, ln):
    for wf, a, b, text in WHY:
        if wf == f and a <= ln <= b:
            return text
    return "REACHABLE? -- add a case"


def main():
    import cases
    import oracle_lib as ol

    names = sys.argv[1:] or [c["name"] for c in cases.CASES]
    with tempfile.TemporaryDirectory() as d:
        objs = []
        for f in SRC_C + SRC_CC:
            cc = ["gcc"] if f.endswith(".c") else ["g++", "-std=c++11"]
            o = os.path.join(d, f + ".o")
            subprocess.run(cc + ["-O0", "-g", "--coverage", "-fPIC", "-c", os.path.join(REF, "aecm", f), "-o", o], check=True, cwd=d)
            objs.append(o)
        so = os.path.join(d, "libaecm_ref_cov.so")
        subprocess.run(["g++", "-shared", "--coverage", "-o", so] + objs, check=True)
        # the runner dlopens the library by path; a child process per batch of cases so that the .gcda files get flushed at exit
        code = (
            "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import cases, oracle_lib as ol\n"
            "ol.REF_SO = %r\n"
            "for name in %r:\n"
            "    c = cases.CASE_BY_NAME[name]\n"
            "    far, near, clean, ms = cases.make_inputs(c)\n"
            "    ol.run_streams(far, near, impl='ref', clean=clean, ms=ms, threads=1, **cases.run_kwargs(c))\n"
        ) % (ROOT, os.path.join(ROOT, "tests"), so, names)
        subprocess.run([sys.executable, "-c", code], check=True)
        # the delay estimator's optional machinery, through the shim (oracle/ref_shim.cc includes the reference's
        # echo_control_mobile.cc and calls the estimator's own setters); the other objects are shared, so their counters add up
        if not sys.argv[1:]:
            shim_o = os.path.join(d, "ref_shim.o")
            subprocess.run(["g++", "-std=c++11", "-O0", "-g", "--coverage", "-fPIC", "-I", REF, "-I", os.path.join(REF, "aecm"), "-c",
                            os.path.join(ROOT, "oracle", "ref_shim.cc"), "-o", shim_o], check=True, cwd=d)
            shim_so = os.path.join(d, "libaecm_ref_shim_cov.so")
            subprocess.run(["g++", "-shared", "--coverage", "-o", shim_so, shim_o] +
                           [o for o in objs if not o.endswith("echo_control_mobile.cc.o")], check=True)
            code = (
                "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
                "import cases, oracle_lib as ol\n"
                "ol.REF_SHIM_SO = %r\n"
                "for cfg in cases.DLY_CONFIGS:\n"
                "    fs, spc, far, near, ms, kw = cases.dly_inputs(cfg)\n"
                "    ol.run_streams(far, near, impl='ref', ms=ms, threads=1, dly=cfg['dly'], soft_reset=cfg['soft_reset'], delay_trace=True, **kw)\n"
            ) % (ROOT, os.path.join(ROOT, "tests"), shim_so)
            subprocess.run([sys.executable, "-c", code], check=True)
        report = []
        for f in SRC_CC + SRC_C:
            subprocess.run(["gcov", "-o", d, os.path.join(d, f + ".o")], cwd=d, capture_output=True, text=True)
            g = os.path.join(d, f + ".gcov")
            if not os.path.exists(g):
                continue
            hit = miss = 0
            missed = []
            for line in open(g, errors="replace"):
                m = re.match(r"\s*([0-9#=\-*]+):\s*(\d+):(.*)", line)
                if not m or m.group(1) == "-":
                    continue
                if m.group(1).startswith("#") or m.group(1).startswith("="):
                    miss += 1
                    missed.append((int(m.group(2)), m.group(3).rstrip()))
                else:
                    hit += 1
            report.append((f, hit, miss, missed))
    out = ["# Reference lines executed by the parity cases", "",
           "Generated by `python tools/ref_coverage.py` (gcov on an -O0 build of the unmodified reference, all cases of",
           "`tests/cases.py` through the oracle's runner, plus the eight delay-estimator configurations `DLY_CONFIGS` -- robust",
           "validation, lookahead, soft resets, SURVEY 8-f4 -- through `oracle/ref_shim.cc`).  \"never reached\" lists source lines of the reference by number;",
           "the annotations say why a line is dead for every caller or which case family would be needed.", ""]
    out.append("| file | executed | never reached | share |")
    out.append("|---|---|---|---|")
    for f, hit, miss, _ in report:
        out.append(f"| `aecm/{f}` | {hit} | {miss} | {100.0 * hit / max(hit + miss, 1):.1f} % |")
    out.append("")
    for f, hit, miss, missed in report:
        if not missed:
            continue
        out.append(f"## aecm/{f}: lines never reached")
        last = None
        for ln, text in missed:
            reason = why(f, ln)
            if reason != last:
                if last is not None:
                    out.append("```")
                out.append("")
                out.append(f"*{reason}*")
                out.append("```")
                last = reason
            out.append(f"{ln:5d}: {text.strip()[:110]}")
        out.append("```")
        out.append("")
    path = os.path.join(ROOT, "oracle", "COVERAGE.md")
    open(path, "w").write("\n".join(out))
    open_items = [(f, ln) for f, _, _, missed in report for ln, _ in missed if why(f, ln).startswith("REACHABLE")]
    print("lines without an explanation:", open_items or "none")
    for f, hit, miss, _ in report:
        print(f"{f:32s} {hit:5d} executed {miss:4d} never reached  {100.0 * hit / max(hit + miss, 1):5.1f} %")
    print("wrote", path)


if __name__ == "__main__":
    main()
