"""The C-ABI library loads and exports every symbol the public headers declare (no compute, no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:WebRtcAecm|AecmB200)_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    from webrtc_aecm_b200 import api

    lib = ctypes.CDLL(api.library_path())
    names = declared_functions("echo_control_mobile.h") + declared_functions("aecm_b200.h")
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_reference_symbol_set_is_complete():
    # the ten entry points of the reference's aecm/echo_control_mobile.h:46-202
    want = {"WebRtcAecm_Create", "WebRtcAecm_Free", "WebRtcAecm_Init", "WebRtcAecm_BufferFarend",
            "WebRtcAecm_GetBufferFarendError", "WebRtcAecm_Process", "WebRtcAecm_set_config",
            "WebRtcAecm_InitEchoPath", "WebRtcAecm_GetEchoPath", "WebRtcAecm_echo_path_size_bytes"}
    assert want == set(declared_functions("echo_control_mobile.h"))


def test_state_mirror_and_constants():
    from webrtc_aecm_b200 import api

    L = api.lib()  # asserts sizeof(InstanceState) == AecmB200_StateBytes()
    assert L.WebRtcAecm_echo_path_size_bytes() == 130
    assert L.AecmB200_StateBytes() % 16 == 0


def test_host_side_argument_errors_need_no_gpu():
    from webrtc_aecm_b200 import api

    L = api.lib()
    # NULL instance / batch -> -1, as in the reference (echo_control_mobile.cc:146-148, 200-201, 252-254)
    assert L.WebRtcAecm_Init(None, 16000) == -1
    assert L.WebRtcAecm_BufferFarend(None, None, 160) == -1
    assert L.WebRtcAecm_Process(None, None, None, None, 160, 0) == -1
    assert L.WebRtcAecm_set_config(None, api.AecmConfig(1, 3)) == -1
    assert L.WebRtcAecm_GetEchoPath(None, None, 130) == -1
    assert L.AecmB200_Init(None, 0, 1, 16000) == -1
    assert L.AecmB200_ProcessStream(None, None) == -1
    L.WebRtcAecm_Free(None)  # no-op, echo_control_mobile.cc:122-124
    L.AecmB200_Free(None)


def test_product_does_not_reference_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "webrtc_aecm_b200")):
        for f in files:
            if f.endswith((".c", ".cu", ".h", ".cuh", ".py", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle/|aecm_oracle|oracle_lib|libaecm_ref", text):
                    bad.append(os.path.join(dirpath, f))
    for f in ("include/aecm_b200.h", "include/echo_control_mobile.h"):
        if re.search(r"oracle", open(os.path.join(ROOT, f)).read()):
            bad.append(f)
    assert not bad, bad


def test_built_kernel_keeps_its_launch_shape_and_hardware_paths():
    """Static checks on the built sm_100a code (cuobjdump, no GPU): the hot kernel still fits 28 warps per SM (<= 72
    registers, nothing spilled to local memory), moves the instance record with the bulk-copy (TMA) unit both ways
    (UBLKCP + mbarrier SYNCS), uses the hardware square-root estimate and warp-wide redux -- so an edit that silently
    costs the residency or falls back to a slower path fails here before it reaches a GPU."""
    import re
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "webrtc_aecm_b200", "libaecm_b200.so")
    res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    # two instantiations: <false> is the hot path (AECM as shipped), <true> additionally drives the delay estimator's
    # optional machinery and may keep a few words on the stack
    usage = {m.group(1): tuple(map(int, m.groups()[1:])) for m in
             re.finditer(r"aecm_call_kernelILb([01])E\S*:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:\d+ LOCAL:(\d+)", res)}
    assert set(usage) == {"0", "1"}, res[:600]
    regs, stack, local = usage["0"]
    assert regs <= 72 and stack == 0 and local == 0, usage
    assert usage["1"][0] <= 72 and usage["1"][1] <= 64 and usage["1"][2] == 0, usage
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "arch = sm_100a" in sass
    funcs = re.split(r"\n\s*Function : ", sass)
    hot = [f for f in funcs if "aecm_call_kernelILb0E" in f.split("\n", 1)[0]]
    assert len(hot) == 1
    call = hot[0]
    for mnemonic in ("UBLKCP.S.G", "UBLKCP.G.S", "SYNCS.PHASECHK", "MUFU.SQRT", "CREDUX", "BAR.SYNC"):
        assert mnemonic in call, mnemonic
    assert not re.search(r"\b(LDL|STL)\b", call)


def test_delay_estimator_example_builds(tmp_path):
    """examples/delay_estimator_options.c (SURVEY 8-f4 from plain C) compiles warning-free against include/; it runs in the
    GPU suite (tests/test_gpu_main_cc.py)."""
    import subprocess

    exe = str(tmp_path / "deo")
    lib_dir = os.path.join(ROOT, "webrtc_aecm_b200")
    r = subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "delay_estimator_options.c"), "-L", lib_dir, "-laecm_b200", f"-Wl,-rpath,{lib_dir}",
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_example_program_builds_against_the_public_headers(tmp_path):
    """examples/batch_stream.c (the INTEGRATION.md throughput loop as a complete program) compiles warning-free against
    include/ and links against the library; without a GPU it must stop with the library's own refusal."""
    import subprocess

    exe = str(tmp_path / "batch_stream")
    lib_dir = os.path.join(ROOT, "webrtc_aecm_b200")
    r = subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "batch_stream.c"), "-L", lib_dir, "-laecm_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch

    if not torch.cuda.is_available():
        run = subprocess.run([exe, "8", "2"], capture_output=True, text=True, timeout=60)
        assert run.returncode == 1 and "no CPU path" in run.stderr
