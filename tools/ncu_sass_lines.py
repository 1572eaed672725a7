#!/usr/bin/env python3
"""Join an ncu report's per-SASS-instruction counters with the source lines of the same build (nvdisasm -g), so that
"which source lines produce the SHF / ISETP / BRA instructions" can be answered.  The report must come from the library
as given (same SASS).   usage: python tools/ncu_sass_lines.py prof.ncu-rep OPCODE-PREFIX|pipe:alu|pipe:fma [launch-index]
[--lib path/to/libaecm_b200.so-or-.o] [--src path/to/csrc]   (defaults: the in-tree library and sources; for an older capture
build that commit into a scratch directory: git archive <commit> webrtc_aecm_b200/csrc include | tar -x -C /tmp/old)"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disasm_lines(lib, kernel="aecm_call_kernel"):
    """[(opcode, file, line)] for the first kernel whose mangled name contains `kernel`, in program order."""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
        cubin = max((f for f in os.listdir(d) if "sm_100a" in f), key=lambda f: os.path.getsize(os.path.join(d, f)))
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, capture_output=True, text=True).stdout
    out, on, cur, seen = [], False, ("?", 0), False
    for line in txt.splitlines():
        if line.startswith("//---") and ".text." in line:
            on = kernel in line and not seen
            seen = seen or on
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if m:
            out.append((m.group(1), cur[0], cur[1]))
    return out


def main():
    argv = list(sys.argv[1:])
    lib, srcdir = os.path.join(ROOT, "webrtc_aecm_b200", "libaecm_b200.so"), os.path.join(ROOT, "webrtc_aecm_b200", "csrc")
    for flag in ("--lib", "--src"):
        if flag in argv:
            k = argv.index(flag)
            if flag == "--lib":
                lib = argv[k + 1]
            else:
                srcdir = argv[k + 1]
            del argv[k:k + 2]
    rep, prefix = argv[0], argv[1]
    want = int(argv[2]) if len(argv) > 2 else 0
    pipe = prefix[5:] if prefix.startswith("pipe:") else None
    if pipe:
        from ncu_opmix import OP2PIPE
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    launches, cur = [], None
    for line in txt.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = []
            launches.append(cur)
        elif cur is not None:
            cur.append(line)
    rows = list(csv.reader(io.StringIO("\n".join(launches[want]))))
    hdr = rows[0]
    i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
    prof = []
    for r in rows[1:]:
        if len(r) <= i_ex:
            continue
        toks = r[i_src].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if toks:
            prof.append((toks[0].rstrip(";"), int(r[i_ex])))
    dis = disasm_lines(lib)
    assert len(dis) == len(prof), (len(dis), len(prof), "report and library are different builds")
    assert sum(a[0] != b[0] for a, b in zip(dis, prof)) == 0, "opcode sequence differs: different builds"
    total = sum(n for _, n in prof)
    agg, flav = collections.Counter(), collections.Counter()
    for (op, f, ln), (_, n) in zip(dis, prof):
        if (OP2PIPE.get(op.split(".")[0]) == pipe) if pipe else op.startswith(prefix):
            agg[(f, ln)] += n
            flav[op] += n
    print(f"{prefix}*: {100.0 * sum(agg.values()) / total:.2f} % of {total} executed warp-instructions")
    for op, n in flav.most_common(8):
        print(f"    {op:<26}{100.0 * n / total:6.2f} %")
    src = {}
    for (f, ln), n in agg.most_common(32):
        if f not in src:
            path = os.path.join(srcdir, f)
            src[f] = open(path).read().splitlines() if os.path.exists(path) else []
        text = src[f][ln - 1].strip()[:96] if 0 < ln <= len(src[f]) else ""
        print(f"{100.0 * n / total:5.2f} %  {f}:{ln:<5} {text}")


if __name__ == "__main__":
    main()
