#!/usr/bin/env python
"""Opcode mix of a kernel from an ncu report (SASS view of the source page), weighted by executed warp-instructions.

usage: python tools/ncu_opmix.py gpurun_out/prof.ncu-rep [launch-index]
Groups opcodes by the issue pipe they use on sm_100 (fma: IMAD/FFMA..., alu: IADD3/LOP3/SHF/PRMT/ISETP/..., lsu,
xu, uniform, control) so that the pipe the kernel leans on can be read off directly."""
import collections
import csv
import io
import subprocess
import sys

PIPE = {
    "fma": ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IDP"),
    "alu": ("IADD", "IADD3", "LOP3", "SHF", "PRMT", "ISETP", "IMNMX", "VIMNMX", "VIMNMX3", "SEL", "FMNMX", "LEA", "MOV", "IABS", "FSETP", "PLOP3",
            "SGXT", "BMSK", "VIADD", "VIADDMNMX", "FSEL", "CS2R", "I2FP", "F2FP", "P2R", "R2P", "LOP"),
    "xu": ("MUFU", "POPC", "FLO", "I2F", "F2I", "BREV", "FRND"),
    "lsu": ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDSM", "ATOMS", "ATOMG", "RED", "SHFL", "LDC", "LDL", "STL", "MATCH", "VOTE", "REDUX", "UBLKCP",
            "SYNCS", "LDGSTS"),
    "uniform": ("S2UR", "UMOV", "ULOP3", "UIADD3", "USHF", "ULEA", "UISETP", "UIMAD", "R2UR", "LDCU", "USEL", "UPRMT", "UFLO", "UBREV", "UPOPC",
                "UPLOP3", "UP2UR", "UR2UP", "UMEMBAR", "UCGABAR_ARV", "UCGABAR_WAIT", "UVIADD", "UVIMNMX"),
    "control": ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "BAR", "NOP", "RET", "CALL", "BRX", "JMP", "YIELD", "NANOSLEEP", "ERRBAR", "MEMBAR",
                "FENCE", "DEPBAR", "S2R", "CCTL", "ELECT", "BREAK"),
}
OP2PIPE = {op: p for p, ops in PIPE.items() for op in ops}


def main():
    rep = sys.argv[1]
    want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    launches, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = []
            launches.append(cur)
        elif cur is not None:
            cur.append(line)
    rows = list(csv.reader(io.StringIO("\n".join(launches[want]))))
    hdr = rows[0]
    i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
    i_thr = hdr.index("Thread Instructions Executed")
    by_op, by_pipe = collections.Counter(), collections.Counter()
    total = thr = 0
    for r in rows[1:]:
        if len(r) <= i_ex:
            continue
        toks = r[i_src].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if not toks:
            continue
        op = toks[0].rstrip(";")
        base = op.split(".")[0]
        n = int(r[i_ex])
        total += n
        thr += int(r[i_thr])
        by_op[op if base in ("IMAD", "SHFL", "MUFU", "LDG", "LDS", "STS") else base] += n
        by_pipe[OP2PIPE.get(base, "other:" + base)] += n
    print(f"launch {want}: {total} warp-instructions, {thr / max(total, 1):.2f} threads/instruction")
    print("\n## by issue pipe")
    for p, n in by_pipe.most_common():
        print(f"    {p:<22}{100.0 * n / total:6.2f} %")
    print("\n## by opcode")
    for op, n in by_op.most_common(45):
        print(f"    {op:<22}{100.0 * n / total:6.2f} %")


if __name__ == "__main__":
    main()
