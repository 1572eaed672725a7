#!/usr/bin/env python3
"""Summarise an Nsight Compute report (.ncu-rep, captured on the B200 box with --set full --import-source on)
into a small text file under profiles/: headline metrics per launch, aggregated warp-stall reasons, and the
executed warp-instructions per source file / hottest source lines.  Needs `ncu` on PATH (no GPU).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xyz.md --blocks-per-launch 12288,8192
"""
import argparse
import collections
import csv
import io
import subprocess

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max"]


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, *args, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--blocks-per-launch", default="", help="comma list: AECM blocks processed by each captured launch")
    ap.add_argument("--title", default="")
    a = ap.parse_args()
    blocks = [int(x) for x in a.blocks_per_launch.split(",") if x]
    lines = [f"# {a.title or a.report}", "", f"source report: `{a.report}` (ncu --set full --clock-control none --import-source on)", ""]
    raw = ncu_csv(a.report, "--page", "raw")
    hdr, units = raw[0], raw[1]
    for li, r in enumerate(raw[2:]):
        lines.append(f"## launch {r[hdr.index('ID')]}: {r[hdr.index('Kernel Name')]}")
        for k in KEYS:
            if k in hdr:
                lines.append(f"    {k:70s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        if li < len(blocks):
            inst = float(r[hdr.index("smsp__inst_executed.sum")])
            t_us = float(r[hdr.index("gpu__time_duration.sum")])
            dram = float(r[hdr.index("dram__bytes_read.sum")]) + float(r[hdr.index("dram__bytes_write.sum")])
            u = units[hdr.index("dram__bytes_read.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            dr = float(r[hdr.index("dram__bytes_read.sum")]) * scale.get(units[hdr.index("dram__bytes_read.sum")], 1)
            dw = float(r[hdr.index("dram__bytes_write.sum")]) * scale.get(units[hdr.index("dram__bytes_write.sum")], 1)
            lines.append(f"    -> {blocks[li]} AECM blocks: {inst / blocks[li]:.0f} warp-instructions/block, "
                         f"{blocks[li] / t_us:.1f} blocks/us under the profiler, DRAM traffic {dr + dw:.0f} B "
                         f"({(dr + dw) / blocks[li]:.0f} B/block)")
            del dram, u
        lines.append("")
    # source-level view
    src = ncu_csv(a.report, "--page", "source", "--print-source", "cuda,sass")
    cur = None
    hdr = None
    per = {}
    stalls = collections.Counter()
    for r in src:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
            scols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or not r:
            continue
        if r[0].isdigit() and len(r) > ie:
            try:
                v, s = int(r[ie]), int(r[sm])
            except ValueError:
                continue
            k = (cur, int(r[0]))
            o = per.get(k, (0, 0, ""))
            per[k] = (o[0] + v, o[1] + s, r[1].strip()[:90])
        elif r[0] == "" and len(r) > len(hdr) - 2:
            for i, h in scols:
                try:
                    stalls[h] += int(r[i])
                except ValueError:
                    pass
    tot = sum(v for v, _, _ in per.values()) or 1
    tots = sum(s for _, s, _ in per.values()) or 1
    lines.append("## warp stall reasons (all sampled launches)")
    st = sum(stalls.values()) or 1
    for h, v in stalls.most_common(10):
        lines.append(f"    {h:28s} {100 * v / st:5.1f} %")
    lines.append("")
    lines.append("## executed warp-instructions by source file")
    byfile = collections.Counter()
    for (f, _), (v, _, _) in per.items():
        byfile[f] += v
    for f, v in byfile.most_common(8):
        lines.append(f"    {f:28s} {100 * v / tot:5.1f} %")
    lines.append("")
    lines.append("## hottest source lines (share of executed warp-instructions | share of stall samples)")
    for (f, ln), (v, s, text) in sorted(per.items(), key=lambda kv: -kv[1][0])[:30]:
        lines.append(f"    {f[:16]:16s}:{ln:<5d} {100 * v / tot:5.1f} % | {100 * s / tots:5.1f} %   {text}")
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
