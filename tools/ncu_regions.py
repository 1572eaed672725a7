#!/usr/bin/env python3
"""Executed warp-instructions per source *section* of the device code (aecm_kernels.cu and its .cuh files) (sections = '// ----' banner comments
and function heads), from an ncu report captured with --import-source on.  Usage:
    python tools/ncu_regions.py gpurun_out/prof.ncu-rep 20480 [dir with the sources of the captured build]"""
import collections, csv, io, os, re, subprocess, sys
rep, nblocks = sys.argv[1], float(sys.argv[2])
src_dir = sys.argv[3] if len(sys.argv) > 3 else None  # the report carries line numbers: read the sources of the captured build
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; per = {}
for r in rows:
    if r and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); continue
    if hdr and r and r[0].isdigit() and len(r) > ie:
        try: v, s = int(r[ie]), int(r[sm])
        except ValueError: continue
        k = (cur, int(r[0])); o = per.get(k, (0, 0)); per[k] = (o[0] + v, o[1] + s)
tot = sum(v for v, _ in per.values()); tots = sum(s for _, s in per.values()) or 1
files = sorted({f for f, _ in per})
for f in files:
    try: src = open(os.path.join(src_dir, os.path.basename(f)) if src_dir else f).read().splitlines()
    except OSError: src = []
    marks = [(1, "(file head)")]
    for i, line in enumerate(src, 1):
        m = re.match(r"\s*// ---- (.*?)(?: ----.*)?$", line)
        if m: marks.append((i, m.group(1)[:60]))
        elif re.match(r"(__device__|__global__|template|static|inline).*\(", line) and not line.startswith(" "):
            name = re.search(r"(\w+)\s*\(", line)
            marks.append((i, "fn " + (name.group(1) if name else line[:40])))
    agg = collections.OrderedDict()
    for (ff, l), (v, s) in per.items():
        if ff != f: continue
        name = [n for (i, n) in marks if i <= l][-1]
        a = agg.get(name, [0, 0]); a[0] += v; a[1] += s; agg[name] = a
    fv = sum(a[0] for a in agg.values())
    print(f"== {f.split('/')[-1]}: {fv / nblocks:.0f} instr/block ({100 * fv / tot:.1f} %)")
    for name, (v, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if v / nblocks >= 5: print(f"   {v / nblocks:7.1f}  {100 * v / tot:5.1f}% | stall samples {100 * s / tots:5.1f}%  {name}")
print(f"total {tot / nblocks:.0f} instr/block")
