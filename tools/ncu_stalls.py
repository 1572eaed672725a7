#!/usr/bin/env python
"""Source lines ranked by one stall reason (or any per-instruction column) of an ncu report taken with
--import-source on.   usage: python tools/ncu_stalls.py prof.ncu-rep stall_long_sb [top-n]"""
import collections
import csv
import io
import subprocess
import sys

rep, col = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = hdr = None
agg, text, sass = collections.Counter(), {}, collections.defaultdict(collections.Counter)
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr, ic = r, r.index(col)
    elif hdr and r and r[0].isdigit() and len(r) > ic:
        try:
            v = int(r[ic])
        except ValueError:
            continue
        k = (cur, int(r[0]))
        agg[k] += v
        text[k] = r[1]
        if v:
            sass[k][r[3].split()[0] if not r[3].strip().startswith("@") else r[3].split()[1]] += v
tot = sum(agg.values()) or 1
print(f"{col}: {tot} samples")
for k, v in agg.most_common(top):
    ops = " ".join(f"{o}:{n}" for o, n in sass[k].most_common(3))
    print(f"{100 * v / tot:5.1f}%  {k[0]}:{k[1]:<5} {text[k].strip()[:90]}   [{ops}]")
