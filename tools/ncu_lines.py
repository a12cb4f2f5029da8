"""Aggregate ncu warp-stall samples per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [top=40]
(needs -lineinfo builds and `ncu --import-source on`)"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = defaultdict(lambda: [0, 0])
cur_line = None
for r in rows:
    if hdr is None:
        if len(r) > 4 and "Warp Stall Sampling (All Samples)" in r:
            hdr = r
            ci, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        continue
    if len(r) <= max(ci, ie):
        continue
    key = r[0]
    try:
        s, e = int(r[ci] or 0), int(r[ie] or 0)
    except ValueError:
        continue
    agg[(key, r[1][:120])][0] += s
    agg[(key, r[1][:120])][1] += e
tot = sum(v[0] for v in agg.values()) or 1
print("header cols:", hdr[:3] if hdr else None, "total samples", tot)
for (k, src), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{s:7d} {100 * s / tot:5.1f}%  inst {e:10d}  {k[-30:]:>30s} | {src.strip()}")
