"""Aggregate ncu per-source-line metrics (needs -lineinfo builds and `ncu --import-source on`).
usage: python tools/ncu_lines.py report.ncu-rep [top=40] [sort=samples|inst]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "samples"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, data = "?", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        ci, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= max(ci, ie) or r[0] == "":
        continue
    try:
        data.append((int(r[ci] or 0), int(r[ie] or 0), cur_file, r[0], r[1].strip()))
    except ValueError:
        pass
ts, ti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
print(f"total samples {ts}, total warp instructions {ti}")
data.sort(key=lambda d: -(d[0] if key == "samples" else d[1]))
for s, e, f, ln, src in data[:top]:
    print(f"{s:7d} {100 * s / ts:5.1f}%  inst {e:11d} {100 * e / ti:5.1f}%  {f}:{ln:>4s} | {src[:105]}")
