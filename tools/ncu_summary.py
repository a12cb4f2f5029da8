"""Summarise ncu outputs into small text files for profiles/.
  python tools/ncu_summary.py launches <launches.csv>         -> per-kernel totals and shares
  python tools/ncu_summary.py report <file.ncu-rep>           -> key metrics of the captured launch"""
import csv
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv:
            continue
        val = float(r[mv].replace(",", ""))
        unit = r[mu]
        us = val / 1e3 if unit in ("nsecond", "ns") else (val if unit in ("usecond", "us") else val * 1e3)
        name = r[kn].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':40s} {'launches':>9s} {'total us':>12s} {'share':>7s} {'avg us':>10s}")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} {c:9d} {us:12.1f} {100 * us / tot:6.1f}% {us / c:10.1f}")
    print(f"{'TOTAL':40s} {sum(v[0] for v in agg.values()):9d} {tot:12.1f}")


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel:", vals[idx["Kernel Name"]] if "Kernel Name" in idx else "?")
    for w in WANT:
        if w in idx:
            print(f"{w:80s} {vals[idx[w]]:>16s} {units[idx[w]]}")


def traffic(path, workload, factor_mode, launches_scale, out_json, how):
    """DRAM bytes (read + write) of the captured launch -> profiles/*traffic*.json, the file bench.py reads `roofline.traffic`
    from.  launches_scale: factor from the captured launch to all launches of that kernel in one pass."""
    import json
    import os

    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(name):
        v, u = float(vals[idx[name]].replace(",", "")), units[idx[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    r, w = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {"workload": workload, "factor_mode": factor_mode, "kernel": vals[idx["Kernel Name"]].split("(")[0],
           "dram_bytes_read_launch": r, "dram_bytes_write_launch": w, "launch_to_pass_scale": float(launches_scale),
           "dram_bytes_per_pass": (r + w) * float(launches_scale), "commit": commit, "ncu_report": os.path.basename(path), "how": how}
    recs = json.load(open(out_json)) if os.path.exists(out_json) else []
    recs = [x for x in recs if not (x["workload"] == workload and x["factor_mode"] == factor_mode)] + [rec]
    json.dump(recs, open(out_json, "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:8])
    else:
        {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
