#!/usr/bin/env python
"""Summarise ncu output for profiles/.

    tools/ncu_summary.py launches <launches.csv> <out.md>      # per-scan launch list -> kernel shares
    tools/ncu_summary.py full <report.ncu-rep> <out.json>      # --set full capture -> the metrics the roofline cites

Runs here (no GPU needed): `ncu -i` only reads the report."""
import collections
import csv
import json
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("dlt::", "")


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    return float(v) * f if f else float(v)


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    names = [short(r[4]) for r in rows]
    ns = [float(r[-1].replace(",", "")) for r in rows]
    starts = [i for i, n in enumerate(names) if n == "k_scan_first"]
    lines = ["# ncu launch list, per scan (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)", "",
             f"source: `{path}`; {len(rows)} launches, {len(starts)} scans; the last scan is summarised, shares are of the scan's own kernels", ""]
    if not starts:
        raise SystemExit("no k_scan_first in the list")
    s, e = starts[-1], len(rows)
    agg = collections.OrderedDict()
    for n, x in zip(names[s:e], ns[s:e]):
        if n.startswith("at::"):
            continue  # torch's L2-flush memset between scans
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += x
    tot = sum(v[1] for v in agg.values())
    lines += ["| kernel | launches | total us | share |", "|---|---|---|---|"]
    for n, (c, x) in agg.items():
        lines.append(f"| `{n}` | {c} | {x / 1e3:.2f} | {100 * x / tot:.1f} % |")
    lines.append(f"| **sum** | {sum(v[0] for v in agg.values())} | {tot / 1e3:.2f} | 100 % |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[idx["Kernel Name"]])}
        for m in FULL_METRICS:
            if m in idx:
                v, u = r[idx[m]].replace(",", ""), units[idx[m]]
                try:
                    d[m] = to_bytes(v, u) if "byte" in u else float(v)
                    if "byte" not in u and u:
                        d[m + "__unit"] = u
                except ValueError:
                    d[m] = v
        if "dram__bytes_read.sum" in d:
            d["dram_bytes_total"] = d["dram__bytes_read.sum"] + d.get("dram__bytes_write.sum", 0.0)
        res.append(d)
    by = collections.OrderedDict()
    for d in res:
        by.setdefault(d["kernel"], []).append(d)
    summary = {}
    for k, lst in by.items():
        summary[k] = {"captures": len(lst)}
        for m in ["gpu__time_duration.sum", "dram_bytes_total", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
                  "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                  "launch__registers_per_thread", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum"]:
            vals = [d[m] for d in lst if isinstance(d.get(m), float)]
            if vals:
                summary[k][m + "__mean"] = sum(vals) / len(vals)
    json.dump({"source": path, "summary": summary, "captures": res}, open(out, "w"), indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
