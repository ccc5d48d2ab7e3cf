#!/usr/bin/env python
"""Refreshes the ncu-derived constants bench.py quotes from a `ncu --set full` capture of `bench.py` itself.

    python profiles/update_traffic.py gpurun_out/rNN_full.ncu-rep

Writes profiles/k3_scan_traffic.json (DRAM bytes per k3_scan_v3 launch: the `roofline.traffic` of the bench line) and
profiles/k6_ncu_summary.json (duration, warp instructions, issue-slot utilisation of the verification kernels).  A
number measured under ncu is never a bench value: only traffic / instruction counts / utilisation are taken from it.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6}.get(u, 1.0)
        return v * scale

    per = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        short = next((k for k in ("k3_scan_v3", "k6a_table", "k6b_components", "k6c_kabsch", "k6d_rows", "k3_topn_sort",
                                  "k3_lookup") if k in name), None)
        if not short:
            continue
        per.setdefault(short, []).append({
            "ms": val(r, "gpu__time_duration.sum"),
            "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
            "warp_insts": val(r, "smsp__inst_executed.sum"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")})
    mean = lambda xs, k: sum(x[k] for x in xs) / len(xs)
    if "k3_scan_v3" in per:
        xs = per["k3_scan_v3"]
        json.dump({"kernel": "k3_scan_v3", "dram_bytes_per_launch": int(mean(xs, "dram_bytes")), "launches_captured": len(xs),
                   "warp_insts_per_launch": int(mean(xs, "warp_insts")), "issue_active_pct": round(mean(xs, "issue_active_pct"), 1),
                   "source": os.path.basename(rep), "command": "bench.py (default workload) under ncu --set full"},
                  open(os.path.join(ROOT, "profiles", "k3_scan_traffic.json"), "w"), indent=1)
    k6 = {k: {"launches_captured": len(xs), "ms_per_launch_under_ncu": round(mean(xs, "ms"), 4),
              "warp_insts_per_launch": int(mean(xs, "warp_insts")), "dram_bytes_per_launch": int(mean(xs, "dram_bytes")),
              "issue_active_pct": round(mean(xs, "issue_active_pct"), 1), "warps_active_pct": round(mean(xs, "warps_active_pct"), 1)}
          for k, xs in per.items() if k != "k3_scan_v3"}
    k6["source"] = os.path.basename(rep)
    json.dump(k6, open(os.path.join(ROOT, "profiles", "k6_ncu_summary.json"), "w"), indent=1)
    print(json.dumps(per.keys().__repr__()))


if __name__ == "__main__":
    main()
