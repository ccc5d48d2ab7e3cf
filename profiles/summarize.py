#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_rNN.csv   > profiles/rNN_launches.txt
    python profiles/summarize.py full gpurun_out/prof_rNN.ncu-rep       > profiles/rNN_ncu_full.txt

`launches`: the `--metrics gpu__time_duration.sum --clock-control none` pass (one row per kernel launch).
`full`: a `--set full` capture; prints the metrics DESIGN.md / bench.py quote (duration, DRAM bytes, issue
utilisation, occupancy limiters) per captured kernel.
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio"]


def short(name):
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name.split("(")[0][-70:]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        if r[ui] == "us":
            v *= 1e3
        elif r[ui] == "ms":
            v *= 1e6
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.3f ms of kernel time (ncu-serialised, cold cache: shares, not absolutes)" %
          (path, len(rows) - 1, tot / 1e6))
    print("%-72s %7s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %7d %12.1f %10.1f %6.1f%%" % (k, a[0], a[1] / 1e3, a[1] / 1e3 / a[0], 100 * a[1] / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    print("# %s (ncu --set full --clock-control none)" % path)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("\n== %s" % short(d["Kernel Name"]))
        for w in WANT:
            if w in d:
                print("  %-70s %18s %s" % (w, d[w], units[hdr.index(w)]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
