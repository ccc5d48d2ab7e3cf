#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu report (SASS page) + nvdisasm line info.

    python profiles/sass_lines.py REPORT.ncu-rep OBJECT.o MANGLED_SUBSTRING [top [DEMANGLED_SUBSTRING]]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, sub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
    line_of = {}
    cur, active, inl = None, False, None
    for ln in sass.splitlines():
        if ln.startswith("//--------------------- .text."):
            active = sub in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m and cur:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr = None
    per = collections.Counter()
    stall = collections.Counter()
    total = 0
    want = sys.argv[5] if len(sys.argv) > 5 else None  # substring of the demangled kernel name (multi-kernel reports)
    active = True
    for r in rows:
        if r and r[0] == "Kernel Name":
            active = want is None or want in r[1]
            continue
        if not active:
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            addr = int(d["Address"], 16) if d["Address"].startswith("0x") else int(d["Address"])
        except ValueError:
            continue
        n = int(float(d["Instructions Executed"] or 0))
        s = int(float(d.get("Warp Stall Sampling (All Samples)") or 0))
        # ncu addresses are absolute; nvdisasm offsets are relative to the function start
        per[addr] = (n, s)  # several launches of the kernel in the report: the last one
    base = min(per) if per else 0
    agg, agg_s = collections.Counter(), collections.Counter()
    for addr, (n, s) in per.items():
        key = line_of.get(addr - base, (("?", 0), ""))[0]
        agg[key] += n
        agg_s[key] += s
    ts = sum(agg_s.values()) or 1
    total = sum(n for n, _ in per.values())
    regions = os.environ.get("SASS_REGIONS")  # "name:first-last,name:first-last" of the main .cu file: region sums
    if regions:
        for spec in regions.split(","):
            name, rng = spec.split(":")
            lo, hi = [int(x) for x in rng.split("-")]
            rn = sum(n for (f, l), n in agg.items() if f.endswith(".cu") and lo <= l <= hi)
            rs = sum(n for (f, l), n in agg_s.items() if f.endswith(".cu") and lo <= l <= hi)
            print("# region %-28s lines %5d-%-5d %12d insts %5.1f%%  stalls %5.1f%%" % (name, lo, hi, rn, 100.0 * rn / max(total, 1), 100.0 * rs / ts))
    print("# %s: %d warp instructions in the captured launch" % (sub, total))
    print("%-22s %14s %7s %8s" % ("file:line", "warp insts", "share", "stalls"))
    for key, n in agg.most_common(top):
        print("%-22s %14d %6.1f%% %7.1f%%" % ("%s:%d" % key, n, 100.0 * n / max(total, 1), 100.0 * agg_s[key] / ts))


if __name__ == "__main__":
    main()
