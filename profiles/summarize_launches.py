"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.

usage: python profiles/summarize_launches.py launches.csv "<command that was profiled>" > profiles/rN_ncu_launches_X.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
        name = re.sub(r"^void\s+", "", name)
        name = re.sub(r"tnb::", "", name)
        rows.append((name, v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    total = sum(v[1] for v in agg.values())
    if note:
        print("# " + note)
    print("# launches %d, total kernel time %.3f ms (cold-cache, serialised under ncu: compare SHARES)" % (len(rows), total / 1e3))
    print("share%  launches  avg_us  total_us  kernel")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%5.1f  %7d  %8.1f  %9.1f  %s" % (100 * us / total, c, us / c, us, n))


if __name__ == "__main__":
    main()
