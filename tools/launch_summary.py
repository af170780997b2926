"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share.
    python tools/launch_summary.py gpurun_out/launches_r1.csv > profiles/r1_launches.txt"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    g, b = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[k]).replace("void ", "").replace("wgb::", "").replace("<unnamed>::", "")
        e = agg.setdefault(name, {"n": 0, "ns": 0.0, "grid": r[g], "block": r[b], "min": 1e30, "max": 0.0})
        t = float(r[v])
        e["n"] += 1
        e["ns"] += t
        e["min"], e["max"] = min(e["min"], t), max(e["max"], t)
    total = sum(e["ns"] for e in agg.values())
    print(f"# {path}: {len(rows) - 1} launches, {total / 1e6:.3f} ms of kernel time (ncu-serialised, cold caches: compare shares)")
    print(f"{'kernel':70s} {'launches':>8s} {'total us':>10s} {'share':>7s} {'min us':>9s} {'max us':>9s}  grid x block")
    for name, e in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"{name[:70]:70s} {e['n']:8d} {e['ns'] / 1e3:10.1f} {100 * e['ns'] / total:6.1f}% {e['min'] / 1e3:9.1f} {e['max'] / 1e3:9.1f}  {e['grid']} x {e['block']}")


if __name__ == "__main__":
    main(sys.argv[1])
