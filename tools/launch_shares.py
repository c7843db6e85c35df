#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel launch counts, total
time and share.   python tools/launch_shares.py launches.csv [first [count]]   (first/count select a window)"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    count = int(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()
        rows.append((name, ms))
    rows = rows[first:first + count] if count else rows[first:]
    agg = collections.OrderedDict()
    for name, ms in rows:
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + ms)
    total = sum(t for _, t in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'share':>7s}")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {c:8d} {t:10.3f} {100 * t / total:6.1f}%")
    print(f"{'TOTAL':60s} {len(rows):8d} {total:10.3f}")


if __name__ == "__main__":
    main()
