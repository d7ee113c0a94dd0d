#!/usr/bin/env python3
"""Per-kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: summarize_launches.py launches.csv > summary.csv"""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((r["Kernel Name"], us))
    agg = OrderedDict()
    for name, us in rows:
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + us)
    total = sum(t for _, t in agg.values())
    print(f"# {len(rows)} launches, total {total:.1f} us (cold-cache, serialised under ncu: compare SHARES)")
    print("kernel,launches,total_us,avg_us,share_pct")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"\"{name}\",{n},{t:.1f},{t / n:.2f},{100.0 * t / total:.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
