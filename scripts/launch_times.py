#!/usr/bin/env python
"""Prints the per-launch device times of an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv
import sys


def load(path):
    with open(path, errors="replace") as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        out.append((row["Kernel Name"], v, row["Grid Size"], row["Block Size"]))
    return out


if __name__ == "__main__":
    rows = load(sys.argv[1])
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
    for name, us, grid, block in rows[lo:hi]:
        print(f"{us:10.1f} us {grid:>16s} {block:>12s}  {name[:100]}")
    print(f"# {len(rows)} launches, {sum(r[1] for r in rows[lo:hi]):.1f} us in the printed range")
