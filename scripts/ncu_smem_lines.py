"""Per-instruction shared-memory wavefronts of the first kernel in an ncu report (scratch tool).
usage: ncu_smem_lines.py report.ncu-rep [items]   -- items = work items per launch, to print per-item figures"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]
items = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ia, isrc, iw, iid, ix, ie, ic = (h.index(n) for n in ("Address", "Source", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                                              "Instructions Executed", "L1 Wavefronts Shared Excessive", "# Samples"))
tot = 0
for r in rows[2:]:
    if len(r) <= iw:
        continue
    w = float(r[iw] or 0)
    if w > 0:
        tot += w
        print(f"{r[isrc].strip()[:70]:70s} exec/item {float(r[ix])/items:6.2f} wf/item {w/items:7.2f} ideal {float(r[iid] or 0)/items:7.2f}")
print("total wavefronts per item", tot / items)
