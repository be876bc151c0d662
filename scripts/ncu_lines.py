"""Per-source-line executed-instruction / stall-sample table from an ncu report (scratch tool).
usage: python scripts/ncu_lines.py report.ncu-rep items [min_per_item]"""
import csv, io, subprocess, sys
rep, items = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 6.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out, tot = [], 0.0
for r in rows:
    if len(r) > 8 and r[0].isdigit():
        try:
            ln, ex, samp = int(r[0]), float(r[7]), float(r[4])
        except ValueError:
            continue
        tot += ex
        out.append((ln, ex / items, samp, r[1][:100]))
sm = sum(o[2] for o in out) or 1.0
print(f"total warp-instructions per item {tot / items:.1f}")
for ln, e, s, src in out:
    if e >= thr or 100 * s / sm >= 1.0:
        print(f"{ln:4d} {e:7.1f} {100 * s / sm:5.1f}%  {src}")
