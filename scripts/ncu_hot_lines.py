"""Top stall-sample instructions of the first kernel in an ncu report (scratch tool).  usage: ncu_hot_lines.py rep [N]"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
isrc, isamp = h.index("Source"), h.index("# Samples")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[2:] if len(r) > isamp]
tot = sum(float(r[isamp] or 0) for r in body)
agg = {}
for i in stall_cols:
    agg[h[i]] = sum(float(r[i] or 0) for r in body)
print("total samples", tot)
print({k: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
idx = sorted(range(len(body)), key=lambda i: -float(body[i][isamp] or 0))[:N]
for i in sorted(idx):
    r = body[i]
    top = sorted(((float(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {r[isrc].strip()[:64]:64s} {float(r[isamp]):7.0f}  " + " ".join(f"{n}:{v:.0f}" for v, n in top))
