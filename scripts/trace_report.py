#!/usr/bin/env python
"""Summarises a KTF_TC_TRACE file (per-tile clock64 stamps of cluster 0 of every pair-kernel launch, see
kaldi_tflite_b200/csrc/tdnn_tc.cu): per launch, the median tile period and where the MMA thread and the epilogue wait.

    KTF_TC_TRACE=gpurun_out/trace.txt python bench.py --no-cpu-baseline --no-stages --steps 1 --warmup 3
    python scripts/trace_report.py gpurun_out/trace.txt [min_tiles]

Columns (SM clocks): period = accumulator-free to accumulator-free of consecutive tiles; free->full = MMA of a tile incl.
its operand waits; free->operands = until the last k-block of the tile has landed; epilogue = accumulator full until epilogue
warp 2 has drained its part; slack = epilogue of tile i done until the MMA of tile i + 2 (same accumulator) starts.
"""
import statistics as st
import sys


def parse(path):
    launches, cur = [], None
    for line in open(path):
        if line.startswith("launch"):
            cur = {"hdr": line.strip(), "rows": []}
            launches.append(cur)
        elif cur is not None:
            cur["rows"].append([int(x) for x in line.split()])
    return launches


def main(path, min_tiles=40):
    seen = set()
    for l in parse(path):
        rows, n = l["rows"], len(l["rows"])
        if n <= min_tiles or l["hdr"] in seen:
            continue
        seen.add(l["hdr"])
        body = range(5, n - 7)
        per = [rows[i + 1][1] - rows[i][1] for i in body]
        mma = [rows[i][3] - rows[i][1] for i in body]
        opw = [rows[i][2] - rows[i][1] for i in body]
        epi = [rows[i][4] - rows[i][3] for i in body]
        lag = [rows[i + 2][1] - rows[i][4] for i in body]
        print(l["hdr"])
        print(f"   tiles {n}  period {st.median(per):.0f}  free->full {st.median(mma):.0f}  free->operands "
              f"{st.median(opw):.0f}  epilogue {st.median(epi):.0f}  slack {st.median(lag):.0f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
