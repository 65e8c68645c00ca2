#!/usr/bin/env python
"""Aggregate an `ncu --set full --import-source on` report per CUDA source line / per phase.

    python profiles/ncu_lines.py gpurun_out/k3.ncu-rep [--top 40] [--phases]

Uses `ncu -i REP --page source --csv --print-source cuda,sass`; stall samples per line are
proportional to warp-time spent at that line.
"""
import csv
import io
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, agg = None, {}
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 10 and r[2] == "-":
            try:
                agg[(cur, int(r[0]))] = (int(r[4]), int(r[7]), r[1].strip()[:100])
            except ValueError:
                pass
    return agg


if __name__ == "__main__":
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    agg = load(rep)
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print(f"total samples {tot}  total warp-instructions {toti}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{k[0]:22s}:{k[1]:4d} samp {100 * v[0] / tot:5.1f}% inst {100 * v[1] / toti:5.1f}%  {v[2]}")
