#!/usr/bin/env python
"""Register the measured DRAM traffic of a dominant kernel for bench.py's `roofline.traffic`.

    python profiles/ncu_traffic.py gpurun_out/X.ncu-rep KEY profiles/r02/X_raw.csv

Reads `dram__bytes_read.sum + dram__bytes_write.sum` of the (single) kernel in an `ncu --set full` report, writes the raw
metric page to the CSV given (committed next to this file) and records {KEY: {dram_bytes, csv, source_sha, kernel}} in
profiles/traffic.json. KEY = "<kernel family>:<bench config>", e.g. "k3_ppo_fast_kernel:c2". `source_sha` is the hash of the
sources of that kernel family at capture time (bench.k3_source_hash): bench.py reports `traffic: null` with the reason when the sources have
changed since, instead of a stale constant.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import k3_source_hash  # noqa: E402


def main(rep, key, csv_out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(csv_out, "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name):
        v, u = float(vals[col[name]].replace(",", "")), units[col[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    path = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[key] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": vals[col["Kernel Name"]],
                  "duration": float(vals[col["gpu__time_duration.sum"]].replace(",", "")),
                  "duration_unit": units[col["gpu__time_duration.sum"]] + " (under ncu, cold cache)",
                  "csv": os.path.relpath(csv_out, ROOT), "source_sha": k3_source_hash(key)}
    json.dump(table, open(path, "w"), indent=1)
    print(key, table[key])


if __name__ == "__main__":
    main(*sys.argv[1:4])
