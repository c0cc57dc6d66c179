#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from an ncu report (needs --import-source on / -lineinfo).
    python scripts/ncu_top_stalls.py gpurun_out/s3/prof_c5.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
idx = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h)]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print(rows[0][1][:120])
print("total samples", tot, "instructions", len(data))
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:n]:
    s = int(r[idx["# Samples"]])
    print(f"{s:7d} {100 * s / tot:5.1f}%  lsb={r[idx['stall_long_sb']]:>6} bar={r[idx['stall_barrier']]:>5} ssb={r[idx['stall_short_sb']]:>5} "
          f"wait={r[idx['stall_wait']]:>5} exec={r[idx['Instructions Executed']]:>9}  {r[idx['Address']][-5:]} {r[idx['Source']][:90]}")
