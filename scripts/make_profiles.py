#!/usr/bin/env python
"""Copies / condenses what scripts/gpu/final_n1.sh brought back in gpurun_out/final into the tracked files under profiles/
(bench lines, ncu launch list, one `ncu --set full` summary + stall list per workload, DRAM traffic per particle).

    python scripts/make_profiles.py [gpurun_out/final] [r2]
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "final")
tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
P = os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def summarize(mode, path):
    return subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_ncu.py"), mode, path], capture_output=True, text=True, check=True).stdout


shutil.copy(os.path.join(src, "bench_n1.json"), os.path.join(P, f"{tag}_bench_n1.json"))
shutil.copy(os.path.join(src, "bench_reference.json"), os.path.join(P, f"{tag}_bench_reference.json"))
shutil.copy(os.path.join(src, "launches.csv"), os.path.join(P, f"{tag}_launches.csv"))
open(os.path.join(P, f"{tag}_launches.md"), "w").write(summarize("launches", os.path.join(src, "launches.csv")))
shutil.copy(os.path.join(src, "host_cost.txt"), os.path.join(P, f"{tag}_host_cost.txt"))
shutil.copy(os.path.join(src, "sprites.txt"), os.path.join(P, f"{tag}_sprites_2d.txt"))

traffic = {}
for w in ("c3", "c3d", "c5", "c2", "c4", "c3_sep", "spawn"):
    raw = os.path.join(src, f"prof_{w}_raw.csv")
    if not os.path.exists(raw):
        continue
    open(os.path.join(P, f"{tag}_ncu_full_{w}.md"), "w").write(summarize("full", raw))
    stalls = os.path.join(src, f"prof_{w}_stalls.txt")
    if os.path.exists(stalls) and os.path.getsize(stalls):
        shutil.copy(stalls, os.path.join(P, f"{tag}_stalls_{w}.txt"))
    if w == "spawn":
        continue
    # live particles at the capture: the bench line the profiled run printed
    log = open(os.path.join(src, f"ncu_{w}.log")).read()
    m = re.search(r'"live_after": (\d+)', log)
    units = int(m.group(1)) if m else None
    rows = list(csv.reader(io.StringIO(open(raw).read())))
    hdr, unit_row = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT[unit_row[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT[unit_row[idx["dram__bytes_write.sum"]]]
        kind = "draw" if "k_draw" in name else ("update_draw" if w != "c3_sep" else "update")
        traffic.setdefault(w.replace("_sep", ""), {})[kind] = {
            "bytes_per_unit": (rd + wr) / units, "dram_read_bytes": rd, "dram_write_bytes": wr, "units": units, "kernel": name,
            "ncu_duration": r[idx["gpu__time_duration.sum"]] + " " + unit_row[idx["gpu__time_duration.sum"]]}
traffic["_about"] = ("DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per live particle of one launch, from one `ncu --set full "
                     "--clock-control none` capture per workload (scripts/gpu/final_n1.sh); units = live particles at the capture. The spawn "
                     "kernel's 147 MB per launch mostly stays in the 126 MB L2 until after the kernel (ncu sees 86-88 MB of DRAM writes), so no "
                     "per-unit figure is given for it. bench.py reads this file for roofline.traffic.")
json.dump(traffic, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
for w, k in traffic.items():
    if w != "_about":
        print(w, {n: round(v["bytes_per_unit"], 1) for n, v in k.items()})
