#!/usr/bin/env python
"""Prints the headline fields of bench.py JSON lines read from stdin (one per line)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print(f"reference: {d['value']:.4g} {d['unit']} cores={d['cpu_baseline']['cores']} kind={d['cpu_baseline']['kind']}")
        continue
    k = {n: (round(v["ms_per_launch"], 4), round(v["frac"], 3)) for n, v in d["rooflines"].items()}
    print(f"n_gpus={d['n_gpus']} value={d['value']:.4g} ms/step={d['ms_per_step']:.3f} kernels(ms,frac)={k} "
          f"e2e={d['e2e']['value']:.4g} launches={d['gpu_launches']} live={d['config']['live_after']} clocks={d['clocks']}")
    print("   workload:", d["config"]["workload"])
    if d.get("cpu_baseline"):
        print("   cpu_baseline:", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
    for name, a in (d.get("also") or {}).items():
        k = {n: (round(v["ms_per_launch"], 4), round(v["frac"], 3)) for n, v in a.get("rooflines", {}).items()}
        print(f"   also.{name}: value={a['value']:.4g} ms/step={a['ms_per_step']:.3f} live={a['particles']} "
              f"replaced/frame={a.get('replaced_per_frame', 0):.4f} kernels(ms,frac)={k}")
    v = d.get("e2e_vertices_to_host")
    if v:
        print("   e2e_vertices_to_host:", {k: (round(x, 4) if isinstance(x, float) else x) for k, x in v.items() if k != "what"})
