"""Reads the fixtures written by tests/golden/make_golden.py."""
import json
import os

import numpy as np

from tractor3d_b200 import abi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    params = abi.EmitterParams.from_dict(json.loads(bytes(z["params"]).decode()))
    lens = z["draw_lens"]
    draws = z["draws"]
    per_step, pos = [], 0
    for n in lens:
        per_step.append(draws[pos:pos + n])
        pos += n
    snaps = []
    for k, (step, cnt, act) in enumerate(zip(z["snap_steps"], z["snap_counts"], z["snap_active"])):
        snaps.append({"step": int(step), "count": int(cnt), "active": bool(act), "state": z[f"state_{k}"],
                      "vertices": z[f"vertices_{k}"].view(np.float32)})
    return {"params": params, "per_step_draws": per_step, "all_draws": draws, "snaps": snaps, "n_steps": int(z["n_steps"])}
