"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libt3dref.so).

Run here, where /root/reference exists and `make -C oracle ref` has been done:
    python tests/golden/make_golden.py
Each file holds one scenario of tests/scenarios.py: the emitter parameters (JSON), the raw draws the
reference's rand() calls consumed at every step, and the snapshots (live count, full particle state,
vertex stream) at every check step.  The fixtures travel to machines without the reference (the GPU box).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cpu_harness as H  # noqa: E402
import scenarios as S  # noqa: E402

SEED = 1234


def main():
    assert H.ref_available(), "build oracle/_ref first: make -C oracle ref"
    for sc in S.all_scenarios():
        R = H.ref(fresh=True)  # fresh statics per scenario: a new process as far as the reference can tell
        R.rand_seed(SEED)
        e = S.new_emitter(R, sc)
        snaps, rec = S.run(e, sc, record_lib=R)
        out = {
            "params": np.frombuffer(S.params_json(sc.params).encode(), dtype=np.uint8),
            "n_steps": np.array(len(sc.steps)),
            "draw_lens": np.array([len(r) for r in rec], dtype=np.int64),
            "draws": np.concatenate(rec).astype(np.int32) if rec else np.zeros(0, np.int32),
            "snap_steps": np.array([s["step"] for s in snaps], dtype=np.int64),
            "snap_counts": np.array([s["count"] for s in snaps], dtype=np.uint32),
            "snap_active": np.array([s["active"] for s in snaps], dtype=np.uint8),
        }
        for k, s in enumerate(snaps):
            out[f"state_{k}"] = s["state"]
            out[f"vertices_{k}"] = s["vertices"].view(np.uint32)
        path = os.path.join(HERE, sc.name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{sc.name}: {len(sc.steps)} steps, {len(out['draws'])} draws, {len(snaps)} snapshots, "
              f"final count {snaps[-1]['count']}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
