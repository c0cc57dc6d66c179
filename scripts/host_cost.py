#!/usr/bin/env python
"""Where the host time of a frame goes (GPU box).  Runs the frames of bench.py's `e2e` leg -- per-frame transforms in,
update + draw, live counts read back -- on one workload and prints, per frame: wall time, device time of the kernels,
and the library's own split of its host time (t3d_ctx_host_profile: enqueue / build / submit / wait), plus the time
spent in Python and ctypes around the calls.

    python scripts/host_cost.py [c2|c4|c3|c5] [frames] [emulated world size]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from tractor3d_b200.emitter import Context  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 100
world = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = Context(0)
emitters, per, _, name = bench.build_population(ctx, workload, 0, world, 0)
handles = Context._handles(emitters)
nodes = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (len(emitters), 1))
cam = np.eye(4, dtype=np.float32).reshape(16)
churn = bench.CHURN.get(workload)
spawn = [churn["spawn"] // len(emitters)] * len(emitters) if churn else None


def frame(read_counts):
    ctx.batch_set_transforms(emitters, nodes, cam, handles)
    if spawn:
        ctx.batch_emit_once(emitters, spawn, handles)
    ctx.batch_update(emitters, bench.DT, handles)
    ctx.batch_draw(emitters, handles)
    return int(ctx.batch_counts(emitters, handles).sum()) if read_counts else 0


for _ in range(churn["warm"] if churn else 10):
    frame(True)
ctx.sync()
print(name)
for read_counts in (True, False):
    p0, k0 = ctx.host_profile(), [ctx.kernel_time(i)[0] for i in range(4)]
    t0 = time.perf_counter()
    for _ in range(frames):
        frame(read_counts)
    ctx.sync()
    wall = time.perf_counter() - t0
    p1, k1 = ctx.host_profile(), [ctx.kernel_time(i)[0] for i in range(4)]
    lib = {k: (p1[k] - p0[k]) / frames * 1e6 for k in p0}
    dev_us = sum(b - a for a, b in zip(k0, k1)) / frames * 1e3
    inside = sum(lib.values())
    print(f"  counts read back every frame: {read_counts}.  per frame: wall {wall / frames * 1e6:.1f} us, kernels {dev_us:.1f} us, "
          f"library host time {inside:.1f} us = " + ", ".join(f"{k} {v:.1f}" for k, v in lib.items())
          + f"; Python/ctypes and the rest {wall / frames * 1e6 - inside:.1f} us")
ctx.close()
