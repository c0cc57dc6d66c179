#!/usr/bin/env python
"""Host-side cost of a frame of many small emitters (GPU box): wall time spent inside batch_update / batch_draw
(both asynchronous: this is enqueue + descriptor building + launch calls), against the device time of the frame."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench  # noqa: E402
from tractor3d_b200.emitter import Context  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ctx = Context(0)
emitters, per, _, name = bench.build_population(ctx, workload, 0, 1, 0)
handles = Context._handles(emitters)
for _ in range(10):
    ctx.batch_update(emitters, bench.DT, handles)
    ctx.batch_draw(emitters, handles)
ctx.sync()
tu = td = 0.0
t_all = time.perf_counter()
for _ in range(frames):
    t0 = time.perf_counter()
    ctx.batch_update(emitters, bench.DT, handles)
    t1 = time.perf_counter()
    ctx.batch_draw(emitters, handles)
    t2 = time.perf_counter()
    tu += t1 - t0
    td += t2 - t1
ctx.flush()
t_host = time.perf_counter() - t_all
ctx.sync()
t_total = time.perf_counter() - t_all
k = [ctx.kernel_time(i) for i in range(4)]
print(f"{name}\n  per frame: batch_update {tu / frames * 1e6:.1f} us, batch_draw {td / frames * 1e6:.1f} us, host loop {t_host / frames * 1e6:.1f} us, "
      f"wall incl. final sync {t_total / frames * 1e6:.1f} us; kernels (ms total, launches): {[(round(a, 2), b) for a, b in k]}")
