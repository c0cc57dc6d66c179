#!/usr/bin/env python
"""Runs a churn workload for N frames and prints per-frame live counts plus checksums of the final state and vertices,
so that two runs (T3D_FUSED=0 / T3D_FUSED=1) can be diffed.  usage: fused_vs_separate.py [frames] [spawn_per_emitter] [emin,emax]"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tractor3d_b200 import abi, presets  # noqa: E402
from tractor3d_b200.emitter import Context, ParticleEmitter  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
spawn = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
emin, emax = (float(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "100,600").split(","))
ctx = Context(0)
es = []
for k in range(4):
    p, sprite = presets.headline_64m(capacity=spawn * 64)
    p.energy_min, p.energy_max = emin, emax
    es.append(ParticleEmitter.create(ctx, p, sprite, (abi.RNG_DEVICE, 100 + k)))
h = Context._handles(es)
for f in range(frames):
    ctx.batch_emit_once(es, [spawn] * len(es), h)
    ctx.batch_update(es, 1000.0 / 60.0, h)
    ctx.batch_draw(es, h)
    if f % 10 == 9 or f == frames - 1:
        print("frame", f, "counts", list(map(int, ctx.batch_counts(es, h))))
for k, e in enumerate(es):
    st = e.state()
    vt = e.vertices()
    print("emitter", k, "state", hashlib.sha1(np.ascontiguousarray(st).tobytes()).hexdigest()[:16], st.shape,
          "vertices", hashlib.sha1(np.ascontiguousarray(vt).tobytes()).hexdigest()[:16], vt.shape)
print("launches fused/update/draw:", ctx.kernel_time(3)[1], ctx.kernel_time(1)[1], ctx.kernel_time(2)[1])
