#!/usr/bin/env python
"""Churn at scale against the oracle (GPU box): K emitters, `spawn` particles spawned per emitter per frame by the device
generator, update + draw every frame through the batch calls, live counts read back only every `check` frames (so the
asynchronous count feedback sizes the grids in between).  usage: churn_vs_oracle.py [frames] [spawn] [emitters] [check]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpu_harness as H  # noqa: E402
from gpu_adapter import GpuLib  # noqa: E402
from tractor3d_b200 import abi, presets  # noqa: E402
from tractor3d_b200.emitter import Context  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
spawn = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
check = int(sys.argv[4]) if len(sys.argv) > 4 else 10
O = H.oracle(fresh=True)
lib = GpuLib(rng_mode=abi.RNG_DEVICE, seed=0)
eos, egs = [], []
for k in range(K):
    p, sprite = presets.headline_64m(capacity=spawn * 64)
    p.energy_min, p.energy_max = 100.0, 600.0
    eo = O.emitter(p)
    eo.set_sprite_frame_coords(*sprite)
    O.emitter_set_device_rng(eo.h, 1, 100 + k)
    eg = lib.emitter(p)
    eg.set_sprite_frame_coords(*sprite)
    eg.pe.set_rng(abi.RNG_DEVICE, 100 + k)
    eos.append(eo)
    egs.append(eg)
pes = [e.pe for e in egs]
h = Context._handles(pes)
ok = True
for f in range(frames):
    for eo in eos:
        eo.emit_once(spawn)
    for eo in eos:
        eo.update(1000.0 / 60.0)
    for eo in eos:
        eo.draw()
    lib.ctx.batch_emit_once(pes, [spawn] * K, h)
    lib.ctx.batch_update(pes, 1000.0 / 60.0, h)
    lib.ctx.batch_draw(pes, h)
    if f % check == check - 1 or f == frames - 1:
        cg = list(map(int, lib.ctx.batch_counts(pes, h)))
        co = [eo.count() for eo in eos]
        print("frame", f, "oracle", co, "gpu", cg, flush=True)
        if co != cg:
            ok = False
            break
if ok:
    for k in range(K):
        assert np.array_equal(egs[k].state(), eos[k].state()), k
        assert np.array_equal(egs[k].vertices().view(np.uint32), eos[k].vertices().view(np.uint32)), k
    print("state and vertices bit-identical")
print("launches fused/update/draw", lib.ctx.kernel_time(3)[1], lib.ctx.kernel_time(1)[1], lib.ctx.kernel_time(2)[1])
lib.close()
sys.exit(0 if ok else 1)
