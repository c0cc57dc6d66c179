"""The C++ host-side mirror of the reference interface (tractor3d_b200/host/ParticleEmitter.{h,cpp}): a small
program written against it the way ParticlesSample is written against the reference class is run on the GPU
and its result compared, bit for bit, with the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cpu_harness as H
import scenarios as S
from test_abi_and_host import PARTICLE_TEXT, _parse, _write_png
from tractor3d_b200 import abi, build

pytestmark = pytest.mark.gpu


def test_cpp_facade_matches_oracle(tmp_path):
    driver = build.build_facade_driver()
    _write_png(str(tmp_path / "sheet.png"), 96, 32)
    f = tmp_path / "sparks.particle"
    f.write_text(PARTICLE_TEXT.replace("looped = true", "looped = false"))
    frames, seed, burst = 90, 4242, 50
    out = tmp_path / "out.bin"
    res = subprocess.run([driver, str(f), str(frames), str(seed), str(out), str(burst)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Failed to create particle emitter from file" in res.stderr      # the nullptr path logged like GP_ERROR
    raw = np.fromfile(out, dtype=np.uint32)
    count, quads, draws, frame_count = [int(x) for x in raw[:4]]
    state = raw[4:4 + abi.NUM_COLUMNS * count].reshape(abi.NUM_COLUMNS, count)
    verts = raw[4 + abi.NUM_COLUMNS * count:].view(np.float32).reshape(quads, 4, 9)

    lib = abi.load()
    rc, p, fc, sw, sh, _ = _parse(lib, str(f))
    assert rc == abi.OK and fc == frame_count
    O = H.oracle(fresh=True)
    eo = O.emitter(p)
    eo.set_sprite_frame_coords(fc, sw, sh)
    O.emitter_set_device_rng(eo.h, 1, seed)
    cam = np.eye(4, dtype=np.float32); cam[:3, 3] = (0.22, 2.7, 13.3)
    eo.set_camera_world(cam.T.reshape(16))
    eo.emit_once(burst)
    eo.start()
    n_draws = 0
    for fr in range(frames):
        w = np.eye(4, dtype=np.float32); w[0, 3] = np.float32(0.05) * np.float32(fr)
        eo.set_node_world(w.T.reshape(16))
        eo.update(S.DT)
        n_draws += eo.draw()
    assert (count, draws) == (eo.count(), n_draws) and count > 20
    so = eo.state()
    fa, fb = state.view(np.float32), so.view(np.float32)
    for c in abi.INT_COLUMNS:
        assert np.array_equal(state[c], so[c])
    # this emitter axis-rotates (rotationSpeed != 0): createRotation runs on the device -> stated tolerance
    for c in range(abi.NUM_COLUMNS):
        if c not in abi.INT_COLUMNS:
            assert np.allclose(fa[c], fb[c], rtol=1e-5, atol=1e-6), abi.COLUMN_NAMES[c]
    assert quads == count and np.allclose(verts, eo.vertices(), rtol=1e-5, atol=1e-6)
