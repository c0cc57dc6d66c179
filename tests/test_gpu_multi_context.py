"""One process, one thread, one context per GPU -- what a C++ engine does (tests/multi_ctx_driver.cpp, C ABI only).
The reference keeps the update rate-cap accumulator (PE.cpp:720) and the latched billboard rotation (SB.cpp:339) in
function-local statics, i.e. once per process; so does the library, and splitting the emitters over contexts must not be
observable: the split run equals the single-context run AND the oracle's (one oracle library = one process)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cpu_harness as H
from tractor3d_b200 import abi, build, presets

pytestmark = pytest.mark.gpu

K = 7   # emitters: not a multiple of the context counts used below


def _records(tmp_path):
    recs = []
    raw = b""
    for e in range(K):
        p, sprite = presets.MIXED[e % 3]()
        p.particle_count_max = 3000
        p.emission_rate = 400 + 70 * e
        p.energy_min, p.energy_max = 300.0, 900.0
        if e % 2 == 0:      # some emitters spin their quads: the first rotated quad of the process latches the rotation
            p.rotation_per_particle_speed_min, p.rotation_per_particle_speed_max = 0.5, 2.5
        seed = 9000 + e
        recs.append((p, sprite, seed))
        # struct Record of multi_ctx_driver.cpp: params, uint32, int32, int32, 4 bytes of padding, uint64
        raw += bytes(p) + np.array([sprite[0]], dtype=np.uint32).tobytes() + np.array(sprite[1:] + (0,), dtype=np.int32).tobytes()
        raw += np.array([seed], dtype=np.uint64).tobytes()
    assert len(raw) == K * (C.sizeof(abi.EmitterParams) + 24)
    f = tmp_path / "params.bin"
    f.write_bytes(raw)
    return recs, str(f)


def _oracle(recs, dt, frames):
    O = H.oracle(fresh=True)
    es = []
    cam = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.2, 2.5, 13.0, 1], dtype=np.float32)
    for p, sprite, seed in recs:
        e = O.emitter(p); e.set_sprite_frame_coords(*sprite)
        O.emitter_set_device_rng(e.h, 1, seed)
        e.set_camera_world(cam); e.start()
        es.append(e)
    for f in range(frames):
        for k, e in enumerate(es):
            w = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, np.float32(0.01) * np.float32(f) * np.float32(k + 1),
                          np.float32(0.5) * np.float32(k), 0, 1], dtype=np.float32)
            e.set_node_world(w)
        for e in es:
            e.update(dt)
        for e in es:
            e.draw()
    return es


def _read_states(path):
    raw = np.fromfile(path, dtype=np.uint32)
    out, pos = [], 0
    for _ in range(K):
        n = int(raw[pos]); pos += 1
        out.append(raw[pos:pos + abi.NUM_COLUMNS * n].reshape(abi.NUM_COLUMNS, n)); pos += abi.NUM_COLUMNS * n
    assert pos == raw.size
    return out


@pytest.mark.parametrize("n_ctx,dt", [(2, 1000.0 / 60.0), (2, 5.0), (3, 3.0)])
def test_contexts_of_one_process_share_the_reference_statics(tmp_path, n_ctx, dt):
    import torch
    driver = build.build_multi_ctx_driver()
    recs, params = _records(tmp_path)
    frames = 90 if dt > 8 else 240
    layouts = ["same"] + (["spread"] if torch.cuda.device_count() >= n_ctx else [])
    want = None
    for layout in layouts:
        out = tmp_path / f"out_{layout}.bin"
        res = subprocess.run([driver, params, str(n_ctx), layout, repr(dt), str(frames), str(out)], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        assert "EQUALS" in res.stdout
        if want is None:
            want = _oracle(recs, np.float32(dt), frames)
        got = _read_states(out)
        for k, e in enumerate(want):
            assert got[k].shape[1] == e.count() > 50, k
            assert np.array_equal(got[k], e.state()), k


def test_private_statics_make_the_split_observable(tmp_path):
    # The opt-out (t3d_ctx_set_private_statics): every context accumulates its own update time, so with 5 ms steps the
    # emitters of different contexts no longer take turns at the 8 ms gate the way the reference's do.
    driver = build.build_multi_ctx_driver()
    recs, params = _records(tmp_path)
    out = tmp_path / "out.bin"
    res = subprocess.run([driver, params, "2", "same", "5.0", "120", str(out), "private"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 1 and "DIFFERS" in res.stdout, res.stdout + res.stderr
    # with steps of a whole frame every update passes the gate either way: the split is invisible again -- except for the
    # rotation each context latches for itself, so compare the states only where that plays no part
    res = subprocess.run([driver, params, "2", "same", "16.666666", "60", str(out), "private"], capture_output=True, text=True, timeout=300)
    assert res.returncode in (0, 1), res.stdout + res.stderr
    got = _read_states(out)
    want = _oracle(recs, np.float32(16.666666), 60)
    for k, e in enumerate(want):
        assert np.array_equal(got[k], e.state()), k
