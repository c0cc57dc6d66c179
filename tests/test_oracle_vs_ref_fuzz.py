"""Randomised API sequences on the CPU: the plain-C oracle against the UNMODIFIED reference (oracle/_ref), sharing one
libc-style draw stream, several emitters interleaved.  The GPU fuzz (tests/test_gpu_fuzz.py) trusts the oracle on exactly
these sequences -- phase changes, parameter edits between frames, dt below the 8 ms rate cap with several emitters
sharing the accumulator, capacity clamping, start/stop, node moves -- so the oracle is pinned on them here, bit for bit."""
import numpy as np
import pytest

import cpu_harness as H
import scenarios as S
from tractor3d_b200 import abi
from test_gpu_fuzz import random_params

pytestmark = pytest.mark.skipif(not H.ref_available(), reason="oracle/_ref/libt3dref.so not built")


@pytest.mark.parametrize("seed,rotate", [(s, r) for s in range(6) for r in (False, True)])
def test_random_api_sequences_oracle_equals_reference(seed, rotate):
    rng = np.random.default_rng(4000 + seed)
    O, R = H.oracle(fresh=True), H.ref(fresh=True)
    O.rand_seed(99 + seed)
    R.rand_seed(99 + seed)
    pairs = []
    for k in range(int(rng.integers(2, 5))):
        p, sprite = random_params(rng, rotate)
        eo, er = O.emitter(p), R.emitter(p)
        for e in (eo, er):
            e.set_sprite_frame_coords(*sprite)
        pairs.append((eo, er, p))
    cam = S.rot_xyz(0.3, -0.2, 0.1, (1, 2, 9))
    for eo, er, _ in pairs:
        eo.set_camera_world(cam); er.set_camera_world(cam)

    def check():
        for i, (eo, er, _) in enumerate(pairs):
            assert eo.count() == er.count(), f"emitter {i}"
            assert eo.is_active() == er.is_active(), f"emitter {i}"
            a, b = eo.state(), er.state()
            if not np.array_equal(a, b):
                c, j = np.argwhere(a != b)[0]
                raise AssertionError(f"emitter {i}: column {abi.COLUMN_NAMES[c]} particle {j}: {a[c, j]:#x} vs {b[c, j]:#x}")

    for step in range(120):
        op = rng.choice(["update", "update", "update", "draw", "emit", "start", "stop", "rate", "node", "params", "dt_small"])
        k = int(rng.integers(0, len(pairs)))
        eo, er, p = pairs[k]
        if op in ("update", "dt_small"):
            dt = float(rng.choice([16.6667, 33.3, 8.0, 50.0] if op == "update" else [2.5, 3.0, 5.5]))
            for a, b, _ in pairs:
                a.update(dt); b.update(dt)
        elif op == "draw":
            for a, b, _ in pairs:
                assert a.draw() == b.draw()
                va, vb = a.vertices(), b.vertices()
                assert va.shape == vb.shape and np.array_equal(va.view(np.uint32), vb.view(np.uint32))
        elif op == "emit":
            n = int(rng.integers(1, 1200))
            eo.emit_once(n); er.emit_once(n)
        elif op == "start":
            eo.start(); er.start()
        elif op == "stop":
            eo.stop(); er.stop()
        elif op == "rate":
            r = int(rng.integers(1, 2000))
            eo.set_emission_rate(r); er.set_emission_rate(r)
        elif op == "node":
            w = S.rot_xyz(*rng.uniform(-1, 1, 3), tuple(rng.uniform(-3, 3, 3)))
            eo.set_node_world(w); er.set_node_world(w)
        elif op == "params":
            q = p.copy()
            q.energy_min, q.energy_max = float(rng.integers(20, 300)), float(rng.integers(300, 1500))
            q.sprite_looped = int(rng.integers(0, 2))
            q.velocity = abi.F3(*rng.uniform(-5, 5, 3))
            q.ellipsoid = int(rng.integers(0, 2))
            eo.set_params(q); er.set_params(q)
            pairs[k] = (eo, er, q)
        if step % 20 == 19:
            check()
    check()
    assert O.rand_calls() == R.rand_calls()   # both consumed the shared stream to the same position


@pytest.mark.parametrize("n", [4609, 200_003])
def test_uploaded_populations_with_heavy_dying_oracle_equals_reference(n):
    # The multi-tile GPU tests upload a synthetic live population and let it die out frame by frame: the removal order
    # (swap-with-last, the moved particle skipped for a frame) is what they check.  Same populations, oracle against the
    # unmodified reference, state and vertices after every frame.
    from test_gpu_parity import _big_population
    from tractor3d_b200 import presets
    p, sprite = presets.headline_64m(capacity=n + 5)
    p.energy_min, p.energy_max = 20.0, 400.0
    O, R = H.oracle(fresh=True), H.ref(fresh=True)
    eo, er = O.emitter(p), R.emitter(p)
    eo.set_sprite_frame_coords(*sprite); er.set_sprite_frame_coords(*sprite)
    cols = _big_population(n, seed=n + 7)
    eo.set_state(cols); er.set_state(cols)
    for f in range(14):
        eo.update(S.DT); er.update(S.DT)
        assert eo.count() == er.count(), f"frame {f}"
        assert np.array_equal(eo.state(), er.state()), f"frame {f}"
        assert eo.draw() == er.draw()
        va, vb = eo.vertices(), er.vertices()
        assert va.shape == vb.shape and np.array_equal(va.view(np.uint32), vb.view(np.uint32)), f"frame {f}"
        if eo.count() == 0:
            break
