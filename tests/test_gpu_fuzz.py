"""Randomised API sequences: several emitters with random parameter sets are driven through random interleavings of
emitOnce / update / draw / start / stop / setters / node and camera moves / capacity changes, on the GPU (through the C ABI) and on the oracle, and
must agree at every checkpoint.  Exercises what the scripted scenarios do not: the command queue (phase changes,
flushes forced mid-phase), the parity bookkeeping of the live counts, count feedback, the shared rate-cap accumulator
with several emitters and dt < 8 ms, capacity clamping in the kernel, parameter changes between frames."""
import numpy as np
import pytest

import cpu_harness as H
import scenarios as S
from tractor3d_b200 import abi

pytestmark = pytest.mark.gpu


def random_params(rng, rotate):
    u = lambda a, b: float(rng.uniform(a, b))
    v3 = lambda s: [u(-s, s) for _ in range(3)]
    animated = int(rng.integers(0, 2))
    p = abi.default_params(
        particle_count_max=int(rng.integers(40, 3000)), emission_rate=int(rng.integers(5, 1500)),
        ellipsoid=int(rng.integers(0, 2)),
        size_start_min=u(0.1, 1), size_start_max=u(1, 3), size_end_min=u(0.1, 1), size_end_max=u(1, 5),
        energy_min=float(rng.integers(30, 600)), energy_max=float(rng.integers(600, 2500)),
        color_start=[u(0, 1) for _ in range(4)], color_start_var=[u(0, 0.3) for _ in range(4)],
        color_end=[u(0, 1) for _ in range(4)], color_end_var=[u(0, 0.3) for _ in range(4)],
        position=v3(2), position_var=v3(2), velocity=v3(5), velocity_var=v3(3), acceleration=v3(4), acceleration_var=v3(2),
        rotation_per_particle_speed_min=u(-2, 0), rotation_per_particle_speed_max=u(0, 2),
        orbit_position=int(rng.integers(0, 2)), orbit_velocity=int(rng.integers(0, 2)), orbit_acceleration=int(rng.integers(0, 2)),
        sprite_animated=animated, sprite_looped=int(rng.integers(0, 2)), sprite_frame_random_offset=int(rng.integers(0, 5)),
        sprite_frame_duration=int(rng.integers(10, 120)), texture_width=256.0, texture_height=128.0)
    if rotate:
        p.rotation_speed_min, p.rotation_speed_max = u(-3, 0), u(0, 3)
        p.rotation_axis = abi.F3(*v3(1))
        p.rotation_axis_var = abi.F3(*v3(0.5))
    frame_count = int(rng.integers(1, 9))
    # create(Properties*) clamps the random start frame to the frame count (PE.cpp:128); beyond it the reference reads
    # past its uv table
    p.sprite_frame_random_offset = min(p.sprite_frame_random_offset, frame_count)
    return p, (frame_count, 64, 64)


@pytest.mark.parametrize("seed,rotate", [(s, r) for s in range(10) for r in (False, True)])
def test_random_api_sequences(seed, rotate):
    from gpu_adapter import GpuLib
    rng = np.random.default_rng(1000 + seed)
    O = H.oracle(fresh=True)
    G = GpuLib(rng_mode=abi.RNG_DEVICE)
    try:
        pairs = []
        for k in range(int(rng.integers(2, 6))):
            p, sprite = random_params(rng, rotate)
            eo, eg = O.emitter(p), G.emitter(p)
            for e in (eo, eg):
                e.set_sprite_frame_coords(*sprite)
            O.emitter_set_device_rng(eo.h, 1, 7000 + k)
            eg.pe.set_rng(abi.RNG_DEVICE, 7000 + k)
            pairs.append((eo, eg, p))
        caps = [int(p.particle_count_max) for _, _, p in pairs]
        cam = S.rot_xyz(0.3, -0.2, 0.1, (1, 2, 9))
        for eo, eg, _ in pairs:
            eo.set_camera_world(cam); eg.set_camera_world(cam)

        def check(full):
            for i, (eo, eg, _) in enumerate(pairs):
                assert eg.count() == eo.count(), f"emitter {i}"
                assert eg.is_active() == eo.is_active(), f"emitter {i}"
                if full:
                    a, b = eg.state(), eo.state()
                    for c in abi.INT_COLUMNS:
                        assert np.array_equal(a[c], b[c]), (i, abi.COLUMN_NAMES[c])
                    if rotate:
                        assert np.allclose(a.view(np.float32), b.view(np.float32), rtol=1e-5, atol=1e-6, equal_nan=True), i
                    else:
                        assert np.array_equal(a, b), i

        drawn = False
        for step in range(160):
            op = rng.choice(["update", "update", "update", "draw", "emit", "start", "stop", "rate", "node", "params", "count", "dt_small",
                             "camera", "cap", "frame"])
            k = int(rng.integers(0, len(pairs)))
            eo, eg, p = pairs[k]
            if op == "update":
                dt = float(rng.choice([16.6667, 33.3, 8.0, 50.0]))
                for a, b, _ in pairs:
                    a.update(dt); b.update(dt)
            elif op == "dt_small":
                dt = float(rng.choice([2.5, 3.0, 5.5]))     # below the 8 ms rate cap: the accumulator is shared (PE.cpp:720)
                for a, b, _ in pairs:
                    a.update(dt); b.update(dt)
            elif op == "draw":
                for a, b, _ in pairs:
                    assert a.draw() == b.draw()
                drawn = True
            elif op == "emit":
                n = int(rng.integers(1, 1200))
                eo.emit_once(n); eg.emit_once(n)
            elif op == "start":
                eo.start(); eg.start()
            elif op == "stop":
                eo.stop(); eg.stop()
            elif op == "rate":
                r = int(rng.integers(1, 2000))
                eo.set_emission_rate(r); eg.set_emission_rate(r)
            elif op == "node":
                w = S.rot_xyz(*rng.uniform(-1, 1, 3), tuple(rng.uniform(-3, 3, 3)))
                eo.set_node_world(w); eg.set_node_world(w)
            elif op == "params":
                q = p.copy()
                q.energy_min, q.energy_max = float(rng.integers(20, 300)), float(rng.integers(300, 1500))
                q.sprite_looped = int(rng.integers(0, 2))
                q.velocity = abi.F3(*rng.uniform(-5, 5, 3))
                q.ellipsoid = int(rng.integers(0, 2))
                eo.set_params(q); eg.set_params(q)
                pairs[k] = (eo, eg, q)
            elif op == "count":
                assert eg.count() == eo.count()
            elif op == "camera":
                # the camera moves between calls (possibly between an update and the draw behind it)
                cam = S.rot_xyz(*rng.uniform(-0.5, 0.5, 3), tuple(rng.uniform(-2, 12, 3)))
                for a, b, _ in pairs:
                    a.set_camera_world(cam); b.set_camera_world(cam)
            elif op == "cap":
                # setParticleCountMax (PE.h:237): anywhere between the live count and the size the storage was created with
                alloc = caps[k]
                n_live = eo.count()
                assert eg.count() == n_live
                q = p.copy()
                q.particle_count_max = int(rng.integers(max(n_live, 1), alloc + 1))
                eo.set_params(q); eg.set_params(q)
                pairs[k] = (eo, eg, q)
            elif op == "frame":
                # one whole game-loop frame on every emitter, and the vertices of one of them checked right away
                dt = float(rng.choice([16.6667, 20.0]))
                for a, b, _ in pairs:
                    a.update(dt); b.update(dt)
                for a, b, _ in pairs:
                    assert a.draw() == b.draw()
                drawn = True
                va, vb = eg.vertices(), eo.vertices()
                assert va.shape == vb.shape
                assert np.allclose(va, vb, rtol=1e-5, atol=1e-6) if rotate else np.array_equal(va.view(np.uint32), vb.view(np.uint32))
            if step % 40 == 39:
                check(full=True)
        check(full=True)
        if drawn:
            for a, b, _ in pairs:
                a.draw(); b.draw()
                va, vb = b.vertices(), a.vertices()
                assert va.shape == vb.shape
                assert np.allclose(va, vb, rtol=1e-5, atol=1e-6) if rotate else np.array_equal(va.view(np.uint32), vb.view(np.uint32))
    finally:
        G.close()
