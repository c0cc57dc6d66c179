"""Scenario scripts shared by the CPU tests, the golden-vector generator and the GPU parity tests.

A scenario is a parameter set plus a list of steps; `run()` plays it on anything that looks like
cpu_harness.CpuEmitter (the oracle, the reference, or GpuAdapter over the CUDA library) and returns the
snapshots taken at ("check",) steps.  When `record_lib` is given, the raw draws each step consumed are
recorded so the same stream can be fed to another implementation step by step.
"""
import json

import numpy as np

from tractor3d_b200 import abi, presets

DT = 1000.0 / 60.0


def rot_y(a, t=(0.0, 0.0, 0.0)):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    m[:3, 3] = t
    return np.ascontiguousarray(m.T.reshape(16))  # column-major


def rot_xyz(ax, ay, az, t=(0.0, 0.0, 0.0)):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    m = np.eye(4)
    m[:3, :3] = rz @ ry @ rx
    m[:3, 3] = t
    return np.ascontiguousarray(m.T.reshape(16).astype(np.float32))


class Scenario:
    def __init__(self, name, params, sprite, steps, exact=True, note=""):
        self.name, self.params, self.sprite, self.steps, self.exact, self.note = name, params, sprite, steps, exact, note


def _frames(n, dt=DT, draw=True, check_every=0, tail_checks=2):
    steps = []
    for f in range(n):
        steps.append(("update", dt))
        if draw:
            steps.append(("draw",))
        if (check_every and f % check_every == check_every - 1) or f >= n - tail_checks:
            steps.append(("check",))
    return steps


def fire(frames=120, rate=None):
    p, sprite = presets.fire()
    if rate:
        p.emission_rate = rate
    cam = np.eye(4, dtype=np.float32)
    cam[:3, 3] = (0.22, 2.7, 13.3)  # editor camera of the sample (ParticlesSample.cpp:1102-1115): translation only
    steps = [("camera", cam.T.reshape(16).copy()), ("start",)] + _frames(frames, check_every=30)
    return Scenario("fire" if not rate else f"fire_rate{rate}", p, sprite, steps)


def smoke(frames=150):
    p, sprite = presets.smoke()
    p.emission_rate = 60
    steps = [("start",)] + _frames(frames, check_every=50)
    return Scenario("smoke", p, sprite, steps)


def explosion(frames=100, looped=0):
    p, sprite = presets.explosion()
    p.sprite_looped = looped
    p.emission_rate = 300
    if looped:
        p.sprite_frame_duration = 40
    steps = [("start",)] + _frames(frames, check_every=25)
    return Scenario("explosion_looped" if looped else "explosion", p, sprite, steps)


def burst_extinction(n=1200, cap=1200):
    p = abi.default_params(particle_count_max=cap, energy_min=50.0, energy_max=400.0, velocity_var=[2, 2, 2],
                           size_end_min=3.0, size_end_max=4.0, position_var=[1, 2, 3],
                           rotation_per_particle_speed_min=-1.0, rotation_per_particle_speed_max=1.0)
    steps = [("emit", n + 300), ("check",)]
    for f in range(30):
        steps += [("update", DT), ("draw",)]
        if f % 3 == 2:
            steps.append(("check",))
    return Scenario("burst_extinction", p, None, steps, note="swap-with-last removal order, SURVEY A.4")


def capacity_saturation():
    # rate > 1000/s: _emitTime is never reduced (PE.cpp:740-743) and the emitter pegs at particleCountMax.
    p, sprite = presets.fire()
    p.particle_count_max = 600
    p.emission_rate = 1200
    steps = [("start",)] + _frames(70, check_every=14)
    return Scenario("capacity_saturation", p, sprite, steps)


def axis_rotation(frames=60):
    p, sprite = presets.fire()
    p.rotation_speed_min, p.rotation_speed_max = -3.0, 3.0
    p.rotation_axis = abi.F3(0.3, 1.0, -0.2)
    p.rotation_axis_var = abi.F3(0.5, 0.5, 0.5)
    p.orbit_acceleration = 1
    p.emission_rate = 600
    steps = [("camera", rot_xyz(0.2, 0.4, -0.1, (1, 2, 3))), ("start",)]
    for f in range(frames):
        steps.append(("node", rot_y(0.05 * f, (0.1 * f, 1.0, -0.05 * f))))
        steps += [("update", DT), ("draw",)]
        if f % 15 == 14 or f >= frames - 2:
            steps.append(("check",))
    return Scenario("axis_rotation", p, sprite, steps, exact=False,
                    note="createRotation per particle per frame (PE.cpp:757-763): cosf/sinf differ by ulps on the GPU")


def stop_and_drain():
    p, sprite = presets.smoke()
    p.emission_rate = 200
    p.energy_min, p.energy_max = 300.0, 900.0
    steps = [("start",)] + _frames(40, check_every=20) + [("stop",)] + _frames(70, check_every=10)
    return Scenario("stop_and_drain", p, sprite, steps)


def small_dt_rate_cap():
    p, sprite = presets.fire()
    steps = [("start",)] + _frames(200, dt=3.0, check_every=40)
    return Scenario("small_dt_rate_cap", p, sprite, steps, note="PE.cpp:720-725 accumulator: updates run every third call")


def tilted_camera_moving_node():
    p, sprite = presets.explosion()
    p.emission_rate = 500
    p.orbit_position = p.orbit_velocity = p.orbit_acceleration = 1
    p.acceleration = abi.F3(0.0, -9.8, 0.0)
    steps = [("camera", rot_xyz(0.3, -0.7, 0.2, (4, 5, 6))), ("start",)]
    for f in range(50):
        steps.append(("node", rot_xyz(0.02 * f, 0.03 * f, 0.01 * f, (0.2 * f, 0.0, 1.0))))
        steps += [("update", DT), ("draw",)]
        if f % 10 == 9:
            steps.append(("check",))
    return Scenario("tilted_camera_moving_node", p, sprite, steps)


ALL = [fire, smoke, explosion, lambda: explosion(looped=1), burst_extinction, capacity_saturation, axis_rotation,
       stop_and_drain, small_dt_rate_cap, tilted_camera_moving_node, lambda: fire(frames=90, rate=1000)]


def all_scenarios():
    return [f() for f in ALL]


def by_name(name):
    for s in all_scenarios():
        if s.name == name:
            return s
    raise KeyError(name)


def new_emitter(lib, sc):
    e = lib.emitter(sc.params)
    if sc.sprite:
        e.set_sprite_frame_coords(*sc.sprite)
    return e


def run(e, sc, record_lib=None, feed=None):
    """Plays the scenario.  Returns (snapshots, per_step_draws).
    record_lib: CpuLib whose draw source is recorded step by step.
    feed: per-step draws (from a recorded run) handed to e.feed_draws before each step."""
    snaps, rec = [], []
    drawn = False
    for i, st in enumerate(sc.steps):
        if feed is not None and len(feed[i]):
            e.feed_draws(feed[i])
        if record_lib is not None:
            record_lib.rand_record(1)
        op = st[0]
        if op == "update":
            e.update(st[1])
        elif op == "draw":
            e.draw()
            drawn = True
        elif op == "emit":
            e.emit_once(st[1])
        elif op == "start":
            e.start()
        elif op == "stop":
            e.stop()
        elif op == "node":
            e.set_node_world(st[1])
        elif op == "camera":
            e.set_camera_world(st[1])
        elif op == "rate":
            e.set_emission_rate(st[1])
        elif op == "check":
            snap = {"step": i, "count": e.count(), "state": e.state().copy(), "active": e.is_active()}
            snap["vertices"] = e.vertices().copy() if drawn else np.zeros((0, 4, 9), np.float32)
            snaps.append(snap)
        else:
            raise ValueError(op)
        if record_lib is not None:
            rec.append(record_lib.recorded())
            record_lib.rand_record(0)
    return snaps, rec


def compare(a, b, exact=True, rtol=1e-5, atol=1e-6, what=""):
    """Counts and integer columns always exact; float columns bit-exact, or within rtol/atol when the
    scenario involves device cosf/sinf."""
    assert len(a) == len(b), what
    for sa, sb in zip(a, b):
        w = f"{what} step {sa['step']}"
        assert sa["count"] == sb["count"], f"{w}: live count {sa['count']} vs {sb['count']}"
        assert sa["active"] == sb["active"], w
        A, B = sa["state"], sb["state"]
        assert A.shape == B.shape, w
        for c in abi.INT_COLUMNS:
            assert np.array_equal(A[c], B[c]), f"{w}: integer column {abi.COLUMN_NAMES[c]} differs"
        if exact:
            if not np.array_equal(A, B):
                bad = np.argwhere(A != B)
                c, i = bad[0]
                raise AssertionError(f"{w}: column {abi.COLUMN_NAMES[c]} particle {i}: {A[c, i]:#010x} vs {B[c, i]:#010x} "
                                     f"({len(bad)} words differ)")
        else:
            fa, fb = A.view(np.float32), B.view(np.float32)
            for c in range(abi.NUM_COLUMNS):
                if c in abi.INT_COLUMNS:
                    continue
                assert np.allclose(fa[c], fb[c], rtol=rtol, atol=atol), f"{w}: column {abi.COLUMN_NAMES[c]} beyond tolerance"
        va, vb = sa["vertices"], sb["vertices"]
        assert va.shape == vb.shape, f"{w}: vertex stream {va.shape} vs {vb.shape}"
        if exact:
            assert np.array_equal(va.view(np.uint32), vb.view(np.uint32)), f"{w}: vertex stream differs"
        else:
            assert np.allclose(va, vb, rtol=rtol, atol=atol), f"{w}: vertex stream beyond tolerance"


def params_json(p):
    return json.dumps(p.as_dict())
