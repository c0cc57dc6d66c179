"""The alternative code paths kept in the library (selected by environment variables read once per process)
must stay parity-green too: other tile shapes of the update kernel (2 or 4 particles per thread), ticket-ordered
tiles, the LSU copy-out of the draw kernel, the fused update+draw launch on and off.  Each runs a slice of the parity suite
in a fresh process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    {"T3D_UPDATE_VARIANT": "u2t256s8b2"},
    {"T3D_UPDATE_VARIANT": "u2t192s12b2", "T3D_UPDATE_TICKET": "1"},
    {"T3D_FUSED": "0", "T3D_UPDATE_TICKET": "1"},
    {"T3D_UPDATE_VARIANT": "u4t128s8b2"},
    {"T3D_DRAW_VARIANT": "lsu"},
    {"T3D_FUSED": "0"},
    {"T3D_FUSED": "1"},
    {"T3D_FUSED": "1", "T3D_UPDATE_TICKET": "1"},
    {"T3D_FUSED": "1", "T3D_FUSED_VARIANT": "f2t192s6b2"},
    {"T3D_FUSED": "1", "T3D_FUSED_VARIANT": "f2t128s12b3"},
]


@pytest.mark.parametrize("env", CASES, ids=lambda e: ",".join(f"{k[4:].lower()}={v}" for k, v in e.items()))
def test_alternative_paths_stay_bit_exact(env):
    e = dict(os.environ)
    e.update(env)
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q",
           "-k", "golden or multi_tile or batch_of_mixed or device_rng or churn_at_scale or headline_parameters"]
    res = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]


FUSED_PROBE = """
import sys
import numpy as np
from tractor3d_b200 import abi, presets
from tractor3d_b200.emitter import Context, ParticleEmitter
n_emitters, per, emit_per_frame = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = Context(0)
p, sprite = presets.fire()
p.particle_count_max = per + 6 * emit_per_frame
p.energy_min, p.energy_max = 1.0e5, 2.0e5
es = [ParticleEmitter.create(ctx, p, sprite, (abi.RNG_DEVICE, 5 + k)) for k in range(n_emitters)]
ctx.batch_emit_once(es, [per] * n_emitters)
for _ in range(6):
    if emit_per_frame:
        ctx.batch_emit_once(es, [emit_per_frame] * n_emitters)
    ctx.batch_update(es, 16.0)
    ctx.batch_draw(es)
ctx.sync()
print("FUSED_LAUNCHES", ctx.kernel_time(3)[1], "SEPARATE", ctx.kernel_time(1)[1], ctx.kernel_time(2)[1])
"""


def _probe(fused, n_emitters, per, emit_per_frame=0):
    e = dict(os.environ)
    e.pop("T3D_FUSED", None)
    if fused is not None:
        e["T3D_FUSED"] = fused
    res = subprocess.run([sys.executable, "-c", FUSED_PROBE, str(n_emitters), str(per), str(emit_per_frame)], env=e,
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("FUSED_LAUNCHES")][0].split()
    return int(line[1]), int(line[3]), int(line[4])


def test_fused_launch_policy():
    """update(set) directly followed by draw(same set): one launch with T3D_FUSED=1, two with T3D_FUSED=0; left alone,
    the runtime fuses (t3d_host_fuse_policy: the fused frame wins at every measured churn level)."""
    # (the very first frame stays apart: the rotation latch has to look at the updated angles before the draw, SB.cpp:339)
    f, u, d = _probe("1", 3, 1000)
    assert f >= 5 and u == 6 - f and d == 6 - f, (f, u, d)
    f, u, d = _probe("0", 3, 1000)
    assert f == 0 and u == 6 and d == 6, (f, u, d)
    f, u, d = _probe(None, 3, 1000)
    assert f >= 5 and u == 6 - f and d == 6 - f, (f, u, d)
    f, u, d = _probe(None, 3, 1000, emit_per_frame=20)    # 2 % of the particles replaced per frame: still fused
    assert f >= 5 and u == 6 - f and d == 6 - f, (f, u, d)


def test_fuzz_with_forced_fusion():
    """The randomised API sequences again with every eligible update+draw pair fused (the emitters there churn too much
    for the automatic policy to pick the fused launch often)."""
    e = dict(os.environ)
    e["T3D_FUSED"] = "1"
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_fuzz.py"), os.path.join(ROOT, "tests", "test_gpu_facade.py"),
           "-x", "-q"]
    res = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]


FAULT_PROBE = """
from tractor3d_b200 import abi, presets
from tractor3d_b200.emitter import Context, ParticleEmitter, T3DError
ctx = Context(0)
p, sprite = presets.headline_64m(capacity=30000)
e = ParticleEmitter.create(ctx, p, sprite, (abi.RNG_DEVICE, 9))
e.emitOnce(20000)
e.update(16.0)
try:
    n = e.getParticlesCount()
    print("NO_FAULT", n)
except T3DError as err:
    print("FAULT", err.code, str(err)[:120])
"""


def test_a_grid_smaller_than_the_live_count_is_reported_not_ignored():
    """The kernels check that the tiles the host launched cover the live count they read; with the test hook that drops
    the last tile of every emitter, the next call that talks to the device must fail with T3D_ERR_INTERNAL."""
    for shrink, expect in (("1", "FAULT 7"), ("0", "NO_FAULT 20000")):
        e = dict(os.environ)
        e["T3D_DEBUG_SHRINK_GRID"] = shrink
        res = subprocess.run([sys.executable, "-c", FAULT_PROBE], env=e, capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        assert expect in res.stdout, res.stdout[-500:]
