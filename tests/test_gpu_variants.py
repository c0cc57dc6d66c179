"""The alternative code paths kept in the library (selected by environment variables read once per process)
must stay parity-green too: the plain-LDG update kernel, ticket-ordered tiles, the blocked struct-of-arrays
layout, the LSU copy-out of the draw kernel.  Each runs a slice of the parity suite in a fresh process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    {"T3D_UPDATE_VARIANT": "v4t128s4b3"},
    {"T3D_UPDATE_VARIANT": "v4t128s4b3", "T3D_UPDATE_TICKET": "1"},
    {"T3D_UPDATE_VARIANT": "a4t64s8b3"},
    {"T3D_LAYOUT_BLOCK": "512"},
    {"T3D_LAYOUT_BLOCK": "256", "T3D_UPDATE_VARIANT": "v2t256s4b4"},
    {"T3D_DRAW_VARIANT": "lsu"},
]


@pytest.mark.parametrize("env", CASES, ids=lambda e: ",".join(f"{k[4:].lower()}={v}" for k, v in e.items()))
def test_alternative_paths_stay_bit_exact(env):
    e = dict(os.environ)
    e.update(env)
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q",
           "-k", "golden or multi_tile or batch_of_mixed or device_rng"]
    res = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
