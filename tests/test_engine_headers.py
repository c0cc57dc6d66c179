"""The facade built against the ENGINE's own headers.  tests/engine_caller.cpp is application code in the style of
src/scene/SceneLoader.cpp:318-324 and samples/browser/src/ParticlesSample.cpp, written against the engine's headers only;
tractor3d_b200.build.build_engine_callers() compiles it once against the reference's class and once, with
-DT3D_WITH_ENGINE_HEADERS, against tractor3d_b200/host (the engine's math / Ref / Properties objects linked in).
CPU: both programs build, and the reference one runs.  GPU: both run and must print the same text."""
import os
import subprocess

import pytest

from test_abi_and_host import PARTICLE_TEXT, _write_png
from tractor3d_b200 import build


def _inputs(tmp_path):
    _write_png(str(tmp_path / "sheet.png"), 96, 32)
    _write_png(str(tmp_path / "other.png"), 128, 64)
    f = tmp_path / "sparks.particle"
    f.write_text(PARTICLE_TEXT)
    return [str(f), str(tmp_path / "other.png")]


def _run(exe, args, cwd):
    return subprocess.run([exe] + args, capture_output=True, text=True, timeout=120, cwd=cwd)


@pytest.mark.skipif(not os.path.isdir(os.path.join(build.REFERENCE, "Tractor3Dlib", "include")), reason="needs the reference tree")
def test_facade_builds_against_the_engine_headers(tmp_path):
    built = build.build_engine_callers()
    assert built is not None
    ref, b200 = built
    assert os.access(ref, os.X_OK) and os.access(b200, os.X_OK)
    # the same source really is valid application code for the reference: it runs against the unmodified class
    res = _run(ref, _inputs(tmp_path), str(tmp_path))
    assert res.returncode == 0, res.stderr
    out = res.stdout
    assert "[loaded] sprite 32x16 frames 6 animated 1 looped 1 offset 6 duration 33 blend BLEND_MULTIPLIED" in out
    assert "[loaded] texture 96x32 started 0 active 0 count 0" in out
    assert "clamped: count 60 max 60" in out                       # a lowered particleCountMax clamps emitOnce (PE.cpp:293-296)
    assert "30 frames: count 0 active 0 draw 0" in out
    assert "[retextured] sprite 128x64 frames 1" in out            # setTexture resets the sprite table (PE.cpp:249-251)
    # the facade's binary imports nothing of the reference's particle path: only the C ABI
    syms = subprocess.run(["nm", "-D", "--undefined-only", b200], capture_output=True, text=True).stdout
    assert "t3d_emitter_update" in syms and "SpriteBatch" not in syms


@pytest.mark.gpu
def test_engine_header_build_behaves_like_the_reference(tmp_path):
    ref, b200 = build.engine_caller_paths()
    if not (os.path.exists(ref) and os.path.exists(b200)):
        pytest.skip("the engine-header programs were not built (they need the reference tree at build time)")
    args = _inputs(tmp_path)
    want = _run(ref, args, str(tmp_path))
    got = _run(b200, args, str(tmp_path))
    assert want.returncode == 0, want.stderr
    assert got.returncode == 0, got.stdout + got.stderr
    assert got.stdout == want.stdout
