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


def test_cpp_facade_600_frames_through_create_properties(tmp_path):
    # BASELINE configs[0] in the facade's terms: a .particle description handed over as a Properties namespace the way
    # SceneLoader.cpp:318-324 does it (`props` mode of the driver: Properties::create, getNextNamespace,
    # ParticleEmitter::create(Properties*)), 600 frames of update + draw, against the oracle.
    driver = build.build_facade_driver()
    _write_png(str(tmp_path / "sheet.png"), 96, 32)
    f = tmp_path / "sparks.particle"
    # no axis rotation here (rotationSpeed keys dropped): nothing calls sincosf, so the comparison is bit for bit
    text = "\n".join(l for l in PARTICLE_TEXT.splitlines() if "rotationSpeed" not in l and "rotationAxis" not in l)
    f.write_text(text.replace("emissionRate = 55", "emissionRate = 300"))
    frames, seed = 600, 99
    out = tmp_path / "out.bin"
    res = subprocess.run([driver, str(f), str(frames), str(seed), str(out), "0", "props"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = np.fromfile(out, dtype=np.uint32)
    count, quads, draws, frame_count = [int(x) for x in raw[:4]]
    state = raw[4:4 + abi.NUM_COLUMNS * count].reshape(abi.NUM_COLUMNS, count)
    verts = raw[4 + abi.NUM_COLUMNS * count:].reshape(quads, 4, 9)

    lib = abi.load()
    rc, p, fc, sw, sh, _ = _parse(lib, str(f))
    assert rc == abi.OK and fc == frame_count
    O = H.oracle(fresh=True)
    eo = O.emitter(p)
    eo.set_sprite_frame_coords(fc, sw, sh)
    O.emitter_set_device_rng(eo.h, 1, seed)
    cam = np.eye(4, dtype=np.float32); cam[:3, 3] = (0.22, 2.7, 13.3)
    eo.set_camera_world(cam.T.reshape(16))
    eo.start()
    n_draws = 0
    for fr in range(frames):
        w = np.eye(4, dtype=np.float32); w[0, 3] = np.float32(0.05) * np.float32(fr)
        eo.set_node_world(w.T.reshape(16))
        eo.update(S.DT)
        n_draws += eo.draw()
    assert (count, draws) == (eo.count(), n_draws) and count > 100
    assert np.array_equal(state, eo.state())
    assert quads == count and np.array_equal(verts, eo.vertices().view(np.uint32))


def test_set_texture_resets_the_sprite_table_and_takes_the_new_size():
    # PE.cpp:229-252: setTexture re-derives the texel ratios from the new texture and goes back to ONE frame covering it,
    # whether or not the size changed; frames cut afterwards use the new size (the uv table is pinned to the reference's
    # on the CPU: test_sprite_frame_coords_follow_the_oracle).
    from gpu_adapter import GpuLib
    from tractor3d_b200 import presets
    lib = GpuLib(rng_mode=abi.RNG_DEVICE, seed=5)
    try:
        p, sprite = presets.headline_64m(capacity=3000)       # 1024 x 256 texture, four 256 x 256 frames
        p.emission_rate = 1
        O = H.oracle(fresh=True)
        eo = O.emitter(p); eo.set_sprite_frame_coords(*sprite)
        O.emitter_set_device_rng(eo.h, 1, 5)
        eg = lib.emitter(p); eg.set_sprite_frame_coords(*sprite)
        assert eg.pe.getSpriteFrameCount() == 4
        eo.emit_once(2500); eg.emit_once(2500)
        eo.start(); eg.start()
        for _ in range(3):
            eo.update(S.DT); eo.draw(); eg.update(S.DT); eg.draw()
        assert np.array_equal(eg.vertices().view(np.uint32), eo.vertices().view(np.uint32))
        # same size again: the table is reset all the same
        eg.pe.setTextureSize(1024, 256, abi.BLEND_ADDITIVE)
        assert eg.pe.getSpriteFrameCount() == 1 and np.array_equal(eg.pe.sprite_tex_coords(), np.array([[0, 1, 1, 0]], dtype=np.float32))
        assert eg.pe.params().blend_mode == abi.BLEND_ADDITIVE
        # a texture of another size, then a 3 x 2 sheet cut from it
        eg.pe.setTextureSize(384, 128, abi.BLEND_ALPHA)
        q = eg.pe.params()
        assert (q.texture_width, q.texture_height) == (384.0, 128.0)
        eg.pe.setSpriteFrameCoords(6, 128, 64)
        want = np.zeros(24, dtype=np.float32)
        assert abi.load().t3d_host_sprite_frame_coords(6, 128, 64, 384.0, 128.0, want.ctypes.data_as(C.POINTER(C.c_float))) == 6
        assert np.array_equal(eg.pe.sprite_tex_coords().reshape(-1), want)
        # the particles kept their frame indices (0..3): the next frames draw with the new table
        eo.set_sprite_tex_coords(want)
        for _ in range(3):
            eo.update(S.DT); eo.draw(); eg.update(S.DT); eg.draw()
        assert np.array_equal(eg.vertices().view(np.uint32), eo.vertices().view(np.uint32))
        with pytest.raises(Exception):
            eg.pe.setTextureSize(0, 128, abi.BLEND_ALPHA)
    finally:
        lib.close()


def test_a_lowered_particle_count_max_clamps_the_spawn():
    # PE.h:237 rewrites the field; PE.cpp:293-296 clamps against it.  The storage keeps the size it was created with, so
    # the value may go down and back up to that size, and never beyond.
    from gpu_adapter import GpuLib
    from tractor3d_b200 import presets
    from tractor3d_b200.emitter import T3DError
    lib = GpuLib(rng_mode=abi.RNG_DEVICE, seed=11)
    try:
        p, sprite = presets.headline_64m(capacity=1000)
        p.emission_rate = 1
        O = H.oracle(fresh=True)
        eo = O.emitter(p); O.emitter_set_device_rng(eo.h, 1, 11)
        eg = lib.emitter(p)
        eo.emit_once(300); eg.emit_once(300)
        q = eo.params(); q.particle_count_max = 450; eo.set_params(q)
        eg.pe.setParticleCountMax(450)
        assert eg.pe.getParticleCountMax() == 450
        eo.emit_once(400); eg.emit_once(400)
        assert eg.count() == eo.count() == 450
        for _ in range(3):
            eo.update(S.DT); eg.update(S.DT)
        assert np.array_equal(eg.state(), eo.state())
        # below the live count: nothing spawns, nothing is lost (the reference's unsigned `max - count` would overrun here)
        eg.pe.setParticleCountMax(100)
        eg.emit_once(50); eg.update(S.DT); eo.update(S.DT)
        assert eg.count() == eo.count()
        assert np.array_equal(eg.state(), eo.state())
        eg.pe.setParticleCountMax(1000)                        # back up to the size at create
        q.particle_count_max = 1000; eo.set_params(q)
        eo.emit_once(700); eg.emit_once(700)
        assert eg.count() == eo.count() == 1000
        with pytest.raises(T3DError) as err:
            eg.pe.setParticleCountMax(1001)
        assert err.value.code == abi.ERR_CAPACITY
    finally:
        lib.close()


def test_create_from_a_hand_built_properties_tree(tmp_path):
    # The C ABI route a caller with its own parsed scene file takes (SceneLoader.cpp:318-324): build the namespace, hand it over.
    from tractor3d_b200.emitter import Context, ParticleEmitter
    lib = abi.load()
    _write_png(str(tmp_path / "flame.png"), 64, 64)
    tree = C.c_void_p()
    assert lib.t3d_properties_new(b"particle", b"flame", C.byref(tree)) == abi.OK
    ctx = Context(0)
    ctx.reset_statics()
    try:
        sprite = C.c_void_p()
        assert lib.t3d_properties_add_namespace(tree, b"sprite", None, C.byref(sprite)) == abi.OK
        for k, v in (("path", str(tmp_path / "flame.png")), ("width", "32"), ("height", "32"), ("frameCount", "4"), ("animated", "true"),
                     ("looped", "true"), ("frameDuration", "40"), ("blendMode", "ADDITIVE")):
            assert lib.t3d_properties_set(sprite, k.encode(), v.encode()) == abi.OK
        for k, v in (("particleCountMax", "2000"), ("emissionRate", "400"), ("energyMin", "300"), ("energyMax", "900"), ("sizeStartMin", "0.5"),
                     ("sizeStartMax", "1.5"), ("sizeEndMin", "0.1"), ("sizeEndMax", "0.2"), ("colorStart", "1, 0.6, 0.1, 1"),
                     ("colorEnd", "0.2, 0, 0, 0"), ("velocity", "0, 2, 0"), ("velocityVar", "0.4, 0.5, 0.4"), ("acceleration", "0, -1, 0")):
            assert lib.t3d_properties_set(tree, k.encode(), v.encode()) == abi.OK
        e = ParticleEmitter.create_from_properties(ctx, tree, (abi.RNG_DEVICE, 21))
        p = e.params()
        assert (p.particle_count_max, p.emission_rate, p.blend_mode, p.sprite_frame_duration) == (2000, 400, abi.BLEND_ADDITIVE, 40)
        assert e.getSpriteFrameCount() == 4
        O = H.oracle(fresh=True)
        eo = O.emitter(p); eo.set_sprite_frame_coords(4, 32, 32)
        O.emitter_set_device_rng(eo.h, 1, 21)
        eo.start(); e.start()
        for _ in range(120):
            eo.update(S.DT); eo.draw(); e.update(S.DT); e.draw()
        assert e.getParticlesCount() == eo.count() > 200
        assert np.array_equal(e.state(), eo.state())
        assert np.array_equal(e.vertices().view(np.uint32), eo.vertices().view(np.uint32))
        # a tree that is not a `particle` namespace is refused like PE.cpp:94-98
        with pytest.raises(Exception):
            ParticleEmitter.create_from_properties(ctx, sprite)
    finally:
        lib.t3d_properties_free(tree)
        ctx.close()


def test_short_lived_emitters_reuse_their_storage():
    # A game creates and destroys explosions next to long-lived emitters: the segment of a destroyed emitter goes back
    # to its slab and is handed out again (it used to be lost until the whole slab was empty), and an emitter
    # that lands in recycled storage starts clean -- same particles as its oracle twin.
    import torch
    from gpu_adapter import GpuLib
    from tractor3d_b200 import presets
    lib = GpuLib(rng_mode=abi.RNG_DEVICE, seed=3)
    try:
        keep_p, keep_sprite = presets.fire()
        keep = lib.emitter(keep_p); keep.set_sprite_frame_coords(*keep_sprite)
        keep.start()
        # (the billboard rotation is latched once per process from the first rotated quad: give the long-running context
        #  and every fresh oracle the same one, so that the vertex streams are comparable)
        axis, angle = np.array([0.0, 0.0, 1.0], dtype=np.float32), 0.3
        lib.ctx.set_frozen_rotation(axis, angle)

        def burst(k, check):
            p, sprite = presets.explosion()
            p.particle_count_max = 50_000
            eg = lib.emitter(p); eg.set_sprite_frame_coords(*sprite)
            eg.pe.set_rng(abi.RNG_DEVICE, 500 + k)
            eg.emit_once(40_000)
            eo = None
            if check:
                O = H.oracle(fresh=True)
                O.set_frozen_rotation(axis.ctypes.data_as(C.POINTER(C.c_float)), angle)
                eo = O.emitter(p); eo.set_sprite_frame_coords(*sprite)
                O.emitter_set_device_rng(eo.h, 1, 500 + k)
                eo.emit_once(40_000)
            for _ in range(3):
                keep.update(S.DT); keep.draw()
                eg.update(S.DT); eg.draw()
                if eo:
                    eo.update(S.DT); eo.draw()
            if eo:
                assert eg.count() == eo.count()
                assert np.array_equal(eg.state(), eo.state())
                assert np.array_equal(eg.vertices().view(np.uint32), eo.vertices().view(np.uint32))
            eg.close()

        for k in range(4):
            burst(k, check=False)        # slabs and vertex storage reach their steady size
        lib.ctx.sync()
        free0 = torch.cuda.mem_get_info()[0]
        for k in range(4, 64):
            burst(k, check=(k % 20 == 5))
        lib.ctx.sync()
        free1 = torch.cuda.mem_get_info()[0]
        assert free0 - free1 < 32 * 1024 * 1024, (free0, free1)     # 60 more explosions of 7 MB each: nothing leaked
        assert keep.count() > 0
    finally:
        lib.close()
