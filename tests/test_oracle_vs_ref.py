"""Pins the plain-C oracle (oracle/particle_oracle.c) against the UNMODIFIED reference compiled headless
(oracle/_ref/libt3dref.so): same parameters, same draw stream, bit-for-bit equal particle state and
vertex stream after every frame.  Needs /root/reference at build time, so these tests skip where the
reference library was never built; tests/test_golden.py covers the same ground from committed vectors.
"""
import ctypes as C
import numpy as np
import pytest

import cpu_harness as H
from tractor3d_b200 import abi, presets

pytestmark = pytest.mark.skipif(not H.ref_available(), reason="oracle/_ref/libt3dref.so not built")

DT = 1000.0 / 60.0


def rot_y(a, t=(0, 0, 0)):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    m[:3, 3] = t
    return m.T.reshape(16).copy()  # column-major


def make_pair(params, sprite=None, seed=1234, fresh=True):
    O, R = H.oracle(fresh=fresh), H.ref(fresh=fresh)
    O.rand_seed(seed)
    R.rand_seed(seed)
    eo, er = O.emitter(params), R.emitter(params)
    if sprite:
        eo.set_sprite_frame_coords(*sprite)
        er.set_sprite_frame_coords(*sprite)
    return O, R, eo, er


def assert_same(eo, er, what=""):
    assert eo.count() == er.count(), what
    so, sr = eo.state(), er.state()
    if not np.array_equal(so, sr):
        bad = np.argwhere(so != sr)
        c, i = bad[0]
        raise AssertionError(f"{what}: state differs in column {abi.COLUMN_NAMES[c]} particle {i}: "
                             f"{so[c, i]:#x} vs {sr[c, i]:#x} ({len(bad)} words)")


def assert_same_vertices(eo, er, what=""):
    vo, vr = eo.vertices(), er.vertices()
    assert vo.shape == vr.shape, what
    assert np.array_equal(vo.view(np.uint32), vr.view(np.uint32)), what


@pytest.mark.parametrize("preset", ["fire", "smoke", "explosion"])
def test_stock_presets_600_frames(preset):
    params, sprite = getattr(presets, preset)()
    O, R, eo, er = make_pair(params, sprite)
    cam = np.eye(4, dtype=np.float32)
    cam[:3, 3] = (0.22, 2.7, 13.3)
    for e in (eo, er):
        e.set_camera_world(cam.T.reshape(16))
        e.start()
    assert np.array_equal(eo.sprite_tex_coords(), er.sprite_tex_coords())
    total = 0
    for f in range(600):
        eo.update(DT); er.update(DT)
        assert eo.draw() == er.draw()
        total += er.count()
        if f % 25 == 0 or f > 590:
            assert_same(eo, er, f"{preset} frame {f}")
            assert_same_vertices(eo, er, f"{preset} frame {f}")
    assert_same(eo, er, preset)
    assert O.rand_calls() == R.rand_calls()
    assert total > 20000


def test_fire_rate_1000_and_rate_1200_quirk():
    # above 1000 particles/s _emitTime is never reduced (PE.cpp:740-743) and the
    # emitter saturates at particleCountMax.
    for rate, lo, hi in ((1000, 800, 950), (1200, 4800, 5000)):
        params, sprite = presets.fire()
        O, R, eo, er = make_pair(params, sprite)
        for e in (eo, er):
            e.set_emission_rate(rate)
            e.start()
        for f in range(600):
            eo.update(DT); er.update(DT)
        assert_same(eo, er, f"rate {rate}")
        assert lo <= er.count() <= hi, er.count()
        assert eo.emit_time() == er.emit_time()


def test_rotation_axis_orbit_and_moving_node():
    params, sprite = presets.fire()
    params.rotation_speed_min, params.rotation_speed_max = -3.0, 3.0
    params.rotation_axis = abi.F3(0.3, 1.0, -0.2)
    params.rotation_axis_var = abi.F3(0.5, 0.5, 0.5)
    params.orbit_acceleration = 1
    params.emission_rate = 600
    O, R, eo, er = make_pair(params, sprite, seed=99)
    for e in (eo, er):
        e.set_camera_world(rot_y(0.4, (1, 2, 3)))
        e.start()
    for f in range(200):
        w = rot_y(0.05 * f, (0.1 * f, 1.0, -0.05 * f))
        eo.set_node_world(w); er.set_node_world(w)
        eo.update(DT); er.update(DT)
        eo.draw(); er.draw()
        if f % 10 == 0:
            assert_same(eo, er, f"frame {f}")
            assert_same_vertices(eo, er, f"frame {f}")
    assert er.count() > 300


def test_burst_until_extinction_swap_with_last_order():
    # removal order (swap-with-last, moved particle skipped that frame).
    params = abi.default_params(particle_count_max=4000, energy_min=50.0, energy_max=400.0,
                                velocity_var=[2, 2, 2], size_end_min=3.0, size_end_max=4.0,
                                rotation_per_particle_speed_min=-1, rotation_per_particle_speed_max=1)
    O, R, eo, er = make_pair(params, None, seed=7)
    eo.emit_once(5000); er.emit_once(5000)  # clamped to 4000, PE.cpp:293-296
    assert eo.count() == er.count() == 4000
    frames = 0
    while er.count() > 0:
        eo.update(DT); er.update(DT)
        assert_same(eo, er, f"frame {frames}")
        frames += 1
        assert frames < 100
    assert eo.count() == 0
    assert not er.is_active() and not eo.is_active()


def test_lowered_particle_count_max_clamps_emit_once():
    # setParticleCountMax (PE.h:237) only rewrites the field the clamp of PE.cpp:293-296 reads; the array keeps its size.
    params = abi.default_params(particle_count_max=1000, energy_min=5000.0, energy_max=6000.0, velocity_var=[1, 1, 1])
    O, R, eo, er = make_pair(params, None, seed=17)
    eo.emit_once(300); er.emit_once(300)
    for e in (eo, er):
        q = e.params(); q.particle_count_max = 450; e.set_params(q)
        assert e.params().particle_count_max == 450
    eo.emit_once(400); er.emit_once(400)
    assert eo.count() == er.count() == 450
    for e in (eo, er):
        q = e.params(); q.particle_count_max = 1000; e.set_params(q)      # back up to the size at create
    eo.emit_once(700); er.emit_once(700)
    assert eo.count() == er.count() == 1000
    for e in (eo, er):
        q = e.params(); q.particle_count_max = 5000; e.set_params(q)      # beyond the array: the harnesses refuse
        assert e.params().particle_count_max == 1000
    for f in range(3):
        eo.update(DT); er.update(DT)
    assert_same(eo, er, "after lowering and raising the cap")


def test_sprite_animation_looped_and_not():
    for looped in (0, 1):
        params, sprite = presets.explosion()
        params.sprite_looped = looped
        params.sprite_frame_duration = 40
        params.emission_rate = 400
        O, R, eo, er = make_pair(params, sprite, seed=5 + looped)
        for e in (eo, er):
            e.start()
        for f in range(150):
            eo.update(DT); er.update(DT)
            eo.draw(); er.draw()
        assert_same(eo, er, f"looped={looped}")
        assert_same_vertices(eo, er)
        frames = er.state()[abi.COL_FRAME]
        assert frames.max() > 3


def test_rate_cap_small_dt_two_emitters_share_accumulator():
    # PE.cpp:720-725: `static double runningTime` is shared by every emitter of the process.
    O, R = H.oracle(fresh=True), H.ref(fresh=True)
    O.rand_seed(3); R.rand_seed(3)
    params, sprite = presets.fire()
    pairs = []
    for k in range(2):
        eo, er = O.emitter(params), R.emitter(params)
        eo.set_sprite_frame_coords(*sprite); er.set_sprite_frame_coords(*sprite)
        eo.start(); er.start()
        pairs.append((eo, er))
    for f in range(400):
        for eo, er in pairs:
            eo.update(3.0); er.update(3.0)
    for eo, er in pairs:
        assert_same(eo, er)
        assert er.count() > 50
    assert pairs[0][1].count() != pairs[1][1].count()


def test_frozen_rotation_is_latched_once_per_process():
    # SB.cpp:339: the billboard rotation matrix is a function-local static.
    params = abi.default_params(particle_count_max=64, rotation_per_particle_speed_min=0.5,
                                rotation_per_particle_speed_max=2.5, position_var=[3, 3, 3])
    O, R, eo, er = make_pair(params, None, seed=11)
    eo.emit_once(32); er.emit_once(32)
    for f in range(3):
        eo.update(DT); er.update(DT)
        eo.draw(); er.draw()
        assert_same_vertices(eo, er, f"frame {f}")
    m = np.zeros(16, dtype=np.float32)
    assert O.get_frozen_rotation(m.ctypes.data_as(H._P(H.C.c_float))) == 1


def test_stop_drains_and_deactivates():
    params, sprite = presets.smoke()
    params.emission_rate = 200
    O, R, eo, er = make_pair(params, sprite, seed=21)
    for e in (eo, er):
        e.start()
    for f in range(60):
        eo.update(DT); er.update(DT)
    for e in (eo, er):
        e.stop()
    for f in range(400):
        eo.update(DT); er.update(DT)
        assert eo.is_active() == er.is_active()
        assert eo.draw() == er.draw()
    assert_same(eo, er)
    assert er.count() == 0


def test_sprite_tex_coord_tables():
    params, _ = presets.fire()
    O, R, eo, er = make_pair(params, None, fresh=False)
    for n, w, h in ((3, 256, 256), (4, 256, 256), (16, 64, 64), (7, 100, 50), (40, 128, 128)):
        eo.set_sprite_frame_coords(n, w, h); er.set_sprite_frame_coords(n, w, h)
        assert np.array_equal(eo.sprite_tex_coords().view(np.uint32), er.sprite_tex_coords().view(np.uint32)), (n, w, h)
    rects = np.array([[0, 0, 10, 20], [13, 7, 100, 50.5]], dtype=np.float32)
    eo.set_sprite_frame_rects(rects); er.set_sprite_frame_rects(rects)
    assert np.array_equal(eo.sprite_tex_coords(), er.sprite_tex_coords())


# ---- the 2-D SpriteBatch overloads and TileSet::draw ------------------------------------------------
def test_sprites_2d_every_overload_matches_the_reference():
    # One SpriteBatch::draw call per record through the reference's own SpriteBatch.cpp (plain, centred, rotated about
    # zero and non-zero pivots, clipped and culled), against the plain-C restatement: bit for bit, including the quads
    # the clip rectangle drops.
    O, R = H.oracle(fresh=True), H.ref(fresh=True)
    for n, seed in ((1, 1), (257, 2), (20000, 3)):
        sp = H.random_sprites(n, seed)
        want = np.zeros((n, 36), dtype=np.float32)
        got = np.zeros((n, 36), dtype=np.float32)
        nr = R.sprites_2d(sp.ctypes.data, n, 512, 256, want.ctypes.data_as(C.POINTER(C.c_float)), n, None)
        no = O.sprites_2d(sp.ctypes.data, n, got.ctypes.data_as(C.POINTER(C.c_float)), n)
        assert no == nr and (n < 100 or 0 < nr < n)          # some sprites lie outside their clip rectangle
        assert np.array_equal(got[:no].view(np.uint32), want[:nr].view(np.uint32))


def test_tileset_draw_matches_the_reference():
    # TileSet::draw (TileSet.cpp:179-236) on a grid with blank cells, node and camera away from the origin, tile sizes
    # that do not add up exactly in binary (the row / column positions are running float sums).
    O, R = H.oracle(fresh=True), H.ref(fresh=True)
    rng = np.random.default_rng(5)
    for rows, cols, tw, th in ((1, 1, 16.0, 16.0), (7, 13, 0.1, 0.3), (64, 96, 33.3, 17.7)):
        src = np.zeros((rows, cols, 2), dtype=np.float32)
        src[..., 0] = rng.integers(0, 8, (rows, cols)) * 32.0
        src[..., 1] = rng.integers(0, 4, (rows, cols)) * 32.0
        src[rng.random((rows, cols)) < 0.3] = -1.0            # blank cells (TileSet.cpp:217)
        h = R.tileset_create(256, 128, tw, th, rows, cols)
        for r in range(rows):
            for c in range(cols):
                if src[r, c, 0] >= 0:
                    R.tileset_set_tile_source(h, c, r, float(src[r, c, 0]), float(src[r, c, 1]))
        node = np.array([12.25, -3.5, 0.75], dtype=np.float32)
        cam = np.array([100.1, 50.7, 9.0], dtype=np.float32)
        color = np.array([0.9, 0.8, 0.7, 0.6], dtype=np.float32)
        opacity = np.float32(0.35)
        n_cells = rows * cols
        want = np.zeros((n_cells, 36), dtype=np.float32)
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        nr = R.tileset_draw(h, fp(node), fp(cam), fp(color), opacity, fp(want), n_cells)
        R.tileset_destroy(h)
        # what TileSet::draw has made of the two translations by line 205, and of colour and opacity at line 226
        pos = np.array([np.float32(0) - cam[0] + node[0], np.float32(0) - cam[1] + node[1], np.float32(0) + node[2]], dtype=np.float32)
        col = color.copy(); col[3] = color[3] * opacity
        got = np.zeros((n_cells, 36), dtype=np.float32)
        no = O.tileset_draw(fp(src), rows, cols, tw, th, 256.0, 128.0, fp(pos), fp(col), fp(got), n_cells)
        assert no == nr == int((src[..., 0] >= 0).sum())
        assert np.array_equal(got[:no].view(np.uint32), want[:nr].view(np.uint32))


def test_font_draw_text_matches_the_reference():
    # Font::drawText(text, x, y, color, size, rightToLeft) through the reference's own Font.cpp, both walks: fonts with fewer glyphs than
    # printable characters, scaled sizes (the floor / truncation paths), character spacing, lines, tabs, control bytes, NULs and
    # bytes >= 128 -- the restatement must leave the same quads, bit for bit.
    O, R = H.oracle(fresh=True), H.ref(fresh=True)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    color = np.array([0.25, 0.5, 0.75, 1.0], dtype=np.float32)
    # (the last case starts the pen beyond 2^24, where a float pen would round: the reference adds the advances in double)
    cases = [(95, 24, 0, 0.0, 1, 3, -17), (95, 24, 31, 0.125, 300, 4, -17), (60, 18, 7, -0.05, 5000, 5, 0), (95, 32, 64, 0.3, 70000, 6, 123456),
             (10, 12, 12, 0.0, 999, 7, -17), (95, 24, 17, 0.0, 4000, 8, (1 << 25) + 1)]
    for glyph_count, font_size, size, spacing, n, seed, x0 in cases:
        g = H.random_font(seed, glyph_count, font_size)
        text = H.random_text(n, seed)
        h = R.font_create(font_size, g.ctypes.data, glyph_count, 512, 512)
        assert h
        R.font_set_character_spacing(h, spacing)
        for rtl in (0, 1):
            # (right to left the walk is over c_str(): it ends at the text's first NUL, which these texts do contain)
            want = np.zeros((n, 36), dtype=np.float32)
            nr = R.font_draw_text(h, text, n, x0, 40, fp(color), size, rtl, fp(want), n, None)
            got = np.zeros((n, 36), dtype=np.float32)
            no = O.font_draw_text(g.ctypes.data, glyph_count, font_size, spacing, text, n, x0, 40, fp(color), size, rtl, fp(got), n)
            assert no == nr and (n < 100 or 0 < nr < n), (glyph_count, n, rtl)
            assert np.array_equal(got[:no].view(np.uint32), want[:nr].view(np.uint32)), (glyph_count, font_size, size, spacing, n, rtl)
            if rtl and n > 100:
                t2 = text.replace(b"\x00", b"\x01")      # the whole text, right to left
                nr = R.font_draw_text(h, t2, n, x0, 40, fp(color), size, 1, fp(want), n, None)
                no = O.font_draw_text(g.ctypes.data, glyph_count, font_size, spacing, t2, n, x0, 40, fp(color), size, 1, fp(got), n)
                assert no == nr > 0
                assert np.array_equal(got[:no].view(np.uint32), want[:nr].view(np.uint32)), (glyph_count, n, "rtl, no NUL")
        R.font_destroy(h)


def test_ui_sprite_records_draw_like_the_reference():
    # ui Sprite::draw (src/ui/Sprite.cpp:495-574) through the reference's own Sprite.cpp on a node under a scene with a camera,
    # against the host helper that turns the same state into ONE t3d_sprite2d record + the oracle's SpriteBatch restatement
    # (the path a scene's sprites take as one t3d_ctx_draw_sprites list): bit for bit
    from tractor3d_b200 import abi
    lib = abi.load()
    O, R = H.oracle(fresh=True), H.ref(fresh=True)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    n = 2000
    ui = H.random_ui_sprites(n, 21)
    rec = np.zeros(n, dtype=abi.SPRITE2D_DTYPE)
    want = np.zeros((n, 36), dtype=np.float32)
    for k in range(n):
        one = ui[k:k + 1]
        assert R.ui_sprite_draw(one.ctypes.data, 512, 256, fp(want[k:k + 1]), 1) == 1
        lib.t3d_host_ui_sprite_record(one.ctypes.data, 512, 256, rec[k:k + 1].ctypes.data)
    assert (rec["flags"] == abi.SPRITE_ROTATED).all() and (rec["rotation_angle"] != 0).mean() > 0.5
    got = np.zeros((n, 36), dtype=np.float32)
    assert O.sprites_2d(rec.ctypes.data, n, fp(got), n) == n
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
