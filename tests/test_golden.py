"""The plain-C oracle against the committed golden vectors (tests/golden/*.npz, generated from the unmodified
reference by tests/golden/make_golden.py).  Runs anywhere: needs neither /root/reference nor a GPU."""
import numpy as np
import pytest

import cpu_harness as H
import golden_io as G
import scenarios as S


@pytest.mark.parametrize("name", G.names())
def test_oracle_reproduces_reference_golden(name):
    g = G.load(name)
    sc = S.by_name(name)
    assert S.params_json(sc.params) == S.params_json(g["params"]), "scenario definition drifted from the fixture: regenerate"
    assert len(sc.steps) == g["n_steps"]
    O = H.oracle(fresh=True)
    O.replay(g["all_draws"])
    e = S.new_emitter(O, sc)
    snaps, _ = S.run(e, sc)
    assert O.rand_replay_remaining() == 0, "the oracle consumed a different number of draws than the reference"
    S.compare(snaps, g["snaps"], exact=True, what=name)  # CPU vs CPU: bit-exact everywhere, trig included


def test_golden_covers_the_edge_cases():
    names = set(G.names())
    assert {"fire", "smoke", "explosion", "burst_extinction", "capacity_saturation", "axis_rotation",
            "stop_and_drain", "small_dt_rate_cap"} <= names
    g = G.load("stop_and_drain")
    assert g["snaps"][-1]["count"] == 0 and not g["snaps"][-1]["active"]
    g = G.load("capacity_saturation")
    assert g["snaps"][-1]["count"] == g["params"].particle_count_max
    g = G.load("burst_extinction")
    counts = [s["count"] for s in g["snaps"]]
    assert counts[0] == g["params"].particle_count_max and counts[-1] < 50 and sorted(counts, reverse=True) == counts


# ---- 2-D sprites and tile sets, fixtures from the reference's own SpriteBatch.cpp / TileSet.cpp ----
def _sprite_fixture(name):
    import os
    return np.load(os.path.join(G.GOLDEN_DIR, "sprites", name + ".npz"))


def tileset_inputs(z):
    """What TileSet::draw has made of node and camera translation by TileSet.cpp:205 and of colour and opacity at :226."""
    node, cam, color = z["node"], z["camera"], z["color"]
    pos = np.array([np.float32(0) - cam[0] + node[0], np.float32(0) - cam[1] + node[1], np.float32(0) + node[2]], dtype=np.float32)
    col = color.copy()
    col[3] = color[3] * z["opacity"]
    return pos, col


def test_oracle_reproduces_reference_golden_sprites():
    import ctypes as C
    from tractor3d_b200 import abi
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    O = H.oracle(fresh=True)
    z = _sprite_fixture("sprites_2d")
    sp = np.ascontiguousarray(z["sprites"]).view(abi.SPRITE2D_DTYPE).reshape(-1)
    want = z["vertices"]
    got = np.zeros((sp.size, 36), dtype=np.float32)
    n = O.sprites_2d(sp.ctypes.data, sp.size, fp(got), sp.size)
    assert n == want.shape[0] < sp.size
    assert np.array_equal(got[:n].view(np.uint32), want)
    z = _sprite_fixture("font_text")
    g = np.ascontiguousarray(z["glyphs"]).view(abi.GLYPH_DTYPE).reshape(-1)
    text = z["text"].tobytes()
    got = np.zeros((len(text), 36), dtype=np.float32)
    n = O.font_draw_text(g.ctypes.data, g.size, int(z["font_size"]), float(z["spacing"]), text, len(text), int(z["origin"][0]), int(z["origin"][1]),
                         fp(np.ascontiguousarray(z["color"])), int(z["size"]), 0, fp(got), len(text))
    assert n == z["vertices"].shape[0] < len(text)
    assert np.array_equal(got[:n].view(np.uint32), z["vertices"])
    z = _sprite_fixture("ui_sprites")
    ui = np.ascontiguousarray(z["sprites"]).view(abi.UI_SPRITE_DTYPE).reshape(-1)
    rec = np.zeros(ui.size, dtype=abi.SPRITE2D_DTYPE)
    lib = abi.load()
    for k in range(ui.size):
        lib.t3d_host_ui_sprite_record(ui[k:k + 1].ctypes.data, int(z["texture"][0]), int(z["texture"][1]), rec[k:k + 1].ctypes.data)
    got = np.zeros((ui.size, 36), dtype=np.float32)
    assert O.sprites_2d(rec.ctypes.data, ui.size, fp(got), ui.size) == ui.size
    assert np.array_equal(got.view(np.uint32), z["vertices"])
    z = _sprite_fixture("tileset")
    src = np.ascontiguousarray(z["sources"])
    rows, cols = src.shape[:2]
    pos, col = tileset_inputs(z)
    got = np.zeros((rows * cols, 36), dtype=np.float32)
    n = O.tileset_draw(fp(src), rows, cols, float(z["tile"][0]), float(z["tile"][1]), float(z["texture"][0]), float(z["texture"][1]),
                       fp(pos), fp(col), fp(got), rows * cols)
    assert n == z["vertices"].shape[0] == int((src[..., 0] >= 0).sum())
    assert np.array_equal(got[:n].view(np.uint32), z["vertices"])
