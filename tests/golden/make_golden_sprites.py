"""Generates tests/golden/sprites/*.npz from the UNMODIFIED reference (oracle/_ref/libt3dref.so): the vertex streams its
own SpriteBatch.cpp / TileSet.cpp produce for a list of 2-D sprites (every overload: plain, centred, rotated, clipped)
and for a tile set with blank cells; and the quads its Font.cpp lays out for a text.  Run here, where /root/reference exists and `make -C oracle ref` has been done:
    python tests/golden/make_golden_sprites.py
The fixtures travel to machines without the reference (the GPU box)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cpu_harness as H  # noqa: E402

fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))


def main():
    assert H.ref_available(), "build oracle/_ref first: make -C oracle ref"
    R = H.ref(fresh=True)
    out_dir = os.path.join(HERE, "sprites")
    os.makedirs(out_dir, exist_ok=True)

    n = 3000
    sp = H.random_sprites(n, seed=2026)
    verts = np.zeros((n, 36), dtype=np.float32)
    quads = R.sprites_2d(sp.ctypes.data, n, 512, 256, fp(verts), n, None)
    np.savez_compressed(os.path.join(out_dir, "sprites_2d.npz"), sprites=sp.view(np.uint32).reshape(n, -1), texture=np.array([512, 256]),
                        vertices=verts[:quads].view(np.uint32))
    print(f"sprites_2d: {n} sprites -> {quads} quads")

    rows, cols, tw, th = 37, 53, 33.3, 17.7
    rng = np.random.default_rng(7)
    src = np.zeros((rows, cols, 2), dtype=np.float32)
    src[..., 0] = rng.integers(0, 8, (rows, cols)) * 32.0
    src[..., 1] = rng.integers(0, 4, (rows, cols)) * 32.0
    src[rng.random((rows, cols)) < 0.25] = -1.0
    h = R.tileset_create(256, 128, tw, th, rows, cols)
    for r in range(rows):
        for c in range(cols):
            if src[r, c, 0] >= 0:
                R.tileset_set_tile_source(h, c, r, float(src[r, c, 0]), float(src[r, c, 1]))
    node = np.array([12.25, -3.5, 0.75], dtype=np.float32)
    cam = np.array([100.1, 50.7, 9.0], dtype=np.float32)
    color = np.array([0.9, 0.8, 0.7, 0.6], dtype=np.float32)
    opacity = np.float32(0.35)
    verts = np.zeros((rows * cols, 36), dtype=np.float32)
    quads = R.tileset_draw(h, fp(node), fp(cam), fp(color), opacity, fp(verts), rows * cols)
    R.tileset_destroy(h)
    np.savez_compressed(os.path.join(out_dir, "tileset.npz"), sources=src, tile=np.array([tw, th], dtype=np.float32), texture=np.array([256, 128]),
                        node=node, camera=cam, color=color, opacity=opacity, vertices=verts[:quads].view(np.uint32))
    print(f"tileset: {rows} x {cols} cells -> {quads} quads")

    # Font::drawText through the reference's own Font.cpp: a scaled size with character spacing, lines, tabs, control bytes,
    # NULs and bytes >= 128
    glyph_count, font_size, size, spacing, n_text = 95, 24, 31, 0.125, 6000
    g = H.random_font(11, glyph_count, font_size)
    text = H.random_text(n_text, 12)
    color = np.array([0.25, 0.5, 0.75, 1.0], dtype=np.float32)
    h = R.font_create(font_size, g.ctypes.data, glyph_count, 512, 512)
    R.font_set_character_spacing(h, spacing)
    verts = np.zeros((n_text, 36), dtype=np.float32)
    quads = R.font_draw_text(h, text, n_text, -17, 40, fp(color), size, 0, fp(verts), n_text, None)
    R.font_destroy(h)
    np.savez_compressed(os.path.join(out_dir, "font_text.npz"), glyphs=g.view(np.uint32).reshape(glyph_count, -1), font_size=np.array(font_size),
                        size=np.array(size), spacing=np.float32(spacing), text=np.frombuffer(text, dtype=np.uint8), origin=np.array([-17, 40]),
                        color=color, vertices=verts[:quads].view(np.uint32))
    print(f"font_text: {n_text} bytes -> {quads} quads")

    # ui Sprite::draw through the reference's own src/ui/Sprite.cpp: one quad per sprite node
    from tractor3d_b200 import abi
    n_ui = 1500
    ui = H.random_ui_sprites(n_ui, 33)
    verts = np.zeros((n_ui, 36), dtype=np.float32)
    for k in range(n_ui):
        assert R.ui_sprite_draw(ui[k:k + 1].ctypes.data, 512, 256, fp(verts[k:k + 1]), 1) == 1
    np.savez_compressed(os.path.join(out_dir, "ui_sprites.npz"), sprites=ui.view(np.uint32).reshape(n_ui, -1), texture=np.array([512, 256]),
                        vertices=verts.view(np.uint32))
    print(f"ui_sprites: {n_ui} sprite nodes -> {n_ui} quads")


if __name__ == "__main__":
    main()
