"""the other producers of SpriteBatch quads on the device -- the 2-D SpriteBatch::draw overloads as a
batched expansion (t3d_ctx_draw_sprites) and TileSet::draw from a device-resident tile table (t3d_tileset_*) -- against
the plain-C oracle, which tests/test_oracle_vs_ref.py pins bit for bit to the reference's own SpriteBatch.cpp /
TileSet.cpp.  Bit-exact, except sprites turned by a non-zero angle: their sine and cosine come from the device's
sincosf instead of glibc's (tolerance 1e-5 relative to the coordinate, stated below)."""
import ctypes as C

import numpy as np
import pytest

import cpu_harness as H
from tractor3d_b200 import abi
from tractor3d_b200.emitter import Context, Font, TileSet

pytestmark = pytest.mark.gpu
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))


def _oracle_sprites(O, sp):
    out = np.zeros((max(sp.size, 1), 4, 9), dtype=np.float32)
    n = O.sprites_2d(sp.ctypes.data, sp.size, fp(out), sp.size)
    return out[:n]


def _compare(got, want, sp_rotated_somewhere):
    assert got.shape == want.shape
    if not sp_rotated_somewhere:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        return
    # everything but the positions of turned quads is bit-exact; those agree to 1e-5 of the coordinate (sincosf)
    assert np.array_equal(got[:, :, 2:].view(np.uint32), want[:, :, 2:].view(np.uint32))
    assert np.allclose(got[:, :, :2], want[:, :, :2], rtol=1e-5, atol=1e-3)
    same = (got[:, :, :2].view(np.uint32) == want[:, :, :2].view(np.uint32)).all(axis=(1, 2))
    assert same.mean() > 0.3           # the unrotated overloads and angle 0 are exact


@pytest.mark.parametrize("n", [1, 255, 256, 257, 5000, 200_003])
def test_sprite_lists_expand_like_spritebatch(n):
    O = H.oracle(fresh=True)
    ctx = Context(0)
    try:
        # (a) no clipping, no rotation: sprite i is quad i, one launch, bit for bit
        sp = H.random_sprites(n, seed=n, clipped=False, rotated=False)
        ptr, nq = ctx.draw_sprites(sp)
        assert nq == n
        _compare(ctx.download_quads(ptr, nq), _oracle_sprites(O, sp), False)
        # (b) every overload, culled sprites leave no gap
        sp = H.random_sprites(n, seed=n + 1)
        want = _oracle_sprites(O, sp)
        ptr, nq = ctx.draw_sprites(sp)
        assert nq == want.shape[0] and (n < 100 or nq < n)
        _compare(ctx.download_quads(ptr, nq), want, True)
        # (c) clipping only: the compaction path, bit for bit
        sp = H.random_sprites(n, seed=n + 2, rotated=False)
        want = _oracle_sprites(O, sp)
        ptr, nq = ctx.draw_sprites(sp)
        assert nq == want.shape[0]
        _compare(ctx.download_quads(ptr, nq), want, False)
    finally:
        ctx.close()


def test_sprites_written_straight_into_the_mapped_staging_buffer():
    # t3d_ctx_map_sprites: the producer fills pinned memory, the library uploads from where the list lies; two buffers
    # alternate, so consecutive lists do not wait for each other
    O = H.oracle(fresh=True)
    ctx = Context(0)
    try:
        for k in range(5):
            n = 10_000 + 1000 * k
            sp = H.random_sprites(n, seed=300 + k, rotated=False, clipped=(k % 2 == 1))
            mapped = ctx.map_sprites(n)
            mapped[:] = sp
            h2d0 = ctx.transfer_bytes()[0]
            ptr, nq = ctx.draw_sprites(mapped, no_clip=(k % 2 == 0))
            assert ctx.transfer_bytes()[0] - h2d0 == n * 84
            want = _oracle_sprites(O, sp)
            assert nq == want.shape[0]
            assert np.array_equal(ctx.download_quads(ptr, nq).view(np.uint32), want.view(np.uint32)), k
    finally:
        ctx.close()


def test_sprite_list_resident_on_the_device_into_a_foreign_buffer():
    import torch
    O = H.oracle(fresh=True)
    ctx = Context(0)
    try:
        n = 30_000
        sp = H.random_sprites(n, seed=77, rotated=False)
        want = _oracle_sprites(O, sp)
        dev_list = torch.from_numpy(sp.view(np.uint8).copy()).cuda()
        dst = torch.full(((n + 4) * 36,), float("nan"), dtype=torch.float32, device="cuda")
        ptr, nq = ctx.draw_sprites(dev_list.data_ptr(), device_dst=dst.data_ptr(), dst_quads=n, on_device=True, n=n)
        assert ptr == dst.data_ptr() and nq == want.shape[0]
        got = dst.cpu().numpy()
        assert np.array_equal(got[:nq * 36].view(np.uint32), want.reshape(-1).view(np.uint32))
        assert np.isnan(got[nq * 36:]).all()                   # nothing past the list's quads
        with pytest.raises(Exception):
            ctx.draw_sprites(dev_list.data_ptr(), device_dst=dst.data_ptr(), dst_quads=n - 1, on_device=True, n=n)
    finally:
        ctx.close()


@pytest.mark.parametrize("rows,cols,tw,th", [(1, 1, 16.0, 16.0), (7, 13, 0.1, 0.3), (64, 96, 33.3, 17.7), (600, 700, 8.0, 8.0)])
def test_tileset_draw_matches_the_oracle(rows, cols, tw, th):
    O = H.oracle(fresh=True)
    ctx = Context(0)
    try:
        rng = np.random.default_rng(rows * 1000 + cols)
        src = np.zeros((rows, cols, 2), dtype=np.float32)
        src[..., 0] = rng.integers(0, 8, (rows, cols)) * 32.0
        src[..., 1] = rng.integers(0, 4, (rows, cols)) * 32.0
        src[rng.random((rows, cols)) < 0.3] = -1.0
        ts = TileSet(ctx, 256, 128, tw, th, rows, cols)
        assert ts.draw((0, 0, 0), (1, 1, 1, 1))[1] == 0        # TileSet::create leaves every cell blank (TileSet.cpp:41-42)
        ts.set_tiles(src)
        color = np.array([0.9, 0.8, 0.7, 0.21], dtype=np.float32)
        for frame in range(3):                                  # the table stays on the device; only the position travels
            pos = np.array([12.25 - 3.3 * frame, -3.5 + 0.7 * frame, 0.75], dtype=np.float32)
            want = np.zeros((rows * cols, 4, 9), dtype=np.float32)
            n = O.tileset_draw(fp(src), rows, cols, tw, th, 256.0, 128.0, fp(pos), fp(color), fp(want), rows * cols)
            ptr, nq = ts.draw(pos, color)
            assert nq == n == int((src[..., 0] >= 0).sum())
            assert np.array_equal(ctx.download_quads(ptr, nq).view(np.uint32), want[:n].view(np.uint32)), frame
        # one cell changed through the per-cell setter (TileSet.cpp:155-165): a blank one comes alive
        blank = np.argwhere(src[..., 0] < 0)
        if len(blank):
            r, c = blank[0]
            ts.setTileSource(int(c), int(r), (64.0, 32.0)); src[r, c] = (64.0, 32.0)
            want = np.zeros((rows * cols, 4, 9), dtype=np.float32)
            n = O.tileset_draw(fp(src), rows, cols, tw, th, 256.0, 128.0, fp(pos), fp(color), fp(want), rows * cols)
            ptr, nq = ts.draw(pos, color)
            assert nq == n
            assert np.array_equal(ctx.download_quads(ptr, nq).view(np.uint32), want[:n].view(np.uint32))
        ts.release()
    finally:
        ctx.close()


def test_sprites_and_tiles_against_the_reference_fixtures():
    # the committed vectors the reference's own SpriteBatch.cpp / TileSet.cpp / Font.cpp produced (tests/golden/make_golden_sprites.py)
    import os
    import golden_io as G
    from test_golden import tileset_inputs
    ctx = Context(0)
    try:
        z = np.load(os.path.join(G.GOLDEN_DIR, "sprites", "sprites_2d.npz"))
        sp = np.ascontiguousarray(z["sprites"]).view(abi.SPRITE2D_DTYPE).reshape(-1)
        want = z["vertices"].view(np.float32).reshape(-1, 4, 9)
        ptr, nq = ctx.draw_sprites(sp)
        assert nq == want.shape[0]
        _compare(ctx.download_quads(ptr, nq), want, True)
        z = np.load(os.path.join(G.GOLDEN_DIR, "sprites", "tileset.npz"))
        src = np.ascontiguousarray(z["sources"])
        rows, cols = src.shape[:2]
        ts = TileSet(ctx, int(z["texture"][0]), int(z["texture"][1]), float(z["tile"][0]), float(z["tile"][1]), rows, cols)
        ts.set_tiles(src)
        pos, col = tileset_inputs(z)
        ptr, nq = ts.draw(pos, col)
        assert nq == z["vertices"].shape[0]
        assert np.array_equal(ctx.download_quads(ptr, nq).view(np.uint32).reshape(nq, 36), z["vertices"])
        ts.release()
        # a scene's ui Sprite nodes as ONE list: the host helper makes the records, the expansion kernel the quads
        z = np.load(os.path.join(G.GOLDEN_DIR, "sprites", "ui_sprites.npz"))
        ui = np.ascontiguousarray(z["sprites"]).view(abi.UI_SPRITE_DTYPE).reshape(-1)
        rec = np.zeros(ui.size, dtype=abi.SPRITE2D_DTYPE)
        for k in range(ui.size):
            ctx.lib.t3d_host_ui_sprite_record(ui[k:k + 1].ctypes.data, int(z["texture"][0]), int(z["texture"][1]), rec[k:k + 1].ctypes.data)
        ptr, nq = ctx.draw_sprites(rec, no_clip=True)
        assert nq == ui.size
        _compare(ctx.download_quads(ptr, nq), z["vertices"].view(np.float32).reshape(-1, 4, 9), True)
        z = np.load(os.path.join(G.GOLDEN_DIR, "sprites", "font_text.npz"))
        font = Font(ctx, int(z["font_size"]), np.ascontiguousarray(z["glyphs"]).view(abi.GLYPH_DTYPE).reshape(-1))
        font.setCharacterSpacing(float(z["spacing"]))
        ptr, nq = font.drawText(z["text"].tobytes(), int(z["origin"][0]), int(z["origin"][1]), z["color"], int(z["size"]))
        assert nq == z["vertices"].shape[0]
        assert np.array_equal(ctx.download_quads(ptr, nq).view(np.uint32).reshape(nq, 36), z["vertices"])
        font.release()
    finally:
        ctx.close()


# ---- Font::drawText (Font.cpp:242-390) ----
def _oracle_text(O, g, font_size, spacing, text, x, y, color, size, rtl=False):
    out = np.zeros((max(len(text), 1), 4, 9), dtype=np.float32)
    n = O.font_draw_text(g.ctypes.data, g.size, font_size, spacing, text, len(text), x, y, fp(color), size, 1 if rtl else 0, fp(out), len(text))
    return out[:n]


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 127, 128, 129, 255, 256, 257, 4095, 4096, 4097, 70_001, 1_300_003])
def test_text_is_laid_out_like_font_drawtext(n):
    # the pen walk as a segmented scan (tile and block summaries, one-CTA scan, expansion of 128-character tiles): texts that end inside a
    # 16-byte load, a tile, a summary CTA (4 096 characters), and over a thousand tiles; bit for bit
    O = H.oracle(fresh=True)
    ctx = Context(0)
    try:
        color = np.array([0.25, 0.5, 0.75, 1.0], dtype=np.float32)
        for glyph_count, font_size, size, spacing, x0 in ((95, 24, 0, 0.0, -17), (60, 18, 31, 0.125, 5), (95, 32, 20, -0.05, (1 << 25) + 1)):
            g = H.random_font(n + glyph_count, glyph_count, font_size)
            text = H.random_text(n, n + 1)
            font = Font(ctx, font_size, g)
            font.setCharacterSpacing(spacing)
            want = _oracle_text(O, g, font_size, spacing, text, x0, 40, color, size)
            ptr, nq = font.drawText(text, x0, 40, color, size)
            assert nq == want.shape[0] and (n < 100 or 0 < nq < n)
            assert np.array_equal(ctx.download_quads(ptr, nq).view(np.uint32), want.view(np.uint32))
            # right to left (Font.cpp:283-327): every line from its last character backwards, the text up to its first NUL
            for t in (text, text.replace(bytes([0]), bytes([1]))):
                want = _oracle_text(O, g, font_size, spacing, t, x0, 40, color, size, rtl=True)
                ptr, nq = font.drawText(t, x0, 40, color, size, right_to_left=True)
                assert nq == want.shape[0]
                if nq:
                    assert np.array_equal(ctx.download_quads(ptr, nq).view(np.uint32), want.view(np.uint32))
            font.release()
    finally:
        ctx.close()


def test_text_edge_cases_and_device_resident_text():
    import torch
    O = H.oracle(fresh=True)
    ctx = Context(0)
    try:
        color = np.array([1.0, 0.5, 0.25, 0.75], dtype=np.float32)
        g = H.random_font(3, 95, 24)
        font = Font(ctx, 24, g)
        # nothing but line breaks and blanks; one long line without a break; only bytes the walk skips; a break as the last byte
        for text in (b"\n\r\n   \t\n", bytes(range(33, 127)) * 300, bytes([200, 0, 7, 255]) * 100, b"ab\ncd\n", b"\tx", b"  a b  \n\t\tc\td \r e"):
            for rtl in (False, True):
                want = _oracle_text(O, g, 24, 0.0, text, 3, -9, color, 0, rtl)
                ptr, nq = font.drawText(text, 3, -9, color, right_to_left=rtl)
                assert nq == want.shape[0], (text[:20], rtl)
                if nq:
                    assert np.array_equal(ctx.download_quads(ptr, nq).view(np.uint32), want.view(np.uint32)), (text[:20], rtl)
        # the text already on the device, the quads into a foreign allocation; the count comes back from the device
        text = H.random_text(50_000, 77)
        want = _oracle_text(O, g, 24, 0.0, text, 0, 0, color, 36)
        dev_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
        dst = torch.zeros(len(text) * 36, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        ptr, nq = font.drawText(None, 0, 0, color, 36, device_dst=dst.data_ptr(), dst_quads=len(text), text_on_device=(dev_text.data_ptr(), len(text)))
        ctx.sync()
        assert ptr == dst.data_ptr() and nq == want.shape[0]
        assert np.array_equal(dst.cpu().numpy()[:nq * 36].view(np.uint32).reshape(nq, 4, 9), want.view(np.uint32))
        assert float(dst[nq * 36:].abs().sum()) == 0.0          # nothing written beyond the text's quads
        # errors are codes: a destination that is too small
        with pytest.raises(Exception):
            font.drawText(b"abc", 0, 0, color, device_dst=dst.data_ptr(), dst_quads=2)
        font.release()
    finally:
        ctx.close()
