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
from tractor3d_b200.emitter import Context, TileSet

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
    # the committed vectors the reference's own SpriteBatch.cpp / TileSet.cpp produced (tests/golden/make_golden_sprites.py)
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
    finally:
        ctx.close()
