#!/usr/bin/env python
"""SURVEY 8(f3) measured: the 2-D SpriteBatch overloads as a batched expansion kernel and TileSet::draw from a
device-resident tile table, against the reference's own SpriteBatch.cpp on one host core (oracle/_ref).

    python scripts/bench_sprites.py [n_sprites] [tile rows] [tile cols]

Algorithmic bytes: 84 B read + 144 B written per sprite (228 B; + 84 B when the list may hold clipped sprites: the count pass); per tile 4 B cell index + 8 B source read + 144 B written."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import cpu_harness as H  # noqa: E402
from tractor3d_b200.emitter import Context, Font, TileSet  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cols = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
PEAK = 6541.1
stream = torch.cuda.Stream()
ctx = Context(0, stream.cuda_stream)


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) / reps * 1e3   # device ms, wall ms


print(f"{n} sprites; tile set {rows} x {cols}")
for label, kw in (("plain + centred (no clipping, no rotation)", dict(clipped=False, rotated=False)),
                  ("every overload (a third clipped, a third rotated)", dict())):
    sp = H.random_sprites(n, 11, **kw)
    dev_list = torch.from_numpy(sp.view(np.uint8).copy()).cuda()
    dst = torch.empty(n * 36, dtype=torch.float32, device="cuda")
    # a list that may hold clipped sprites takes three launches (kept sprites per tile -- a second read of the list --, scan,
    # expansion); a caller that knows its list has none says so (T3D_SPRITES_NO_CLIP): one launch
    no_clip = "clipped" in kw
    d_ms, w_ms = timed(lambda: ctx.draw_sprites(dev_list.data_ptr(), device_dst=dst.data_ptr(), dst_quads=n, on_device=True, n=n, want_count=False,
                                                no_clip=no_clip))
    alg = (228 if no_clip else 228 + 84) * n
    print(f"  {label}\n    list resident in HBM ({'one launch, T3D_SPRITES_NO_CLIP' if no_clip else 'three launches: culled sprites leave no gap; the list is read twice'}): "
          f"{d_ms:.3f} ms per list = {n / d_ms / 1e6:.2f} G sprites/s, {alg / d_ms / 1e6:.0f} GB/s algorithmic ({alg / d_ms / 1e6 / PEAK:.2f} of the measured HBM peak)")
    d_ms, w_ms = timed(lambda: ctx.draw_sprites(sp, device_dst=dst.data_ptr(), dst_quads=n, want_count=False), reps=10)
    print(f"    list in pageable host memory (copied into pinned staging, then uploaded: 84 B per sprite): {w_ms:.3f} ms wall per list = "
          f"{n / w_ms / 1e6:.3f} G sprites/s; host -> device {84 * n / 1e6:.1f} MB per list")

    def mapped_list():
        m = ctx.map_sprites(n)          # (a real producer writes its records here; the fill is not part of the hand-off)
        ctx.draw_sprites(m, device_dst=dst.data_ptr(), dst_quads=n, want_count=False, no_clip="clipped" in kw)
    d_ms, w_ms = timed(mapped_list, reps=10)
    print(f"    list written in place into the mapped pinned buffer (t3d_ctx_map_sprites): {w_ms:.3f} ms wall per list = "
          f"{n / w_ms / 1e6:.3f} G sprites/s = {84 * n / w_ms / 1e6:.1f} GB/s over PCIe")
    R = H.ref(fresh=True)
    m = min(n, 1 << 18)
    t = C.c_double()
    out = np.zeros((m, 36), dtype=np.float32)
    R.sprites_2d(sp.ctypes.data, m, 512, 256, out.ctypes.data_as(C.POINTER(C.c_float)), m, C.byref(t))
    R.sprites_2d(sp.ctypes.data, m, 512, 256, out.ctypes.data_as(C.POINTER(C.c_float)), m, C.byref(t))
    print(f"    reference SpriteBatch.cpp, one host core, {m} sprites: {t.value * 1e3:.2f} ms = {m / t.value / 1e6:.2f} M sprites/s")

rng = np.random.default_rng(3)
src = np.zeros((rows, cols, 2), dtype=np.float32)
src[..., 0] = rng.integers(0, 8, (rows, cols)) * 32.0
src[..., 1] = rng.integers(0, 4, (rows, cols)) * 32.0
src[rng.random((rows, cols)) < 0.1] = -1.0
ts = TileSet(ctx, 256, 128, 16.0, 16.0, rows, cols)
ts.set_tiles(src)
cells = int((src[..., 0] >= 0).sum())
dst = torch.empty(cells * 36, dtype=torch.float32, device="cuda")
color = (1.0, 1.0, 1.0, 0.8)
frame = [0]


def draw_tiles():
    frame[0] += 1
    ts.draw((0.25 * frame[0], -0.5 * frame[0], 0.0), color, device_dst=dst.data_ptr(), dst_quads=cells)


d_ms, w_ms = timed(draw_tiles)
print(f"  tile set, {cells} non-blank cells, table resident in HBM, ({cols} + {rows}) x 4 B uploaded per frame:\n"
      f"    {d_ms:.3f} ms device, {w_ms:.3f} ms wall per frame = {cells / d_ms / 1e6:.2f} G tiles/s, "
      f"{156 * cells / d_ms / 1e6:.0f} GB/s algorithmic ({156 * cells / d_ms / 1e6 / PEAK:.2f} of the measured HBM peak)")
R = H.ref(fresh=True)
r2, c2 = min(rows, 512), min(cols, 512)
h = R.tileset_create(256, 128, 16.0, 16.0, r2, c2)
for r in range(r2):
    for c in range(c2):
        if src[r, c, 0] >= 0:
            R.tileset_set_tile_source(h, c, r, float(src[r, c, 0]), float(src[r, c, 1]))
z3 = np.zeros(3, dtype=np.float32); col = np.ones(4, dtype=np.float32)
fpp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
R.tileset_draw(h, fpp(z3), fpp(z3), fpp(col), 0.8, None, 0)
t0 = time.perf_counter()
for _ in range(3):
    nq = R.tileset_draw(h, fpp(z3), fpp(z3), fpp(col), 0.8, None, 0)
t = (time.perf_counter() - t0) / 3
print(f"    reference TileSet.cpp, one host core, {r2} x {c2} grid ({nq} tiles): {t * 1e3:.2f} ms per frame = {nq / t / 1e6:.2f} M tiles/s")

# ---- Font::drawText: 1 byte read per character (+ the tile summaries), 144 B written per glyph ----
n_text = n
g = H.random_font(5, 95, 24)
text = H.random_text(n_text, 6)
font = Font(ctx, 24, g)
color4 = np.array([1.0, 1.0, 1.0, 1.0], dtype=np.float32)
glyphs_drawn = sum(1 for b in text if 32 < b < 127)
dev_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
dst = torch.empty(n_text * 36, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
_, nq = font.drawText(None, 0, 0, color4, 30, device_dst=dst.data_ptr(), dst_quads=n_text, text_on_device=(dev_text.data_ptr(), n_text))
assert nq == glyphs_drawn
# (the timed calls do not ask for the count back: no wait)
lib, fh = ctx.lib, font.h
cp = color4.ctypes.data_as(C.POINTER(C.c_float))
from tractor3d_b200 import abi  # noqa: E402


def text_on_device():
    lib.t3d_font_draw_text(fh, C.c_void_p(dev_text.data_ptr()), n_text, abi.TEXT_ON_DEVICE, 0, 0, cp, 30, 0, C.c_void_p(dst.data_ptr()), n_text, None, None)


def text_on_host():
    lib.t3d_font_draw_text(fh, text, n_text, 0, 0, 0, cp, 30, 0, C.c_void_p(dst.data_ptr()), n_text, None, None)


bytes_alg = n_text + 144 * glyphs_drawn
d_ms, w_ms = timed(text_on_device)
print(f"  text, {n_text} bytes -> {glyphs_drawn} glyph quads, text resident in HBM (3 launches: tile summaries, scan, expansion):\n"
      f"    {d_ms:.3f} ms device per text = {glyphs_drawn / d_ms / 1e6:.2f} G glyphs/s, {bytes_alg / d_ms / 1e6:.0f} GB/s algorithmic "
      f"({bytes_alg / d_ms / 1e6 / PEAK:.2f} of the measured HBM peak)")
d_ms, w_ms = timed(text_on_host, reps=10)
print(f"    text in host memory (1 B per character uploaded, glyphs counted on the host): {w_ms:.3f} ms wall per text = {glyphs_drawn / w_ms / 1e6:.3f} G glyphs/s")
# a longer text: the three launches weigh less
big = 16 * n_text
big_text = text * 16
big_glyphs = 16 * glyphs_drawn
dev_big = torch.frombuffer(bytearray(big_text), dtype=torch.uint8).cuda()
dst_big = torch.empty(big * 36, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()


def big_on_device():
    lib.t3d_font_draw_text(fh, C.c_void_p(dev_big.data_ptr()), big, abi.TEXT_ON_DEVICE, 0, 0, cp, 30, 0, C.c_void_p(dst_big.data_ptr()), big, None, None)


d_ms, w_ms = timed(big_on_device, reps=10)
print(f"    {big} bytes -> {big_glyphs} glyph quads, resident: {d_ms:.3f} ms = {big_glyphs / d_ms / 1e6:.2f} G glyphs/s, "
      f"{(big + 144 * big_glyphs) / d_ms / 1e6:.0f} GB/s algorithmic ({(big + 144 * big_glyphs) / d_ms / 1e6 / PEAK:.2f} of the measured HBM peak)")
del dev_big, dst_big
R = H.ref(fresh=True)
m = min(n_text, 1 << 18)
h = R.font_create(24, g.ctypes.data, 95, 512, 512)
t = C.c_double()
out = np.zeros((m, 36), dtype=np.float32)
for _ in range(2):
    nq = R.font_draw_text(h, text, m, 0, 0, cp, 30, 0, out.ctypes.data_as(C.POINTER(C.c_float)), m, C.byref(t))
print(f"    reference Font.cpp + SpriteBatch.cpp, one host core, {m} bytes ({nq} glyphs): {t.value * 1e3:.2f} ms = {nq / t.value / 1e6:.2f} M glyphs/s")
R.font_destroy(h)
font.release()
ctx.close()
