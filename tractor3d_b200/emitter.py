"""Python mirror of the reference's ParticleEmitter public interface over the C ABI.

Method names and argument meaning follow Tractor3Dlib/include/graphics/ParticleEmitter.h:179-733
(camelCase kept so parity tests read like calls on the reference class).  Every call goes through
libt3d_particles.so; nothing here computes particle state.
"""
import ctypes as C

import numpy as np

from . import abi

_P = C.POINTER
IDENTITY = np.eye(4, dtype=np.float32).reshape(16)


class T3DError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"t3d error {code}: {message}")
        self.code = code


def _check(lib, rc):
    if rc != abi.OK:
        raise T3DError(rc, lib.t3d_last_error().decode(errors="replace"))


def _fptr(a):
    return a.ctypes.data_as(_P(C.c_float))


class Context:
    """One per GPU (t3d_ctx).  `stream` may be a raw cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=0, stream=None):
        self.lib = abi.load()
        self.h = C.c_void_p()
        _check(self.lib, self.lib.t3d_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self.h)))
        self.device = device
        self.vertex_mode = abi.VERTEX_WORLD

    def close(self):
        if self.h:
            self.lib.t3d_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def flush(self):
        _check(self.lib, self.lib.t3d_ctx_flush(self.h))

    def sync(self):
        _check(self.lib, self.lib.t3d_ctx_sync(self.h))

    def set_rotation_mode(self, mode):
        _check(self.lib, self.lib.t3d_ctx_set_rotation_mode(self.h, mode))

    def set_vertex_mode(self, mode):
        """abi.VERTEX_WORLD (SpriteVertex, default) or abi.VERTEX_CLIP (position already projected, sprite.vert:19)."""
        _check(self.lib, self.lib.t3d_ctx_set_vertex_mode(self.h, mode))
        self.vertex_mode = mode

    def copy_to_host_async(self, host_ptr, device_ptr, nbytes):
        """Device -> pinned host copy on the context's download stream, beside the launches that follow.  Returns a ticket."""
        t = C.c_uint64()
        _check(self.lib, self.lib.t3d_ctx_copy_to_host_async(self.h, C.c_void_p(host_ptr), C.c_void_p(device_ptr), nbytes, C.byref(t)))
        return t.value

    def copy_wait(self, ticket):
        _check(self.lib, self.lib.t3d_ctx_copy_wait(self.h, ticket))

    def set_frozen_rotation(self, u, angle):
        a = np.ascontiguousarray(u, dtype=np.float32)
        _check(self.lib, self.lib.t3d_ctx_set_frozen_rotation(self.h, _fptr(a), angle))

    def frozen_rotation(self):
        m = np.zeros(16, dtype=np.float32)
        latched = C.c_int()
        _check(self.lib, self.lib.t3d_ctx_get_frozen_rotation(self.h, _fptr(m), C.byref(latched)))
        return m, bool(latched.value)

    def reset_statics(self):
        _check(self.lib, self.lib.t3d_ctx_reset_statics(self.h))

    def set_private_statics(self, on):
        _check(self.lib, self.lib.t3d_ctx_set_private_statics(self.h, int(bool(on))))

    def last_kernel_ms(self, which):
        ms = C.c_float()
        _check(self.lib, self.lib.t3d_ctx_last_kernel_ms(self.h, which, C.byref(ms)))
        return ms.value

    def kernel_time(self, which):
        """(total device ms, launches) of kind `which` (0 spawn, 1 update, 2 draw) since the context was created."""
        ms, n = C.c_double(), C.c_uint64()
        _check(self.lib, self.lib.t3d_ctx_kernel_time(self.h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def strip_indices(self, n_quads):
        """MeshBatch's triangle-strip index pattern for n_quads quads (uint32 numpy copy)."""
        n = C.c_uint64()
        out = np.zeros(max(6 * n_quads - 2, 1), dtype=np.uint32)
        _check(self.lib, self.lib.t3d_ctx_download_strip_indices(self.h, n_quads, out.ctypes.data_as(_P(C.c_uint32)), out.size, C.byref(n)))
        return out[: n.value]

    def triangle_indices(self, n_quads):
        """Triangle-list indices for n_quads quads (uint32 numpy copy)."""
        n = C.c_uint64()
        out = np.zeros(max(6 * n_quads, 1), dtype=np.uint32)
        _check(self.lib, self.lib.t3d_ctx_download_triangle_indices(self.h, n_quads, out.ctypes.data_as(_P(C.c_uint32)), out.size, C.byref(n)))
        return out[: n.value]

    def launch_count(self):
        n = C.c_uint64()
        _check(self.lib, self.lib.t3d_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def draw_sprites(self, sprites, device_dst=None, dst_quads=0, on_device=False, n=None, want_count=True, no_clip=False):
        """One SpriteBatch::draw 2-D call per record (numpy array of abi.SPRITE2D_DTYPE -- possibly the one map_sprites
        returned -- or a device pointer with on_device=True and n).  Returns (device pointer of the quads, quads written or None)."""
        if on_device:
            src, count = C.c_void_p(sprites), n
        else:
            a = np.ascontiguousarray(sprites)
            assert a.dtype.itemsize == C.sizeof(abi.Sprite2D)
            src, count = C.c_void_p(a.ctypes.data), a.size
        flags = (abi.SPRITES_ON_DEVICE if on_device else 0) | (abi.SPRITES_NO_CLIP if no_clip else 0)
        ptr, nq = C.c_void_p(), C.c_uint32()
        _check(self.lib, self.lib.t3d_ctx_draw_sprites(self.h, src, count, flags, C.c_void_p(device_dst) if device_dst else None,
                                                       dst_quads, C.byref(ptr), C.byref(nq) if want_count else None))
        return ptr.value, (nq.value if want_count else None)

    def map_sprites(self, n):
        """The pinned staging buffer of the next list as a numpy array of n abi.SPRITE2D_DTYPE records: fill it, then pass
        it to draw_sprites (no further host copy)."""
        p = C.c_void_p()
        _check(self.lib, self.lib.t3d_ctx_map_sprites(self.h, n, C.byref(p)))
        buf = (C.c_char * (n * C.sizeof(abi.Sprite2D))).from_address(p.value)
        return np.frombuffer(buf, dtype=abi.SPRITE2D_DTYPE, count=n)

    def download_quads(self, device_ptr, n_quads):
        out = np.zeros((n_quads, 4, 9), dtype=np.float32)
        _check(self.lib, self.lib.t3d_ctx_download_quads(self.h, C.c_void_p(device_ptr), n_quads, _fptr(out.reshape(-1)) if n_quads else None))
        return out

    def host_profile(self):
        """Seconds the library has spent on the host: {enqueue, build, submit, wait} (t3d_ctx_host_profile)."""
        out = (C.c_double * 4)()
        _check(self.lib, self.lib.t3d_ctx_host_profile(self.h, out))
        return dict(zip(("enqueue", "build", "submit", "wait"), out))

    def transfer_bytes(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(self.lib, self.lib.t3d_ctx_transfer_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- batches --
    @staticmethod
    def _handles(emitters):
        arr = (C.c_void_p * len(emitters))(*[e.h for e in emitters])
        return arr

    def batch_update(self, emitters, elapsed_ms, handles=None):
        h = handles if handles is not None else self._handles(emitters)
        _check(self.lib, self.lib.t3d_batch_update(h, len(h), elapsed_ms))

    def batch_draw(self, emitters, handles=None):
        h = handles if handles is not None else self._handles(emitters)
        _check(self.lib, self.lib.t3d_batch_draw(h, len(h)))

    def batch_emit_once(self, emitters, counts, handles=None):
        h = handles if handles is not None else self._handles(emitters)
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        _check(self.lib, self.lib.t3d_batch_emit_once(h, len(h), c.ctypes.data_as(_P(C.c_uint32))))

    def batch_set_transforms(self, emitters, node_worlds=None, camera_world=None, handles=None):
        """node_worlds: (n, 16) float32 or None; camera_world: (16,) float32 shared by all, or None."""
        h = handles if handles is not None else self._handles(emitters)
        nw = np.ascontiguousarray(node_worlds, dtype=np.float32).reshape(len(h), 16) if node_worlds is not None else None
        cw = np.ascontiguousarray(camera_world, dtype=np.float32).reshape(16) if camera_world is not None else None
        _check(self.lib, self.lib.t3d_batch_set_transforms(h, len(h), _fptr(nw) if nw is not None else None,
                                                           _fptr(cw) if cw is not None else None))

    def batch_frame(self, emitters, elapsed_ms, node_worlds=None, camera_world=None, handles=None, counts_out=None):
        """One frame for many emitters in one call (t3d_batch_frame): transforms in, update + draw, and -- when counts_out
        (a uint32 array of len(emitters)) is given -- the live counts back.  node_worlds / camera_world must already be
        contiguous float32 arrays (no conversion here: this is the per-frame call)."""
        h = handles if handles is not None else self._handles(emitters)
        _check(self.lib, self.lib.t3d_batch_frame(h, len(h), _fptr(node_worlds) if node_worlds is not None else None,
                                                  _fptr(camera_world) if camera_world is not None else None, elapsed_ms,
                                                  counts_out.ctypes.data_as(_P(C.c_uint32)) if counts_out is not None else None))
        return counts_out

    def batch_frame_pipelined(self, emitters, elapsed_ms, node_worlds=None, camera_world=None, handles=None, previous_counts=None):
        """t3d_batch_frame_pipelined: the frame is enqueued without waiting for it; previous_counts (uint32 array) receives the
        live counts the previous such call left.  Returns True when previous_counts was filled."""
        h = handles if handles is not None else self._handles(emitters)
        have = C.c_int(0)
        _check(self.lib, self.lib.t3d_batch_frame_pipelined(h, len(h), _fptr(node_worlds) if node_worlds is not None else None,
                                                            _fptr(camera_world) if camera_world is not None else None, elapsed_ms,
                                                            previous_counts.ctypes.data_as(_P(C.c_uint32)) if previous_counts is not None else None,
                                                            C.byref(have)))
        return bool(have.value)

    def batch_counts(self, emitters, handles=None):
        h = handles if handles is not None else self._handles(emitters)
        out = np.zeros(len(h), dtype=np.uint32)
        _check(self.lib, self.lib.t3d_batch_get_counts(h, len(h), out.ctypes.data_as(_P(C.c_uint32))))
        return out


class Font:
    """One size of tractor::Font (include/renderer/Font.h): drawText(text, x, y, color, size) laid out and expanded on the device."""

    def __init__(self, ctx, size, glyphs):
        """glyphs: numpy array of abi.GLYPH_DTYPE (Font::Glyph: code, width, bearingX, advance, uvs[4]); glyph k is character 32 + k."""
        self.ctx, self.lib = ctx, ctx.lib
        g = np.ascontiguousarray(glyphs, dtype=abi.GLYPH_DTYPE)
        self._glyphs, self._size = g, size
        self.h = C.c_void_p()
        _check(self.lib, self.lib.t3d_font_create(ctx.h, size, g.ctypes.data, len(g), C.byref(self.h)))

    def setCharacterSpacing(self, spacing):
        self._spacing = spacing
        _check(self.lib, self.lib.t3d_font_set_character_spacing(self.h, spacing))

    def measureText(self, text, size=0):
        """Font::measureText(text, size, &width, &height): (width, height) in pixels -- host logic, no GPU work."""
        w, h = C.c_uint32(), C.c_uint32()
        self.lib.t3d_host_font_measure_text(self._glyphs.ctypes.data, self._glyphs.size, self._size, getattr(self, "_spacing", 0.0), bytes(text), len(text),
                                            size, C.byref(w), C.byref(h))
        return w.value, h.value

    def drawText(self, text, x, y, color, size=0, right_to_left=False, device_dst=None, dst_quads=0, text_on_device=None):
        """text: bytes (uploaded) -- or, with text_on_device=(pointer, length), text already on the device.  Returns
        (device pointer of the quads, number of quads)."""
        col = np.ascontiguousarray(color, dtype=np.float32)
        ptr, nq = C.c_void_p(), C.c_uint32()
        if text_on_device is not None:
            tp, length, flags = C.c_void_p(text_on_device[0]), text_on_device[1], abi.TEXT_ON_DEVICE
        else:
            buf = C.create_string_buffer(bytes(text), len(text))
            tp, length, flags = C.cast(buf, C.c_void_p), len(text), 0
        _check(self.lib, self.lib.t3d_font_draw_text(self.h, tp, length, flags, x, y, _fptr(col), size, 1 if right_to_left else 0,
                                                     C.c_void_p(device_dst) if device_dst else None, dst_quads, C.byref(ptr), C.byref(nq)))
        return ptr.value, nq.value

    def release(self):
        if self.h:
            self.lib.t3d_font_destroy(self.h)
            self.h = None


class TileSet:
    """tractor::TileSet's draw path (include/graphics/TileSet.h) with the tile table resident on the device."""

    def __init__(self, ctx, texture_width, texture_height, tile_width, tile_height, row_count, column_count):
        self.ctx, self.lib = ctx, ctx.lib
        self.rows, self.cols = row_count, column_count
        self.h = C.c_void_p()
        _check(self.lib, self.lib.t3d_tileset_create(ctx.h, texture_width, texture_height, tile_width, tile_height, row_count, column_count,
                                                     C.byref(self.h)))

    def setTileSource(self, column, row, source):
        _check(self.lib, self.lib.t3d_tileset_set_tile_source(self.h, column, row, source[0], source[1]))

    def set_tiles(self, sources):
        a = np.ascontiguousarray(sources, dtype=np.float32).reshape(-1)
        assert a.size == self.rows * self.cols * 2
        _check(self.lib, self.lib.t3d_tileset_set_tiles(self.h, _fptr(a)))

    def draw(self, position, color, device_dst=None, dst_quads=0):
        """Returns (device pointer, quads).  position / color as TileSet::draw has them by TileSet.cpp:205 / :226."""
        p = np.ascontiguousarray(position, dtype=np.float32)
        c = np.ascontiguousarray(color, dtype=np.float32)
        ptr, n = C.c_void_p(), C.c_uint32()
        _check(self.lib, self.lib.t3d_tileset_draw(self.h, _fptr(p), _fptr(c), C.c_void_p(device_dst) if device_dst else None, dst_quads,
                                                   C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def release(self):
        if self.h:
            self.lib.t3d_tileset_destroy(self.h)
            self.h = C.c_void_p()


class ParticleEmitter:
    """Drop-in for tractor::ParticleEmitter.  A fresh emitter is attached to an identity node and looks
    through an identity camera node (what the reference takes from Node / Scene, PE.cpp:299, 860-861)."""

    BLEND_NONE, BLEND_ALPHA, BLEND_ADDITIVE, BLEND_MULTIPLIED = range(4)

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.lib = ctx.lib
        self.h = handle
        self.set_node_world(IDENTITY)
        self.set_camera_world(IDENTITY)

    # -- creation (ParticleEmitter.h:179-205) --
    @classmethod
    def create(cls, ctx, params, sprite=None, rng=(abi.RNG_LIBC, 0)):
        h = C.c_void_p()
        _check(ctx.lib, ctx.lib.t3d_emitter_create(ctx.h, C.byref(params), C.byref(h)))
        e = cls(ctx, h)
        if sprite:
            e.setSpriteFrameCoords(*sprite)
        e.set_rng(*rng)
        return e

    @classmethod
    def create_from_file(cls, ctx, url, rng=(abi.RNG_LIBC, 0)):
        h = C.c_void_p()
        _check(ctx.lib, ctx.lib.t3d_emitter_create_from_file(ctx.h, url.encode(), C.byref(h)))
        e = cls(ctx, h)
        e.set_rng(*rng)
        return e

    @classmethod
    def create_from_properties(cls, ctx, particle_namespace, rng=(abi.RNG_LIBC, 0)):
        """ParticleEmitter::create(Properties*) (PE.cpp:92-211): `particle_namespace` is a t3d_properties handle."""
        h = C.c_void_p()
        _check(ctx.lib, ctx.lib.t3d_emitter_create_from_properties(ctx.h, particle_namespace, C.byref(h)))
        e = cls(ctx, h)
        e.set_rng(*rng)
        return e

    def clone(self):
        h = C.c_void_p()
        _check(self.lib, self.lib.t3d_emitter_clone(self.h, C.byref(h)))
        return ParticleEmitter(self.ctx, h)

    def release(self):
        if self.h:
            self.lib.t3d_emitter_destroy(self.h)
            self.h = C.c_void_p()

    # -- configuration --
    def params(self):
        p = abi.EmitterParams()
        _check(self.lib, self.lib.t3d_emitter_get_params(self.h, C.byref(p)))
        return p

    def set_params(self, p):
        _check(self.lib, self.lib.t3d_emitter_set_params(self.h, C.byref(p)))

    def _edit(self, **kw):
        p = self.params()
        for k, v in kw.items():
            if isinstance(v, (list, tuple, np.ndarray)):
                typ = dict(abi.EmitterParams._fields_)[k]
                setattr(p, k, typ(*[float(x) for x in v]))
            else:
                setattr(p, k, v)
        self.set_params(p)

    def getParticleCountMax(self):
        return self.params().particle_count_max

    def setParticleCountMax(self, n):
        # PE.h:237 rewrites the field only: lowering it (and raising it back up to the size at create) is all that works
        self._edit(particle_count_max=int(n))

    def setTextureSize(self, width, height, blend_mode):
        """setTexture(Texture*, BlendMode) (PE.cpp:229-252): new texel ratios, sprite table reset to one full frame."""
        _check(self.lib, self.lib.t3d_emitter_set_texture_size(self.h, width, height, blend_mode))

    def setEmissionRate(self, rate):
        _check(self.lib, self.lib.t3d_emitter_set_emission_rate(self.h, rate))

    def getEmissionRate(self):
        return self.params().emission_rate

    def setEllipsoid(self, v):
        self._edit(ellipsoid=int(bool(v)))

    def setSize(self, start_min, start_max, end_min, end_max):
        self._edit(size_start_min=start_min, size_start_max=start_max, size_end_min=end_min, size_end_max=end_max)

    def setEnergy(self, energy_min, energy_max):
        # ParticleEmitter::setEnergy takes long and stores float (PE.cpp:381-385).
        self._edit(energy_min=float(int(energy_min)), energy_max=float(int(energy_max)))

    def setColor(self, start, start_var, end, end_var):
        self._edit(color_start=start, color_start_var=start_var, color_end=end, color_end_var=end_var)

    def setPosition(self, position, variance):
        self._edit(position=position, position_var=variance)

    def setVelocity(self, velocity, variance):
        self._edit(velocity=velocity, velocity_var=variance)

    def setAcceleration(self, acceleration, variance):
        self._edit(acceleration=acceleration, acceleration_var=variance)

    def setRotationPerParticle(self, speed_min, speed_max):
        self._edit(rotation_per_particle_speed_min=speed_min, rotation_per_particle_speed_max=speed_max)

    def setRotation(self, speed_min, speed_max, axis, axis_var):
        self._edit(rotation_speed_min=speed_min, rotation_speed_max=speed_max, rotation_axis=axis, rotation_axis_var=axis_var)

    def setSpriteAnimated(self, v):
        self._edit(sprite_animated=int(bool(v)))

    def setSpriteLooped(self, v):
        self._edit(sprite_looped=int(bool(v)))

    def setSpriteFrameRandomOffset(self, v):
        self._edit(sprite_frame_random_offset=int(v))

    def setSpriteFrameDuration(self, ms):
        self._edit(sprite_frame_duration=int(ms))

    def setOrbit(self, position, velocity, acceleration):
        self._edit(orbit_position=int(bool(position)), orbit_velocity=int(bool(velocity)), orbit_acceleration=int(bool(acceleration)))

    def setBlendMode(self, mode):
        self._edit(blend_mode=int(mode))

    def setSpriteTexCoords(self, tex_coords):
        a = np.ascontiguousarray(tex_coords, dtype=np.float32).reshape(-1)
        _check(self.lib, self.lib.t3d_emitter_set_sprite_tex_coords(self.h, a.size // 4, _fptr(a)))

    def setSpriteFrameRects(self, rects):
        a = np.ascontiguousarray(rects, dtype=np.float32).reshape(-1)
        _check(self.lib, self.lib.t3d_emitter_set_sprite_frame_rects(self.h, a.size // 4, _fptr(a)))

    def setSpriteFrameCoords(self, frame_count, width, height):
        _check(self.lib, self.lib.t3d_emitter_set_sprite_frame_coords(self.h, frame_count, width, height))

    def getSpriteFrameCount(self):
        n = C.c_uint32()
        _check(self.lib, self.lib.t3d_emitter_get_sprite_tex_coords(self.h, C.byref(n), None, 0))
        return n.value

    def sprite_tex_coords(self):
        n = self.getSpriteFrameCount()
        out = np.zeros(n * 4, dtype=np.float32)
        _check(self.lib, self.lib.t3d_emitter_get_sprite_tex_coords(self.h, None, _fptr(out), n))
        return out.reshape(n, 4)

    def depth_sorted_indices(self, view_dir):
        """Triangle-list indices (uint32 numpy copy) that walk the current particles' quads farthest first along view_dir."""
        v = np.ascontiguousarray(view_dir, dtype=np.float32)
        n = C.c_uint64()
        _check(self.lib, self.lib.t3d_emitter_download_depth_sorted_indices(self.h, _fptr(v), None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.uint32)
        _check(self.lib, self.lib.t3d_emitter_download_depth_sorted_indices(self.h, _fptr(v), out.ctypes.data_as(_P(C.c_uint32)), out.size, C.byref(n)))
        return out[: n.value]

    def set_rng(self, mode, seed=0):
        _check(self.lib, self.lib.t3d_emitter_set_rng(self.h, mode, seed))

    def feed_draws(self, draws):
        a = np.ascontiguousarray(draws, dtype=np.int32)
        _check(self.lib, self.lib.t3d_emitter_feed_draws(self.h, a.ctypes.data_as(_P(C.c_int32)), a.size))

    def pending_draws(self):
        n = C.c_size_t()
        _check(self.lib, self.lib.t3d_emitter_pending_draws(self.h, C.byref(n)))
        return n.value

    def set_node_world(self, m):
        a = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
        _check(self.lib, self.lib.t3d_emitter_set_node_world(self.h, _fptr(a)))

    def set_camera_world(self, m):
        a = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
        _check(self.lib, self.lib.t3d_emitter_set_camera_world(self.h, _fptr(a)))

    def set_view_projection(self, m):
        """Node::getViewProjectionMatrix() of the emitter's node (PE.cpp:846-849)."""
        a = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
        _check(self.lib, self.lib.t3d_emitter_set_view_projection(self.h, _fptr(a)))

    def enable_bounds(self, on=True):
        _check(self.lib, self.lib.t3d_emitter_enable_bounds(self.h, int(bool(on))))

    def bounds(self):
        """(lo, hi) of pos -+ size over the live particles as the last update() left them, or None."""
        lo, hi, ok = np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_int()
        _check(self.lib, self.lib.t3d_emitter_bounds(self.h, _fptr(lo), _fptr(hi), C.byref(ok)))
        return (lo, hi) if ok.value else None

    def set_culling(self, on=True):
        _check(self.lib, self.lib.t3d_emitter_set_culling(self.h, int(bool(on))))

    # -- the path --
    def start(self):
        _check(self.lib, self.lib.t3d_emitter_start(self.h))

    def stop(self):
        _check(self.lib, self.lib.t3d_emitter_stop(self.h))

    def isStarted(self):
        v = C.c_int()
        _check(self.lib, self.lib.t3d_emitter_is_started(self.h, C.byref(v)))
        return bool(v.value)

    def isActive(self):
        v = C.c_int()
        _check(self.lib, self.lib.t3d_emitter_is_active(self.h, C.byref(v)))
        return bool(v.value)

    def emitOnce(self, n):
        _check(self.lib, self.lib.t3d_emitter_emit_once(self.h, n))

    def update(self, elapsed_ms):
        _check(self.lib, self.lib.t3d_emitter_update(self.h, elapsed_ms))

    def draw(self, wireframe=False):
        v = C.c_uint32()
        _check(self.lib, self.lib.t3d_emitter_draw(self.h, C.byref(v)))
        return v.value

    def draw_to(self, device_ptr, capacity_quads):
        """draw() with the quads written into the caller's device buffer (e.g. a mapped interop vertex buffer)."""
        v = C.c_uint32()
        _check(self.lib, self.lib.t3d_emitter_draw_to(self.h, C.c_void_p(device_ptr), capacity_quads, C.byref(v)))
        return v.value

    def getParticlesCount(self):
        v = C.c_uint32()
        _check(self.lib, self.lib.t3d_emitter_get_count(self.h, C.byref(v)))
        return v.value

    # -- results --
    def state(self):
        """(NUM_COLUMNS, n) uint32 copy of the live particles (column order of include/t3d_particles.h)."""
        n = self.getParticlesCount()
        out = np.zeros((abi.NUM_COLUMNS, max(n, 1)), dtype=np.uint32)
        cnt = C.c_uint32()
        _check(self.lib, self.lib.t3d_emitter_download_state(self.h, out.ctypes.data_as(_P(C.c_uint32)), out.shape[1], C.byref(cnt)))
        return out[:, : cnt.value]

    def upload_state(self, cols):
        a = np.ascontiguousarray(cols, dtype=np.uint32)
        assert a.ndim == 2 and a.shape[0] == abi.NUM_COLUMNS
        _check(self.lib, self.lib.t3d_emitter_upload_state(self.h, a.ctypes.data_as(_P(C.c_uint32)), a.shape[1], a.shape[1]))

    def vertices_device(self):
        ptr, n = C.c_void_p(), C.c_uint32()
        _check(self.lib, self.lib.t3d_emitter_vertices(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def vertices(self, out=None):
        """(n_quads, 4, 9) float32 copy of the vertex stream of the last draw() -- (n_quads, 4, 10) in clip-space mode."""
        _, n = self.vertices_device()
        per = 10 if self.ctx.vertex_mode == abi.VERTEX_CLIP else 9
        buf = out if out is not None else np.zeros((max(n, 1), 4, per), dtype=np.float32)
        got = C.c_uint32()
        _check(self.lib, self.lib.t3d_emitter_download_vertices(self.h, _fptr(buf), buf.shape[0], C.byref(got)))
        return buf[: min(got.value, buf.shape[0])]
