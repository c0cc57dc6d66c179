"""ctypes drivers for the two CPU checkers (TEST INFRASTRUCTURE):

  * `oracle()`  -> oracle/_build/libt3doracle.so, the plain-C restatement (prefix orc_)
  * `ref()`     -> oracle/_ref/libt3dref.so, the UNMODIFIED reference compiled headless (prefix t3dref_);
                   None when it has not been built (it needs /root/reference at build time).

Both expose the same calls through `CpuEmitter`, so a scenario written once can be run on either.
The function-local statics of the reference (update rate-cap accumulator, frozen billboard rotation)
are per loaded library image; `fresh=True` loads a private copy of the .so so a test starts from a
"new process".
"""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

from tractor3d_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libt3doracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libt3dref.so")

_P = C.POINTER
_vp = C.c_void_p
_tmpdir = None


def _private_copy(path):
    global _tmpdir
    if _tmpdir is None:
        _tmpdir = tempfile.mkdtemp(prefix="t3d_cpu_")
    fd, dst = tempfile.mkstemp(suffix=".so", dir=_tmpdir)
    os.close(fd)
    shutil.copyfile(path, dst)
    return dst


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


class CpuLib:
    """One loaded image of a CPU checker."""

    def __init__(self, path, prefix, fresh=False):
        self.prefix = prefix
        self.kind = "reference" if prefix == "t3dref_" else "port"
        self.lib = C.CDLL(_private_copy(path) if fresh else path, mode=C.RTLD_LOCAL)
        L, p = self.lib, prefix

        def proto(name, res, args):
            fn = getattr(L, p + name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

        proto("rand_seed", None, [C.c_uint64])
        proto("rand_replay", None, [_P(C.c_int32), C.c_size_t])
        proto("rand_replay_remaining", C.c_size_t, [])
        proto("rand_record", None, [C.c_int])
        proto("rand_recorded", C.c_size_t, [_P(C.c_int32), C.c_size_t])
        proto("rand_calls", C.c_uint64, [])
        proto("emitter_create", _vp, [_P(abi.EmitterParams)])
        proto("emitter_destroy", None, [_vp])
        proto("emitter_set_params", None, [_vp, _P(abi.EmitterParams)])
        proto("emitter_get_params", None, [_vp, _P(abi.EmitterParams)])
        proto("emitter_set_emission_rate", None, [_vp, C.c_uint32])
        proto("emitter_set_sprite_tex_coords", None, [_vp, C.c_uint32, _P(C.c_float)])
        proto("emitter_set_sprite_frame_rects", None, [_vp, C.c_uint32, _P(C.c_float)])
        proto("emitter_set_sprite_frame_coords", None, [_vp, C.c_uint32, C.c_int, C.c_int])
        proto("emitter_get_sprite_tex_coords", C.c_uint32, [_vp, _P(C.c_float), C.c_size_t])
        proto("emitter_set_node_world", None, [_vp, _P(C.c_float)])
        proto("emitter_set_camera_world", None, [_vp, _P(C.c_float)])
        proto("emitter_get_emit_time", C.c_float, [_vp])
        proto("emitter_start", None, [_vp])
        proto("emitter_stop", None, [_vp])
        proto("emitter_is_started", C.c_int, [_vp])
        proto("emitter_is_active", C.c_int, [_vp])
        proto("emitter_emit_once", None, [_vp, C.c_uint32])
        proto("emitter_update", None, [_vp, C.c_float])
        proto("emitter_draw", C.c_uint32, [_vp])
        proto("emitter_get_count", C.c_uint32, [_vp])
        proto("emitter_get_state", C.c_uint32, [_vp, _P(C.c_uint32), C.c_size_t])
        proto("emitter_set_state", C.c_int, [_vp, _P(C.c_uint32), C.c_size_t, C.c_uint32])
        proto("emitter_get_vertices", C.c_uint32, [_vp, _P(C.c_float), C.c_size_t])
        proto("time_update", C.c_double, [_vp, C.c_float, C.c_int])
        proto("time_draw", C.c_double, [_vp, C.c_int])
        proto("time_frames", C.c_double, [_vp, C.c_float, C.c_int])
        if self.kind == "port":
            proto("emitter_depth_sorted_indices", C.c_uint64, [_vp, _P(C.c_float), _P(C.c_uint32), C.c_uint64])
            proto("sprites_2d", C.c_uint32, [_vp, C.c_size_t, _P(C.c_float), C.c_size_t])
            proto("tileset_draw", C.c_uint32, [_P(C.c_float), C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float,
                                               _P(C.c_float), _P(C.c_float), _P(C.c_float), C.c_size_t])
            proto("font_draw_text", C.c_uint32, [_vp, C.c_uint32, C.c_uint32, C.c_float, C.c_char_p, C.c_size_t, C.c_int, C.c_int, _P(C.c_float),
                                                 C.c_uint32, C.c_int, _P(C.c_float), C.c_size_t])
            proto("device_rng_draw", C.c_int32, [C.c_uint64, C.c_uint64, C.c_uint32])
            proto("reset_statics", None, [])
            proto("set_rotation_mode", None, [C.c_int])
            proto("set_frozen_rotation", None, [_P(C.c_float), C.c_float])
            proto("get_frozen_rotation", C.c_int, [_P(C.c_float)])
            proto("emitter_set_device_rng", None, [_vp, C.c_int, C.c_uint64])
            proto("strip_indices", C.c_uint64, [C.c_uint32, _P(C.c_uint32)])
            proto("emitter_get_clip_vertices", C.c_uint32, [_vp, _P(C.c_float), _P(C.c_float), C.c_size_t])
            proto("emitter_bounds", C.c_int, [_vp, _P(C.c_float), _P(C.c_float)])
        else:
            proto("sprites_2d", C.c_uint32, [_vp, C.c_size_t, C.c_uint32, C.c_uint32, _P(C.c_float), C.c_size_t, _P(C.c_double)])
            proto("tileset_create", _vp, [C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_uint32, C.c_uint32])
            proto("tileset_destroy", None, [_vp])
            proto("tileset_set_tile_source", None, [_vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float])
            proto("tileset_draw", C.c_uint32, [_vp, _P(C.c_float), _P(C.c_float), _P(C.c_float), C.c_float, _P(C.c_float), C.c_size_t])
            proto("ui_sprite_draw", C.c_uint32, [_vp, C.c_uint32, C.c_uint32, _P(C.c_float), C.c_size_t])
            proto("font_create", _vp, [C.c_uint32, _vp, C.c_uint32, C.c_uint32, C.c_uint32])
            proto("font_destroy", None, [_vp])
            proto("font_set_character_spacing", None, [_vp, C.c_float])
            proto("font_measure_text", None, [_vp, C.c_char_p, C.c_size_t, C.c_uint32, _P(C.c_uint32), _P(C.c_uint32)])
            proto("font_draw_text", C.c_uint32, [_vp, C.c_char_p, C.c_size_t, C.c_int, C.c_int, _P(C.c_float), C.c_uint32, C.c_int, _P(C.c_float),
                                                 C.c_size_t, _P(C.c_double)])
            proto("emitter_create_from_file", _vp, [C.c_char_p])
            proto("emitter_clone", _vp, [_vp])
            proto("set_quiet", None, [C.c_int])
            proto("set_resource_path", None, [C.c_char_p])
            proto("sizeof_particle", C.c_size_t, [])
            proto("emitter_set_discard_vertices", None, [_vp, C.c_int])
            proto("emitter_get_sprite_width", C.c_uint32, [_vp])
            proto("emitter_get_sprite_height", C.c_uint32, [_vp])
            proto("emitter_get_time_per_emission", C.c_float, [_vp])

    # -- draw-source helpers --
    def replay(self, draws):
        if draws is None:
            self.rand_replay(None, 0)
            return
        a = np.ascontiguousarray(draws, dtype=np.int32)
        self.rand_replay(a.ctypes.data_as(_P(C.c_int32)), a.size)

    def recorded(self):
        n = self.rand_recorded(None, 0)
        out = np.empty(n, dtype=np.int32)
        if n:
            self.rand_recorded(out.ctypes.data_as(_P(C.c_int32)), n)
        return out

    def emitter(self, params):
        return CpuEmitter(self, params)


def _m16(m):
    a = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    return a, a.ctypes.data_as(_P(C.c_float))


class CpuEmitter:
    """Thin object view over one emitter of a CpuLib (oracle port or the real reference)."""

    def __init__(self, lib, params=None, handle=None):
        self.L = lib
        self.h = handle if handle is not None else lib.emitter_create(C.byref(params))
        if not self.h:
            raise RuntimeError("emitter_create failed")
        self.capacity = self.params().particle_count_max

    def close(self):
        if self.h:
            self.L.emitter_destroy(self.h)
            self.h = None

    def params(self):
        p = abi.EmitterParams()
        self.L.emitter_get_params(self.h, C.byref(p))
        return p

    def set_params(self, p):
        self.L.emitter_set_params(self.h, C.byref(p))

    def set_emission_rate(self, r):
        self.L.emitter_set_emission_rate(self.h, r)

    def set_sprite_tex_coords(self, coords):
        a = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1)
        self.L.emitter_set_sprite_tex_coords(self.h, a.size // 4, a.ctypes.data_as(_P(C.c_float)))

    def set_sprite_frame_rects(self, rects):
        a = np.ascontiguousarray(rects, dtype=np.float32).reshape(-1)
        self.L.emitter_set_sprite_frame_rects(self.h, a.size // 4, a.ctypes.data_as(_P(C.c_float)))

    def set_sprite_frame_coords(self, n, w, h):
        self.L.emitter_set_sprite_frame_coords(self.h, n, w, h)

    def sprite_tex_coords(self):
        n = self.L.emitter_get_sprite_tex_coords(self.h, None, 0)
        out = np.empty(n * 4, dtype=np.float32)
        self.L.emitter_get_sprite_tex_coords(self.h, out.ctypes.data_as(_P(C.c_float)), n)
        return out.reshape(n, 4)

    def set_node_world(self, m):
        a, p = _m16(m)
        self.L.emitter_set_node_world(self.h, p)

    def set_camera_world(self, m):
        a, p = _m16(m)
        self.L.emitter_set_camera_world(self.h, p)

    def start(self):
        self.L.emitter_start(self.h)

    def stop(self):
        self.L.emitter_stop(self.h)

    def is_started(self):
        return bool(self.L.emitter_is_started(self.h))

    def is_active(self):
        return bool(self.L.emitter_is_active(self.h))

    def emit_once(self, n):
        self.L.emitter_emit_once(self.h, n)

    def update(self, ms):
        self.L.emitter_update(self.h, ms)

    def draw(self):
        return self.L.emitter_draw(self.h)

    def count(self):
        return self.L.emitter_get_count(self.h)

    def emit_time(self):
        return self.L.emitter_get_emit_time(self.h)

    def state(self):
        """(NUM_COLUMNS, n) uint32 view of the live particles."""
        n = self.count()
        out = np.zeros((abi.NUM_COLUMNS, max(n, 1)), dtype=np.uint32)
        self.L.emitter_get_state(self.h, out.ctypes.data_as(_P(C.c_uint32)), out.shape[1])
        return out[:, :n]

    def set_state(self, cols):
        a = np.ascontiguousarray(cols, dtype=np.uint32)
        assert a.shape[0] == abi.NUM_COLUMNS
        rc = self.L.emitter_set_state(self.h, a.ctypes.data_as(_P(C.c_uint32)), a.shape[1], a.shape[1])
        assert rc == 0

    def vertices(self):
        """(n_quads, 4, 9) float32 vertex stream captured from the last draw()."""
        n = self.L.emitter_get_vertices(self.h, None, 0)
        out = np.zeros((max(n, 1), 4, 9), dtype=np.float32)
        if n:
            self.L.emitter_get_vertices(self.h, out.ctypes.data_as(_P(C.c_float)), n)
        return out[:n]


    def clip_vertices(self, vp):
        """(n_quads, 4, 10) float32: the last draw()'s vertices with sprite.vert:19 applied (oracle port only)."""
        m = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
        n = self.L.emitter_get_clip_vertices(self.h, m.ctypes.data_as(_P(C.c_float)), None, 0)
        out = np.zeros((max(n, 1), 4, 10), dtype=np.float32)
        if n:
            self.L.emitter_get_clip_vertices(self.h, m.ctypes.data_as(_P(C.c_float)), out.ctypes.data_as(_P(C.c_float)), n)
        return out[:n]

    def depth_sorted_indices(self, view_dir):
        v = np.ascontiguousarray(view_dir, dtype=np.float32)
        n = self.L.emitter_depth_sorted_indices(self.h, v.ctypes.data_as(_P(C.c_float)), None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self.L.emitter_depth_sorted_indices(self.h, v.ctypes.data_as(_P(C.c_float)), out.ctypes.data_as(_P(C.c_uint32)), n)
        return out[:n]

    def bounds(self):
        lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
        ok = self.L.emitter_bounds(self.h, lo.ctypes.data_as(_P(C.c_float)), hi.ctypes.data_as(_P(C.c_float)))
        return (lo, hi) if ok else None


_oracle = None
_ref = None


def oracle(fresh=False):
    global _oracle
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    if fresh:
        return CpuLib(ORACLE_SO, "orc_", fresh=True)
    if _oracle is None:
        _oracle = CpuLib(ORACLE_SO, "orc_")
    return _oracle


def ref_available():
    return os.path.exists(REF_SO)


def ref(fresh=False):
    global _ref
    if not ref_available():
        return None
    if fresh:
        lib = CpuLib(REF_SO, "t3dref_", fresh=True)
        lib.set_quiet(1)
        return lib
    if _ref is None:
        _ref = CpuLib(REF_SO, "t3dref_")
        _ref.set_quiet(1)
    return _ref


# -- 2-D sprites and tile sets --
def random_sprites(n, seed, clipped=True, rotated=True):
    """n t3d_sprite2d records (numpy structured array) covering every overload: plain, centred, rotated (angle 0 and not,
    zero and non-zero pivots), clipped (inside, straddling every edge, outside)."""
    rng = np.random.default_rng(seed)
    s = np.zeros(n, dtype=abi.SPRITE2D_DTYPE)
    s["x"] = rng.uniform(-50, 800, n); s["y"] = rng.uniform(-50, 600, n); s["z"] = rng.uniform(0, 1, n)
    s["width"] = rng.uniform(1, 120, n); s["height"] = rng.uniform(1, 90, n)
    s["u1"] = rng.uniform(0, 0.5, n); s["u2"] = s["u1"] + rng.uniform(0.01, 0.5, n)
    s["v1"] = rng.uniform(0.5, 1, n); s["v2"] = s["v1"] - rng.uniform(0.01, 0.5, n)
    s["color"] = rng.uniform(0, 1, (n, 4))
    kind = rng.integers(0, 6, n)
    flags = np.zeros(n, dtype=np.uint32)
    flags[kind == 1] = abi.SPRITE_CENTERED
    if rotated:
        flags[kind == 2] = abi.SPRITE_ROTATED
        flags[kind == 3] = abi.SPRITE_ROTATED | abi.SPRITE_CENTERED
        s["rotation_point"] = rng.uniform(0, 1, (n, 2))
        s["rotation_point"][rng.random(n) < 0.2] = 0.0            # pivot (0, 0) with the sprite at the origin: Vector2::rotate's other branch
        at_origin = rng.random(n) < 0.1
        s["x"][at_origin] = 0; s["y"][at_origin] = 0
        s["rotation_angle"] = np.where(rng.random(n) < 0.3, 0.0, rng.uniform(-7, 7, n))
    if clipped:
        flags[kind >= 4] = abi.SPRITE_CLIPPED
        s["clip"][:, 0] = rng.uniform(0, 300, n); s["clip"][:, 1] = rng.uniform(0, 200, n)
        s["clip"][:, 2] = rng.uniform(50, 400, n); s["clip"][:, 3] = rng.uniform(50, 300, n)
    s["flags"] = flags
    return s


# -- Font::drawText --
def random_font(seed, glyph_count=95, size=24):
    """A glyph table like a font bundle's: character 32 + k at index k; widths, bearings (some negative) and advances in
    pixels, uvs anywhere in the atlas."""
    rng = np.random.default_rng(seed)
    g = np.zeros(glyph_count, dtype=abi.GLYPH_DTYPE)
    g["code"] = 32 + np.arange(glyph_count)
    g["width"] = rng.integers(1, size + 6, glyph_count)
    g["bearing_x"] = rng.integers(-3, 4, glyph_count)
    g["advance"] = rng.integers(1, size + 8, glyph_count)
    g["uvs"][:, 0] = rng.uniform(0, 0.9, glyph_count); g["uvs"][:, 2] = g["uvs"][:, 0] + rng.uniform(0.01, 0.1, glyph_count)
    g["uvs"][:, 1] = rng.uniform(0.1, 1, glyph_count); g["uvs"][:, 3] = g["uvs"][:, 1] - rng.uniform(0.01, 0.1, glyph_count)
    return g


def random_text(n, seed, newline_every=60):
    """n bytes: mostly printable ASCII, with spaces, tabs, '\\n', '\\r', control bytes, NULs and bytes >= 128 (which the
    reference's signed `char` skips) mixed in."""
    rng = np.random.default_rng(seed)
    t = rng.integers(33, 127, n).astype(np.uint8)
    r = rng.random(n)
    t[r < 0.15] = 32
    t[(r >= 0.15) & (r < 0.17)] = 9
    t[(r >= 0.17) & (r < 0.17 + 1.0 / newline_every)] = 10
    t[(r >= 0.20) & (r < 0.205)] = 13
    t[(r >= 0.21) & (r < 0.215)] = rng.integers(128, 256, n)[(r >= 0.21) & (r < 0.215)]
    t[(r >= 0.22) & (r < 0.223)] = rng.integers(0, 9, n)[(r >= 0.22) & (r < 0.223)]
    return t.tobytes()


def random_ui_sprites(n, seed):
    """n t3d_ui_sprite records: every offset combination, anchors, node rotations (none, about z, about a tilted axis, not
    normalised), node scales (1 and not), both flips, frames inside a 512 x 256 texture."""
    rng = np.random.default_rng(seed)
    s = np.zeros(n, dtype=abi.UI_SPRITE_DTYPE)
    s["node_translation"] = rng.uniform(-200, 900, (n, 3))
    s["camera_translation"] = np.where(rng.random((n, 1)) < 0.3, 0.0, rng.uniform(-50, 50, (n, 2)))
    s["width"] = rng.uniform(1, 300, n); s["height"] = rng.uniform(1, 200, n)
    s["offset"] = rng.choice([0x01, 0x02, 0x04], n) | rng.choice([0x10, 0x20, 0x40], n) | np.where(rng.random(n) < 0.3, 0x80, 0)
    s["anchor"] = rng.uniform(0, 1, (n, 2))
    kind = rng.integers(0, 4, n)
    q = np.zeros((n, 4), dtype=np.float32); q[:, 3] = 1.0
    half = rng.uniform(-3, 3, n) * 0.5
    q[kind == 1, 2] = np.sin(half[kind == 1]); q[kind == 1, 3] = np.cos(half[kind == 1])                  # about z
    axis = rng.normal(size=(n, 3)); axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    q[kind == 2, :3] = (axis * np.sin(half)[:, None])[kind == 2]; q[kind == 2, 3] = np.cos(half[kind == 2])  # about a tilted axis
    q[kind == 3] = rng.uniform(-2, 2, (int((kind == 3).sum()), 4))                                       # not normalised
    s["node_rotation"] = q
    s["node_scale"] = np.where(rng.random((n, 2)) < 0.4, 1.0, rng.uniform(0.25, 3, (n, 2)))
    s["flip"] = rng.integers(0, 4, n)
    s["frame"][:, 0] = rng.integers(0, 400, n); s["frame"][:, 1] = rng.integers(0, 200, n)
    s["frame"][:, 2] = rng.integers(1, 112, n); s["frame"][:, 3] = rng.integers(1, 56, n)
    s["color"] = rng.uniform(0, 1, (n, 4)); s["opacity"] = rng.uniform(0, 1, n)
    return s
