"""ctypes view of include/t3d_particles.h (the C ABI of libt3d_particles.so).

Only declarations live here: the structure layout, the column order, the enums and the function
prototypes.  The library itself is CUDA; there is no Python or CPU implementation behind these names
and `load()` raises if the shared object is missing.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# T3D_LIBRARY: another build of the same library (experiments: scripts/gpu/knobs.sh); the product is the in-tree file
LIB_PATH = os.environ.get("T3D_LIBRARY") or os.path.join(HERE, "csrc", "libt3d_particles.so")

# t3d_status
OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_RNG_UNDERFLOW, ERR_IO, ERR_NO_NODE, ERR_INTERNAL = range(8)
# blend modes (reference ParticleEmitter.h:157-163)
BLEND_NONE, BLEND_ALPHA, BLEND_ADDITIVE, BLEND_MULTIPLIED = range(4)
# draw sources
RNG_LIBC, RNG_STREAM, RNG_DEVICE = range(3)
# billboard rotation modes
ROT_FROZEN, ROT_PER_PARTICLE = range(2)
# vertex layouts of draw
VERTEX_WORLD, VERTEX_CLIP = range(2)
FLOATS_PER_CLIP_QUAD = 40

# column order of the struct-of-arrays particle record
COL_COLOR_START, COL_COLOR_END, COL_COLOR = 0, 4, 8
COL_POSITION, COL_VELOCITY, COL_ACCELERATION, COL_ROTATION_AXIS = 12, 15, 18, 21
COL_ENERGY_START, COL_ENERGY = 24, 25
COL_SIZE_START, COL_SIZE_END, COL_SIZE = 26, 27, 28
COL_ROT_PER_PARTICLE_SPEED, COL_ROTATION_SPEED, COL_ANGLE, COL_TIME_ON_FRAME, COL_FRAME = 29, 30, 31, 32, 33
NUM_COLUMNS = 34
INT_COLUMNS = (COL_ENERGY_START, COL_ENERGY, COL_FRAME)
FLOATS_PER_QUAD = 36

COLUMN_NAMES = (
    ["color_start." + c for c in "rgba"] + ["color_end." + c for c in "rgba"] + ["color." + c for c in "rgba"]
    + ["position." + c for c in "xyz"] + ["velocity." + c for c in "xyz"] + ["acceleration." + c for c in "xyz"]
    + ["rotation_axis." + c for c in "xyz"]
    + ["energy_start", "energy", "size_start", "size_end", "size", "rotation_per_particle_speed",
       "rotation_speed", "angle", "time_on_current_frame", "frame"]
)

F3 = C.c_float * 3
F4 = C.c_float * 4
M16 = C.c_float * 16


class EmitterParams(C.Structure):
    """t3d_emitter_params: the reference's private configuration members (ParticleEmitter.h:824-877)."""

    _fields_ = [
        ("particle_count_max", C.c_uint32),
        ("emission_rate", C.c_uint32),
        ("ellipsoid", C.c_int32),
        ("size_start_min", C.c_float), ("size_start_max", C.c_float),
        ("size_end_min", C.c_float), ("size_end_max", C.c_float),
        ("energy_min", C.c_float), ("energy_max", C.c_float),
        ("color_start", F4), ("color_start_var", F4), ("color_end", F4), ("color_end_var", F4),
        ("position", F3), ("position_var", F3),
        ("velocity", F3), ("velocity_var", F3),
        ("acceleration", F3), ("acceleration_var", F3),
        ("rotation_per_particle_speed_min", C.c_float), ("rotation_per_particle_speed_max", C.c_float),
        ("rotation_speed_min", C.c_float), ("rotation_speed_max", C.c_float),
        ("rotation_axis", F3), ("rotation_axis_var", F3),
        ("orbit_position", C.c_int32), ("orbit_velocity", C.c_int32), ("orbit_acceleration", C.c_int32),
        ("sprite_animated", C.c_int32), ("sprite_looped", C.c_int32),
        ("sprite_frame_random_offset", C.c_uint32),
        ("sprite_frame_duration", C.c_int64),
        ("texture_width", C.c_float), ("texture_height", C.c_float),
        ("blend_mode", C.c_int32),
    ]

    def as_dict(self):
        out = {}
        for name, typ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for name, typ in cls._fields_:
            if name not in d:
                continue
            v = d[name]
            if isinstance(v, (list, tuple)):
                setattr(p, name, typ(*v))
            else:
                setattr(p, name, v)
        return p

    def copy(self):
        q = EmitterParams()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(self))
        return q


def default_params(**overrides):
    """Reference defaults (ParticleEmitter.h:824-877; capacity 100 and rate 10 as ParticleEmitter.cpp:133-143)."""
    d = dict(
        particle_count_max=100, emission_rate=10, ellipsoid=0,
        size_start_min=1.0, size_start_max=1.0, size_end_min=1.0, size_end_max=1.0,
        energy_min=1000.0, energy_max=1000.0,
        color_start=[0, 0, 0, 0], color_start_var=[0, 0, 0, 0], color_end=[1, 1, 1, 1], color_end_var=[0, 0, 0, 0],
        position=[0, 0, 0], position_var=[0, 0, 0], velocity=[0, 0, 0], velocity_var=[1, 1, 1],
        acceleration=[0, 0, 0], acceleration_var=[0, 0, 0],
        rotation_per_particle_speed_min=0.0, rotation_per_particle_speed_max=0.0,
        rotation_speed_min=0.0, rotation_speed_max=0.0, rotation_axis=[0, 0, 0], rotation_axis_var=[0, 0, 0],
        orbit_position=0, orbit_velocity=0, orbit_acceleration=0,
        sprite_animated=0, sprite_looped=0, sprite_frame_random_offset=0, sprite_frame_duration=0,
        texture_width=64.0, texture_height=64.0, blend_mode=BLEND_ALPHA,
    )
    d.update(overrides)
    return EmitterParams.from_dict(d)


class Sprite2D(C.Structure):
    """t3d_sprite2d: one SpriteBatch::draw 2-D call (include/t3d_particles.h)."""
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("width", C.c_float), ("height", C.c_float),
                ("u1", C.c_float), ("v1", C.c_float), ("u2", C.c_float), ("v2", C.c_float), ("color", C.c_float * 4),
                ("rotation_point", C.c_float * 2), ("rotation_angle", C.c_float), ("flags", C.c_uint32), ("clip", C.c_float * 4)]


SPRITE_ROTATED, SPRITE_CENTERED, SPRITE_CLIPPED = 1, 2, 4
SPRITES_ON_DEVICE, SPRITES_NO_CLIP = 1, 2
# numpy view of t3d_ui_sprite (ui Sprite::draw's inputs, 26 words)
UI_SPRITE_DTYPE = [("node_translation", "f4", (3,)), ("camera_translation", "f4", (2,)), ("width", "f4"), ("height", "f4"), ("offset", "u4"),
                   ("anchor", "f4", (2,)), ("node_rotation", "f4", (4,)), ("node_scale", "f4", (2,)), ("flip", "u4"), ("frame", "f4", (4,)),
                   ("color", "f4", (4,)), ("opacity", "f4")]

# numpy view of an array of t3d_glyph (Font::Glyph, 8 words each)
GLYPH_DTYPE = [("code", "u4"), ("width", "u4"), ("bearing_x", "i4"), ("advance", "u4"), ("uvs", "f4", (4,))]
TEXT_ON_DEVICE = 1

# numpy view of an array of t3d_sprite2d (21 words each)
SPRITE2D_DTYPE = [("x", "f4"), ("y", "f4"), ("z", "f4"), ("width", "f4"), ("height", "f4"), ("u1", "f4"), ("v1", "f4"), ("u2", "f4"),
                  ("v2", "f4"), ("color", "f4", 4), ("rotation_point", "f4", 2), ("rotation_angle", "f4"), ("flags", "u4"), ("clip", "f4", 4)]

_P = C.POINTER
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/t3d_particles.h declares.
PROTOTYPES = {
    "t3d_emitter_params_default": (None, [_P(EmitterParams)]),
    "t3d_abi_version": (C.c_int, []),
    "t3d_last_error": (C.c_char_p, []),
    "t3d_ctx_create": (C.c_int, [C.c_int, _vp, _P(_vp)]),
    "t3d_ctx_destroy": (C.c_int, [_vp]),
    "t3d_ctx_flush": (C.c_int, [_vp]),
    "t3d_ctx_sync": (C.c_int, [_vp]),
    "t3d_ctx_set_rotation_mode": (C.c_int, [_vp, C.c_int]),
    "t3d_ctx_set_vertex_mode": (C.c_int, [_vp, C.c_int]),
    "t3d_ctx_copy_to_host_async": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _P(C.c_uint64)]),
    "t3d_ctx_copy_wait": (C.c_int, [_vp, C.c_uint64]),
    "t3d_ctx_set_frozen_rotation": (C.c_int, [_vp, _P(C.c_float), C.c_float]),
    "t3d_ctx_get_frozen_rotation": (C.c_int, [_vp, _P(C.c_float), _P(C.c_int)]),
    "t3d_ctx_reset_statics": (C.c_int, [_vp]),
    "t3d_ctx_set_private_statics": (C.c_int, [_vp, C.c_int]),
    "t3d_ctx_last_kernel_ms": (C.c_int, [_vp, C.c_int, _P(C.c_float)]),
    "t3d_ctx_kernel_time": (C.c_int, [_vp, C.c_int, _P(C.c_double), _P(C.c_uint64)]),
    "t3d_ctx_launch_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "t3d_ctx_host_profile": (C.c_int, [_vp, _P(C.c_double)]),
    "t3d_ctx_transfer_bytes": (C.c_int, [_vp, _P(C.c_uint64), _P(C.c_uint64)]),
    "t3d_emitter_create": (C.c_int, [_vp, _P(EmitterParams), _P(_vp)]),
    "t3d_emitter_create_from_file": (C.c_int, [_vp, C.c_char_p, _P(_vp)]),
    "t3d_emitter_clone": (C.c_int, [_vp, _P(_vp)]),
    "t3d_emitter_destroy": (C.c_int, [_vp]),
    "t3d_emitter_set_params": (C.c_int, [_vp, _P(EmitterParams)]),
    "t3d_emitter_get_params": (C.c_int, [_vp, _P(EmitterParams)]),
    "t3d_emitter_set_emission_rate": (C.c_int, [_vp, C.c_uint32]),
    "t3d_emitter_set_sprite_tex_coords": (C.c_int, [_vp, C.c_uint32, _P(C.c_float)]),
    "t3d_emitter_set_sprite_frame_rects": (C.c_int, [_vp, C.c_uint32, _P(C.c_float)]),
    "t3d_emitter_set_sprite_frame_coords": (C.c_int, [_vp, C.c_uint32, C.c_int, C.c_int]),
    "t3d_emitter_get_sprite_tex_coords": (C.c_int, [_vp, _P(C.c_uint32), _P(C.c_float), C.c_size_t]),
    "t3d_emitter_set_rng": (C.c_int, [_vp, C.c_int, C.c_uint64]),
    "t3d_emitter_feed_draws": (C.c_int, [_vp, _P(C.c_int32), C.c_size_t]),
    "t3d_emitter_pending_draws": (C.c_int, [_vp, _P(C.c_size_t)]),
    "t3d_emitter_set_node_world": (C.c_int, [_vp, _P(C.c_float)]),
    "t3d_emitter_set_camera_world": (C.c_int, [_vp, _P(C.c_float)]),
    "t3d_emitter_set_view_projection": (C.c_int, [_vp, _P(C.c_float)]),
    "t3d_emitter_draw_to": (C.c_int, [_vp, _vp, C.c_size_t, _P(C.c_uint32)]),
    "t3d_emitter_enable_bounds": (C.c_int, [_vp, C.c_int]),
    "t3d_emitter_bounds": (C.c_int, [_vp, _P(C.c_float), _P(C.c_float), _P(C.c_int)]),
    "t3d_emitter_set_culling": (C.c_int, [_vp, C.c_int]),
    "t3d_emitter_start": (C.c_int, [_vp]),
    "t3d_emitter_stop": (C.c_int, [_vp]),
    "t3d_emitter_is_started": (C.c_int, [_vp, _P(C.c_int)]),
    "t3d_emitter_is_active": (C.c_int, [_vp, _P(C.c_int)]),
    "t3d_emitter_emit_once": (C.c_int, [_vp, C.c_uint32]),
    "t3d_emitter_update": (C.c_int, [_vp, C.c_float]),
    "t3d_emitter_draw": (C.c_int, [_vp, _P(C.c_uint32)]),
    "t3d_emitter_get_count": (C.c_int, [_vp, _P(C.c_uint32)]),
    "t3d_batch_update": (C.c_int, [_P(_vp), C.c_size_t, C.c_float]),
    "t3d_batch_draw": (C.c_int, [_P(_vp), C.c_size_t]),
    "t3d_batch_emit_once": (C.c_int, [_P(_vp), C.c_size_t, _P(C.c_uint32)]),
    "t3d_batch_get_counts": (C.c_int, [_P(_vp), C.c_size_t, _P(C.c_uint32)]),
    "t3d_font_create": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint32, _P(_vp)]),
    "t3d_font_destroy": (C.c_int, [_vp]),
    "t3d_font_set_character_spacing": (C.c_int, [_vp, C.c_float]),
    "t3d_font_draw_text": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, C.c_int, C.c_int, _P(C.c_float), C.c_uint32, C.c_int, _vp, C.c_size_t,
                                     _P(_vp), _P(C.c_uint32)]),
    "t3d_batch_frame": (C.c_int, [_P(_vp), C.c_size_t, _P(C.c_float), _P(C.c_float), C.c_float, _P(C.c_uint32)]),
    "t3d_batch_frame_pipelined": (C.c_int, [_P(_vp), C.c_size_t, _P(C.c_float), _P(C.c_float), C.c_float, _P(C.c_uint32), _P(C.c_int)]),
    "t3d_batch_set_transforms": (C.c_int, [_P(_vp), C.c_size_t, _P(C.c_float), _P(C.c_float)]),
    "t3d_emitter_vertices": (C.c_int, [_vp, _P(_vp), _P(C.c_uint32)]),
    "t3d_emitter_depth_sorted_indices": (C.c_int, [_vp, _P(C.c_float), _P(_vp), _P(C.c_uint64)]),
    "t3d_emitter_download_depth_sorted_indices": (C.c_int, [_vp, _P(C.c_float), _P(C.c_uint32), C.c_uint64, _P(C.c_uint64)]),
    "t3d_ctx_strip_indices": (C.c_int, [_vp, C.c_uint32, _P(_vp), _P(C.c_uint64)]),
    "t3d_ctx_triangle_indices": (C.c_int, [_vp, C.c_uint32, _P(_vp), _P(C.c_uint64)]),
    "t3d_ctx_download_triangle_indices": (C.c_int, [_vp, C.c_uint32, _P(C.c_uint32), C.c_uint64, _P(C.c_uint64)]),
    "t3d_ctx_download_strip_indices": (C.c_int, [_vp, C.c_uint32, _P(C.c_uint32), C.c_uint64, _P(C.c_uint64)]),
    "t3d_emitter_download_vertices": (C.c_int, [_vp, _P(C.c_float), C.c_size_t, _P(C.c_uint32)]),
    "t3d_emitter_download_state": (C.c_int, [_vp, _P(C.c_uint32), C.c_size_t, _P(C.c_uint32)]),
    "t3d_emitter_upload_state": (C.c_int, [_vp, _P(C.c_uint32), C.c_size_t, C.c_uint32]),
    "t3d_ctx_draw_sprites": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _vp, C.c_size_t, _P(_vp), _P(C.c_uint32)]),
    "t3d_ctx_map_sprites": (C.c_int, [_vp, C.c_size_t, _P(_vp)]),
    "t3d_tileset_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_uint32, C.c_uint32, _P(_vp)]),
    "t3d_tileset_destroy": (C.c_int, [_vp]),
    "t3d_tileset_set_tile_source": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float]),
    "t3d_tileset_set_tiles": (C.c_int, [_vp, _P(C.c_float)]),
    "t3d_tileset_draw": (C.c_int, [_vp, _P(C.c_float), _P(C.c_float), _vp, C.c_size_t, _P(_vp), _P(C.c_uint32)]),
    "t3d_ctx_download_quads": (C.c_int, [_vp, _vp, C.c_uint32, _P(C.c_float)]),
    "t3d_device_rng_draw": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_uint32]),
    "t3d_host_fuse_policy": (C.c_int, [C.c_uint64, C.c_uint64]),
    "t3d_host_rate_cap": (C.c_int, [_P(C.c_double), C.c_float, _P(C.c_float)]),
    "t3d_host_emission_count": (C.c_uint32, [_P(C.c_float), C.c_float, C.c_float]),
    "t3d_host_ui_sprite_record": (None, [_vp, C.c_uint32, C.c_uint32, _vp]),
    "t3d_host_font_measure_text": (None, [_vp, C.c_uint32, C.c_uint32, C.c_float, C.c_char_p, C.c_size_t, C.c_uint32, _P(C.c_uint32), _P(C.c_uint32)]),
    "t3d_host_sprite_frame_coords": (C.c_uint32, [C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_float, _P(C.c_float)]),
    "t3d_host_particle_draws": (C.c_size_t, [_P(C.c_int32), C.c_size_t, C.c_int, C.c_uint32]),
    "t3d_host_png_size": (C.c_int, [C.c_char_p, _P(C.c_uint32), _P(C.c_uint32)]),
    "t3d_particle_file_parse": (C.c_int, [C.c_char_p, _P(EmitterParams), _P(C.c_uint32), _P(C.c_int), _P(C.c_int), C.c_char_p, C.c_size_t]),
    "t3d_emitter_create_from_properties": (C.c_int, [_vp, _vp, _P(_vp)]),
    "t3d_emitter_set_texture_size": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_int]),
    "t3d_properties_load": (C.c_int, [C.c_char_p, _P(_vp)]),
    "t3d_properties_new": (C.c_int, [C.c_char_p, C.c_char_p, _P(_vp)]),
    "t3d_properties_free": (None, [_vp]),
    "t3d_properties_set": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "t3d_properties_add_namespace": (C.c_int, [_vp, C.c_char_p, C.c_char_p, _P(_vp)]),
    "t3d_properties_namespace": (C.c_char_p, [_vp]),
    "t3d_properties_id": (C.c_char_p, [_vp]),
    "t3d_properties_namespace_count": (C.c_size_t, [_vp]),
    "t3d_properties_get_namespace": (_vp, [_vp, C.c_size_t]),
    "t3d_properties_count": (C.c_size_t, [_vp]),
    "t3d_properties_key": (C.c_char_p, [_vp, C.c_size_t]),
    "t3d_properties_value": (C.c_char_p, [_vp, C.c_size_t]),
    "t3d_properties_get": (C.c_char_p, [_vp, C.c_char_p]),
    "t3d_properties_get_bool": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "t3d_properties_get_int": (C.c_int, [_vp, C.c_char_p]),
    "t3d_properties_get_long": (C.c_longlong, [_vp, C.c_char_p]),
    "t3d_properties_get_float": (C.c_float, [_vp, C.c_char_p]),
    "t3d_properties_get_vector": (C.c_int, [_vp, C.c_char_p, _P(C.c_float), C.c_int]),
    "t3d_properties_get_path": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.c_size_t]),
    "t3d_particle_properties_parse": (C.c_int, [_vp, _P(EmitterParams), _P(C.c_uint32), _P(C.c_int), _P(C.c_int), C.c_char_p, C.c_size_t]),
    "t3d_particle_file_save": (C.c_int, [C.c_char_p, C.c_char_p, _P(EmitterParams), C.c_uint32, C.c_int, C.c_int, C.c_char_p]),
}

_lib = None


def load(path=None):
    """Loads libt3d_particles.so and applies the prototypes.  Raises if the extension was not built:
    there is no fallback implementation."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). tractor3d_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(p, mode=C.RTLD_LOCAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
