/*
 * t3d_particles.h -- C ABI of the B200-native particle backend (libt3d_particles.so).
 *
 * This is the drop-in boundary for Tractor3D's ParticleEmitter hot path: per-frame
 * emitOnce (spawn) -> update (integrate + lerp + sprite step + dead-particle removal) -> draw
 * (billboard quads in SpriteBatch's vertex layout).  The reference has no FFI for this path
 * (ParticleEmitter is a concrete C++ class), so each entry point names the reference member it
 * replaces; "PE.cpp"/"PE.h" = Tractor3Dlib/{src,include}/graphics/ParticleEmitter.{cpp,h},
 * "SB.cpp"/"SB.h" = .../graphics/SpriteBatch.{cpp,h}, "MB.cpp" = .../graphics/MeshBatch.cpp.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no engine types, no torch types.
 *   - every function returns a t3d_status (0 = ok) and never exits the process; the text of the
 *     last failure on the calling thread is available from t3d_last_error().  The C++ facade
 *     (tractor3d_b200/host/ParticleEmitter.h) maps non-zero codes to GP_ERROR / GP_WARN
 *     (reference include/pch.h:83-103).
 *   - one t3d_ctx per GPU, driven from one host thread (the reference path is single-threaded and
 *     non-reentrant: function-local statics at PE.cpp:720, PE.cpp:855, SB.cpp:337-339).
 *   - emit/update/draw ENQUEUE work: consecutive calls on different emitters are coalesced into
 *     one kernel launch per phase and flushed when a result is needed (get_count, download_*,
 *     t3d_ctx_flush, t3d_ctx_sync) or when the phase changes for an emitter already queued.
 *   - time is in milliseconds exactly as in the reference (PE.h:724, PE.h:415-420).
 *   - matrices are column-major float[16] (reference include/math/Matrix.h:62).
 *   - there is no CPU fallback: without a CUDA device t3d_ctx_create fails with T3D_ERR_CUDA.
 */
#ifndef T3D_PARTICLES_H
#define T3D_PARTICLES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T3D_ABI_VERSION 4 /* 4: + t3d_batch_frame_pipelined, t3d_font_* (additive) */

typedef enum t3d_status {
    T3D_OK = 0,
    T3D_ERR_INVALID = 1,      /* bad argument (the reference would assert, e.g. PE.cpp:264, 289) */
    T3D_ERR_CUDA = 2,         /* a CUDA call failed; t3d_last_error() names it */
    T3D_ERR_CAPACITY = 3,     /* growth beyond the capacity fixed at create (PE.h:237 rewrites the field only) */
    T3D_ERR_RNG_UNDERFLOW = 4,/* T3D_RNG_STREAM: not enough recorded draws were fed for this emit */
    T3D_ERR_IO = 5,           /* .particle file missing or malformed (PE.cpp:79-83, 94-115) */
    T3D_ERR_NO_NODE = 6,      /* emitOnce/draw before a node world matrix was supplied (PE.cpp:289, 858) */
    T3D_ERR_INTERNAL = 7      /* a kernel found its grid too small for the live count it read: the host's upper bound was
                                 wrong (a bug in this library); the context refuses further work instead of dropping particles */
} t3d_status;

typedef struct t3d_ctx t3d_ctx;         /* one per GPU */
typedef struct t3d_emitter t3d_emitter; /* replaces one tractor::ParticleEmitter */
typedef struct t3d_properties t3d_properties; /* one namespace of a tractor::Properties tree (see "Properties" below) */

/* Blend modes, PE.h:157-163.  Carried for API fidelity; rasterisation is out of scope. */
enum { T3D_BLEND_NONE = 0, T3D_BLEND_ALPHA = 1, T3D_BLEND_ADDITIVE = 2, T3D_BLEND_MULTIPLIED = 3 };

/* Source of the emitter's random draws (the reference calls libc rand(), PE.cpp:359 and
 * include/pch.h:151-152).  Every draw is a raw integer in [0, 2^31-1]. */
enum {
    T3D_RNG_LIBC = 0,   /* the library calls rand() itself, in the reference's order: same stream as the reference under the same srand() */
    T3D_RNG_STREAM = 1, /* draws recorded from the reference are fed with t3d_emitter_feed_draws() */
    T3D_RNG_DEVICE = 2  /* counter-based generator evaluated in the spawn kernel: draw j of particle serial s = t3d_device_rng_draw(seed, s, j) */
};

/* Billboard rotation behaviour in draw (SB.cpp:326-352). */
enum {
    T3D_ROT_FROZEN = 0,      /* reference-identical: one rotation matrix per process, latched from the first rotated quad (SB.cpp:339) */
    T3D_ROT_PER_PARTICLE = 1 /* intended behaviour: createRotation(right x up, angle_i) for every quad */
};

/* What draw writes per vertex. */
enum {
    T3D_VERTEX_WORLD = 0, /* SpriteBatch::SpriteVertex in world space, as the reference hands it to MeshBatch (SB.cpp:355-362) */
    T3D_VERTEX_CLIP = 1   /* t3d_clip_vertex: position already multiplied by the node's view-projection matrix, i.e. the
                             vertex shader's one line (res/shaders/sprite.vert:19, matrix set at PE.cpp:846-849) */
};

/* Emitter configuration = the reference's private configuration members, PE.h:824-877, in the
 * numeric types the reference stores them in (energy bounds are float there, PE.h:840-841). */
typedef struct t3d_emitter_params {
    uint32_t particle_count_max;   /* _particleCountMax; storage is sized for the value at create: it may be lowered (and raised back) later, never grown */
    uint32_t emission_rate;        /* _emissionRate, particles/s (PE.cpp:262-267) */
    int32_t ellipsoid;             /* _ellipsoid */
    float size_start_min, size_start_max, size_end_min, size_end_max;
    float energy_min, energy_max;  /* ms */
    float color_start[4], color_start_var[4], color_end[4], color_end_var[4];
    float position[3], position_var[3];
    float velocity[3], velocity_var[3];
    float acceleration[3], acceleration_var[3];
    float rotation_per_particle_speed_min, rotation_per_particle_speed_max;
    float rotation_speed_min, rotation_speed_max;
    float rotation_axis[3], rotation_axis_var[3];
    int32_t orbit_position, orbit_velocity, orbit_acceleration;
    int32_t sprite_animated, sprite_looped;
    uint32_t sprite_frame_random_offset; /* _spriteFrameRandomOffset (unsigned in PE.h:867) */
    int64_t sprite_frame_duration;       /* ms; _spriteFrameDurationSecs = (float)d / 1000.0f (PE.cpp:491-495) */
    float texture_width, texture_height; /* _spriteTextureWidth/Height (PE.cpp:244-245) */
    int32_t blend_mode;
} t3d_emitter_params;

/* Fills *p with the reference's defaults (PE.h:824-877; capacity 100, rate 10 as PE.cpp:133-143). */
void t3d_emitter_params_default(t3d_emitter_params* p);

/* Per-particle record as 34 columns of 32-bit words (struct-of-arrays; replaces
 * ParticleEmitter::Particle, PE.h:792-822).  Energies are int32 on the device (the reference's
 * `long` holds milliseconds; values must stay below 2^31). */
enum {
    T3D_COL_COLOR_START = 0,  /* 4 columns r,g,b,a */
    T3D_COL_COLOR_END = 4,    /* 4 */
    T3D_COL_COLOR = 8,        /* 4 */
    T3D_COL_POSITION = 12,    /* 3 columns x,y,z */
    T3D_COL_VELOCITY = 15,    /* 3 */
    T3D_COL_ACCELERATION = 18,/* 3 */
    T3D_COL_ROTATION_AXIS = 21,/* 3 */
    T3D_COL_ENERGY_START = 24,/* int32 */
    T3D_COL_ENERGY = 25,      /* int32 */
    T3D_COL_SIZE_START = 26,
    T3D_COL_SIZE_END = 27,
    T3D_COL_SIZE = 28,
    T3D_COL_ROT_PER_PARTICLE_SPEED = 29,
    T3D_COL_ROTATION_SPEED = 30,
    T3D_COL_ANGLE = 31,
    T3D_COL_TIME_ON_FRAME = 32,
    T3D_COL_FRAME = 33,       /* uint32 */
    T3D_NUM_COLUMNS = 34
};

/* Output vertex, identical to SpriteBatch::SpriteVertex (SB.h:360-380): 36 bytes, 4 per quad in the
 * order SB.cpp:355-359 emits them.  The triangle-strip index pattern of MB.cpp:126-141 is implicit
 * ({4i..4i+3} stitched by two degenerate indices) and is not materialised per frame. */
typedef struct t3d_sprite_vertex { float x, y, z, u, v, r, g, b, a; } t3d_sprite_vertex;
#define T3D_FLOATS_PER_QUAD 36
/* T3D_VERTEX_CLIP: gl_Position = u_projectionMatrix * vec4(a_position, 1) (sprite.vert:19) evaluated per vertex,
 * each component as the left-to-right sum x*m[r] + y*m[4+r] + z*m[8+r] + 1*m[12+r] (MathUtil.h:195-200). */
typedef struct t3d_clip_vertex { float x, y, z, w, u, v, r, g, b, a; } t3d_clip_vertex;
#define T3D_FLOATS_PER_CLIP_QUAD 40

/* ---- library / context ------------------------------------------------------------------ */
int t3d_abi_version(void);
const char* t3d_last_error(void);

/* Creates the per-GPU context.  `cuda_stream` may be NULL (the context creates its own stream) or a
 * cudaStream_t owned by the caller on which all kernels and copies are then issued. */
int t3d_ctx_create(int device, void* cuda_stream, t3d_ctx** out);
int t3d_ctx_destroy(t3d_ctx* ctx);
int t3d_ctx_flush(t3d_ctx* ctx);  /* launch everything queued; does not wait */
int t3d_ctx_sync(t3d_ctx* ctx);   /* flush, then wait for the stream */
int t3d_ctx_set_rotation_mode(t3d_ctx* ctx, int mode);            /* T3D_ROT_*; default FROZEN */
/* T3D_VERTEX_*; default WORLD.  In CLIP mode draw() multiplies every vertex by the matrix given with
 * t3d_emitter_set_view_projection (update and draw then go out as two launches: the fused pass writes SpriteVertex). */
int t3d_ctx_set_vertex_mode(t3d_ctx* ctx, int mode);
/* Latches the process-wide billboard rotation explicitly, as if the first rotated quad ever drawn
 * had axis u = right x up and this angle (SB.cpp:337-339). */
int t3d_ctx_set_frozen_rotation(t3d_ctx* ctx, const float u[3], float angle);
/* Returns the latched matrix (column-major 4x4) and whether it is latched yet. */
int t3d_ctx_get_frozen_rotation(t3d_ctx* ctx, float m[16], int* latched);
/* Forgets the function-local statics the reference keeps per process: the update rate-cap
 * accumulator (PE.cpp:720) and the latched rotation (SB.cpp:339) -- i.e. "a fresh process".
 * These two live ONCE PER PROCESS here as well: all contexts of a process (one per GPU) share them, so a host that drives
 * N contexts from its game thread gets the reference's results however the emitters are split over the GPUs. */
int t3d_ctx_reset_statics(t3d_ctx* ctx);
/* on: this context keeps its own copy of those statics (fresh at the switch) instead of the process-wide one -- for hosts
 * that drive their contexts from different threads (the shared copy is not synchronised), at the price that results with
 * update steps under 8 ms, and the frame that latches the rotation, then depend on the split.  Off by default. */
int t3d_ctx_set_private_statics(t3d_ctx* ctx, int on);
/* Device time in ms of the last flushed launch of each kind (CUDA events on the context's stream):
 * which = 0 spawn, 1 update, 2 draw, 3 fused update+draw (the default frame).  Waits for that launch to finish. */
int t3d_ctx_last_kernel_ms(t3d_ctx* ctx, int which, float* ms);
/* Cumulative device time (ms) and number of launches of one kind since the context was created; every
 * launch is bracketed by its own event pair, so a caller can difference two readings around a timed region. */
int t3d_ctx_kernel_time(t3d_ctx* ctx, int which, double* total_ms, uint64_t* n_launches);
/* Number of kernels this library has launched on the context since creation. */
int t3d_ctx_launch_count(t3d_ctx* ctx, uint64_t* n);
/* Bytes staged host->device / device->host by the library since creation. */
int t3d_ctx_transfer_bytes(t3d_ctx* ctx, uint64_t* h2d, uint64_t* d2h);

/* Host time in seconds, since the context was created, that calls on this context have spent inside the library, split
 * by what it went into (each slot exclusive of the others); the batch calls and everything behind flush / read-backs are
 * counted, the per-emitter setters are not.  For finding where a frame's host time goes (scripts/host_cost.py). */
enum {
    T3D_HOST_ENQUEUE = 0, /* t3d_batch_* calls: the per-emitter host logic (isActive, rate cap, emission count) and queueing */
    T3D_HOST_BUILD = 1,   /* cutting launch descriptors and the bookkeeping around a launch */
    T3D_HOST_SUBMIT = 2,  /* inside CUDA calls that enqueue work: descriptor uploads, launches, event records */
    T3D_HOST_WAIT = 3,    /* blocked on the device: count read-backs, t3d_ctx_sync, a full staging ring */
    T3D_HOST_PROFILE_SLOTS = 4
};
int t3d_ctx_host_profile(t3d_ctx* ctx, double seconds[T3D_HOST_PROFILE_SLOTS]);

/* ---- emitter lifecycle (PE.cpp:43-211) --------------------------------------------------- */
/* Replaces ParticleEmitter::create(texture, blendMode, particleCountMax) + the 16 setters
 * (PE.cpp:63-73, 193-208).  Installs one full-texture sprite frame like setTexture (PE.cpp:249-251). */
int t3d_emitter_create(t3d_ctx* ctx, const t3d_emitter_params* params, t3d_emitter** out);
/* Replaces ParticleEmitter::create(url) (PE.cpp:76-211): parses a `.particle` Properties file.
 * The texture is not decoded; only the PNG header is read for width/height (PE.cpp:244-247). */
int t3d_emitter_create_from_file(t3d_ctx* ctx, const char* url, t3d_emitter** out);
/* Replaces ParticleEmitter::clone (PE.cpp:891-932): same configuration, no particles. */
/* Replaces ParticleEmitter::create(Properties*) (PE.cpp:92-211, the overload src/scene/SceneLoader.cpp:320 calls):
 * `particle` is a namespace named "particle" whose first child namespace is "sprite" (t3d_properties below). */
int t3d_emitter_create_from_properties(t3d_ctx* ctx, const t3d_properties* particle, t3d_emitter** out);
int t3d_emitter_clone(t3d_emitter* e, t3d_emitter** out);
int t3d_emitter_destroy(t3d_emitter* e);

/* ---- configuration (PE.cpp:262-267, 372-448, 485-590) ------------------------------------- */
int t3d_emitter_set_params(t3d_emitter* e, const t3d_emitter_params* params); /* capacity must not grow */
int t3d_emitter_get_params(t3d_emitter* e, t3d_emitter_params* params);
int t3d_emitter_set_emission_rate(t3d_emitter* e, uint32_t rate);                   /* PE.cpp:262-267 */
/* Replaces ParticleEmitter::setTexture(Texture*, BlendMode) (PE.cpp:229-252) minus the GL objects: the new texture's size
 * re-derives the texel ratios and the sprite table is reset to ONE frame covering the whole texture -- also when the size
 * did not change (PE.cpp:249-251).  Nothing on the device depends on the texture size (particles carry a frame index). */
int t3d_emitter_set_texture_size(t3d_emitter* e, uint32_t width, uint32_t height, int blend_mode);
int t3d_emitter_set_sprite_tex_coords(t3d_emitter* e, uint32_t frame_count, const float* tex_coords); /* PE.cpp:512-523 */
/* rects = frame_count x {x, y, width, height} in texels (PE.cpp:526-547) */
int t3d_emitter_set_sprite_frame_rects(t3d_emitter* e, uint32_t frame_count, const float* rects);
int t3d_emitter_set_sprite_frame_coords(t3d_emitter* e, uint32_t frame_count, int width, int height); /* PE.cpp:550-582 */
int t3d_emitter_get_sprite_tex_coords(t3d_emitter* e, uint32_t* frame_count, float* tex_coords, size_t max_frames);
int t3d_emitter_set_rng(t3d_emitter* e, int mode, uint64_t seed);
int t3d_emitter_feed_draws(t3d_emitter* e, const int32_t* draws, size_t n);  /* T3D_RNG_STREAM */
int t3d_emitter_pending_draws(t3d_emitter* e, size_t* n);

/* ---- what the reference pulls from its Node / Scene each call ----------------------------- */
/* Node::getWorldMatrix() of the emitter's node (PE.cpp:299).  Supplying it = Node::setDrawable. */
int t3d_emitter_set_node_world(t3d_emitter* e, const float world[16]);
/* getWorldMatrix() of the active camera's node (PE.cpp:860-864); only right=(m0,m1,m2), up=(m4,m5,m6) are used. */
int t3d_emitter_set_camera_world(t3d_emitter* e, const float camera_world[16]);

/* Node::getViewProjectionMatrix() of the emitter's node (PE.cpp:846-849); identity until set.  Used by
 * T3D_VERTEX_CLIP draws and by the frustum test of t3d_emitter_set_culling. */
int t3d_emitter_set_view_projection(t3d_emitter* e, const float view_projection[16]);

/* ---- the hot path ------------------------------------------------------------------------- */
int t3d_emitter_start(t3d_emitter* e);                   /* PE.cpp:270-274 */
int t3d_emitter_stop(t3d_emitter* e);                    /* PE.h:272 */
int t3d_emitter_is_started(t3d_emitter* e, int* out);    /* PE.h:279 */
int t3d_emitter_is_active(t3d_emitter* e, int* out);     /* PE.cpp:277-284 (may read the count back) */
int t3d_emitter_emit_once(t3d_emitter* e, uint32_t particle_count);   /* PE.cpp:287-369 */
int t3d_emitter_update(t3d_emitter* e, float elapsed_ms);             /* PE.cpp:713-832 */
/* PE.cpp:835-888.  *draw_calls receives what the reference returns (0 if inactive, else 1); for a stopped emitter
 * that may need the live count from the device.  Pass NULL (as t3d_batch_draw does) when the value is not needed:
 * the draw is then queued without waiting, and writes no quads if every particle has died. */
int t3d_emitter_draw(t3d_emitter* e, uint32_t* draw_calls);
int t3d_emitter_get_count(t3d_emitter* e, uint32_t* count);           /* PE.h:306; flushes + reads back */
/* draw() with the quads written straight into the caller's device buffer instead of the emitter's own vertex stream:
 * a mapped graphics-interop buffer (cudaGraphicsResourceGetMappedPointer of a GL/Vulkan vertex buffer) or any device
 * allocation.  `device_dst` must be 16-byte aligned and hold `dst_quads` quads of the context's vertex layout; the call
 * fails with T3D_ERR_CAPACITY when the emitter may hold more live particles than that.  Quad i of the frame lands at
 * quad index i, exactly as in t3d_emitter_vertices(); nothing is written beyond the frame's quads. */
int t3d_emitter_draw_to(t3d_emitter* e, void* device_dst, size_t dst_quads, uint32_t* draw_calls);

/* ---- batches: the same calls over many emitters, one launch per phase --------------------- */
int t3d_batch_update(t3d_emitter* const* emitters, size_t n, float elapsed_ms);
int t3d_batch_draw(t3d_emitter* const* emitters, size_t n);
int t3d_batch_emit_once(t3d_emitter* const* emitters, size_t n, const uint32_t* particle_counts);
int t3d_batch_get_counts(t3d_emitter* const* emitters, size_t n, uint32_t* counts);
/* Per-frame scene inputs for many emitters at once: node_worlds = n x float[16] (or NULL to leave them),
 * camera_world = one float[16] shared by all (or NULL).  Same meaning as the two per-emitter setters above. */
int t3d_batch_set_transforms(t3d_emitter* const* emitters, size_t n, const float* node_worlds, const float* camera_world);

/* One frame of the game loop for many emitters in one call: t3d_batch_set_transforms, t3d_batch_update, t3d_batch_draw and
 * -- when `counts` is not NULL -- t3d_batch_get_counts, in that order and with exactly their effects (the update and the
 * draw go out as one fused launch).  What an engine's particle system calls once per frame. */
int t3d_batch_frame(t3d_emitter* const* emitters, size_t n, const float* node_worlds, const float* camera_world, float elapsed_ms,
                    uint32_t* counts);

/* The same frame without the wait: the call returns as soon as the frame is enqueued, and `previous_counts` (n words, or NULL)
 * receives the live counts that the PREVIOUS call of this function left for the same emitters -- their read-back travelled
 * while this frame was being prepared, so the host never idles the device.  `*have_previous` is 0 when there is no such frame
 * (first call, another emitter set, or a frame whose update the 8 ms rate cap of PE.cpp:720-725 swallowed) and
 * `previous_counts` is then left untouched.  State, vertices and the order of effects are exactly t3d_batch_frame's; only the
 * counts reach the caller one frame later (ParticleEmitter::getParticlesCount, PE.h:306, stays exact through
 * t3d_emitter_get_count / t3d_batch_get_counts at any time). */
int t3d_batch_frame_pipelined(t3d_emitter* const* emitters, size_t n, const float* node_worlds, const float* camera_world,
                              float elapsed_ms, uint32_t* previous_counts, int* have_previous);

/* ---- results ------------------------------------------------------------------------------ */
/* Device-resident vertex stream of the last draw: 4*n_quads t3d_sprite_vertex, quad i at [4i, 4i+4)
 * in live-array order (what SpriteBatch::draw appended to MeshBatch, SB.cpp:355-362).  Flushes and
 * reads the quad count back. */
int t3d_emitter_vertices(t3d_emitter* e, const t3d_sprite_vertex** device_ptr, uint32_t* n_quads);
int t3d_emitter_download_vertices(t3d_emitter* e, float* out, size_t max_quads, uint32_t* n_quads);
/* Copies `bytes` from device memory to (pinned) host memory on the context's download stream, ordered after everything
 * enqueued on the context so far, without blocking later launches: frame k's vertices travel over PCIe while frame k+1
 * is computed (alternate between two t3d_emitter_draw_to buffers).  t3d_ctx_copy_wait blocks until that copy has landed. */
int t3d_ctx_copy_to_host_async(t3d_ctx* ctx, void* host_dst, const void* device_src, size_t bytes, uint64_t* ticket);
int t3d_ctx_copy_wait(t3d_ctx* ctx, uint64_t ticket);
/* Bounding box of the emitter's live particles as the last update() left them: min / max over pos -+ size per axis (the
 * engine leaves this out: "TODO: bounds ignore particle emitters", src/scene/Node.cpp:777-780).  Reduced inside the
 * update kernel while the particle is in registers; enable it before the update whose box is wanted.  *valid = 0 when
 * no update has run since enabling or no particle is live. */
int t3d_emitter_enable_bounds(t3d_emitter* e, int on);
int t3d_emitter_bounds(t3d_emitter* e, float lo[3], float hi[3], int* valid);
/* on: draw() of this emitter is skipped (no quads, returns what the reference returns) while its last bounding box lies
 * entirely outside the clip volume of its view-projection matrix.  Implies t3d_emitter_enable_bounds. */
int t3d_emitter_set_culling(t3d_emitter* e, int on);
/* SURVEY 8(f4), back-to-front order for BLEND_ALPHA (the reference draws in array order, PE.cpp:866-882, which only
 * additive blending forgives): a triangle-list index buffer (6 indices per quad, 32-bit, the pattern of
 * t3d_ctx_triangle_indices) that walks the quads of the emitter's current particles by decreasing depth
 * position . view_dir (x*vx + y*vy + z*vz, left to right); quads of equal depth keep their live-array order.  The indices
 * address the vertex stream of a draw() of the same state (quad i = live particle i).  Device-resident, owned by the
 * context, valid until the next call of this function on the context.  Flushes and reads the live count back. */
int t3d_emitter_depth_sorted_indices(t3d_emitter* e, const float view_dir[3], const uint32_t** device_ptr, uint64_t* n_indices);
int t3d_emitter_download_depth_sorted_indices(t3d_emitter* e, const float view_dir[3], uint32_t* out, uint64_t max_indices, uint64_t* n_indices);
/* The index buffer MeshBatch builds for n_quads billboards (TRIANGLE_STRIP, MB.cpp:117-141): quad 0 contributes
 * {0,1,2,3}; every later quad i the degenerate pair {4(i-1)+3, 4i} and then {4i..4i+3}: 6*n_quads-2 indices.
 * 32-bit (the reference's unsigned short indices cap a batch at 10 922 quads, MB.cpp:213-220).  The buffer is
 * cached by the context, device-resident, and valid until the next call that asks for more quads. */
int t3d_ctx_strip_indices(t3d_ctx* ctx, uint32_t n_quads, const uint32_t** device_ptr, uint64_t* n_indices);
/* The same quads as an indexed triangle LIST for renderers without strips: {4i, 4i+1, 4i+2, 4i+2, 4i+1, 4i+3} per quad
 * = the two non-degenerate triangles of quad i's strip with the strip's winding (6*n_quads indices, 32-bit). */
int t3d_ctx_triangle_indices(t3d_ctx* ctx, uint32_t n_quads, const uint32_t** device_ptr, uint64_t* n_indices);
int t3d_ctx_download_triangle_indices(t3d_ctx* ctx, uint32_t n_quads, uint32_t* out, uint64_t max_indices, uint64_t* n_indices);
int t3d_ctx_download_strip_indices(t3d_ctx* ctx, uint32_t n_quads, uint32_t* out, uint64_t max_indices, uint64_t* n_indices);
/* Live particles as T3D_NUM_COLUMNS columns of `stride` words each: out[c*stride + i] for i < *count. */
int t3d_emitter_download_state(t3d_emitter* e, uint32_t* out, size_t stride, uint32_t* count);
int t3d_emitter_upload_state(t3d_emitter* e, const uint32_t* in, size_t stride, uint32_t count);

/* ---- the other SpriteBatch producers -- 2-D quads ---------------------------------------------
 * TileSet::draw (src/graphics/TileSet.cpp:179-236), Font::drawText (src/renderer/Font.cpp:242-420) and the UI controls
 * (src/ui/Control.cpp, Container.cpp) feed SpriteBatch's 2-D overloads, one call per quad.  One t3d_sprite2d = one such
 * call; a list of them expands on the device into the same SpriteVertex stream, in call order, culled sprites leaving no
 * gap.  "SB.cpp" line numbers as above. */
enum {
    T3D_SPRITE_ROTATED = 1,  /* the overloads that take rotationPoint / rotationAngle (SB.cpp:235-289): vertices leave in the
                                order downLeft, upLeft, downRight, upRight and turn about the pivot when the angle is not 0 */
    T3D_SPRITE_CENTERED = 2, /* positionIsCenter (SB.cpp:248-252, :487-491) */
    T3D_SPRITE_CLIPPED = 4   /* the overloads with a clip rectangle (SB.cpp:366-413): clipped against it by clipSprite
                                (SB.cpp:523-597), dropped when entirely outside; not combined with T3D_SPRITE_ROTATED */
};
typedef struct t3d_sprite2d {
    float x, y, z, width, height;   /* destination */
    float u1, v1, u2, v2;           /* source, already in texture coordinates (SB.cpp:176-181 is the caller's) */
    float color[4];
    float rotation_point[2];        /* T3D_SPRITE_ROTATED */
    float rotation_angle;
    uint32_t flags;
    float clip[4];                  /* T3D_SPRITE_CLIPPED: x, y, width, height */
} t3d_sprite2d;
/* how the list arrives */
enum {
    T3D_SPRITES_ON_DEVICE = 1, /* `sprites` is a device array the caller keeps resident (16-byte aligned) */
    T3D_SPRITES_NO_CLIP = 2    /* the caller promises that no sprite carries T3D_SPRITE_CLIPPED: sprite i becomes quad i, one
                                  launch, and the list is not scanned for the flag */
};
/* Expands n sprites.  `sprites` is host memory -- copied through pinned staging, or, when it is the buffer
 * t3d_ctx_map_sprites handed out, used where it lies -- or a device array (T3D_SPRITES_ON_DEVICE).  Quads go to device_dst
 * (16-byte aligned, dst_quads quads; T3D_ERR_CAPACITY if n does not fit) or, when it is NULL, to a buffer the context
 * owns; *device_ptr receives where.  *n_quads (may be NULL) receives the number of quads written, which waits for the
 * device when some sprite may be clipped; otherwise it is n. */
int t3d_ctx_draw_sprites(t3d_ctx* ctx, const t3d_sprite2d* sprites, size_t n, int list_flags, void* device_dst,
                         size_t dst_quads, const t3d_sprite_vertex** device_ptr, uint32_t* n_quads);
/* Pinned host memory for the next list of up to n sprites: a producer (text layout, UI, a tile walk) writes its records
 * straight into it and passes the same pointer to t3d_ctx_draw_sprites, which then uploads without a copy of its own.
 * Two such buffers alternate, so list k + 1 can be written while list k is still on its way to the device. */
int t3d_ctx_map_sprites(t3d_ctx* ctx, size_t n, t3d_sprite2d** host_list);

/* TileSet (include/graphics/TileSet.h): a grid of equally sized tiles cut from one texture; blank cells are skipped.  The
 * tile table lives on the device, so a frame uploads nothing but its position. */
typedef struct t3d_tileset t3d_tileset;
int t3d_tileset_create(t3d_ctx* ctx, uint32_t texture_width, uint32_t texture_height, float tile_width, float tile_height,
                       uint32_t row_count, uint32_t column_count, t3d_tileset** out);            /* TileSet.cpp:31-58: every cell blank */
int t3d_tileset_destroy(t3d_tileset* t);
int t3d_tileset_set_tile_source(t3d_tileset* t, uint32_t column, uint32_t row, float x, float y); /* TileSet.cpp:155-165 */
/* all cells at once: sources = row_count * column_count * {x, y}, row-major; negative = blank (TileSet.cpp:217) */
int t3d_tileset_set_tiles(t3d_tileset* t, const float* sources);
/* TileSet::draw (TileSet.cpp:179-236).  position = what the function derives from its node and the active camera's node
 * (node world translation minus the camera node's x and y, :182-205); color = _color with _opacity folded into alpha (:226).
 * Destination as in t3d_ctx_draw_sprites; *n_quads = the number of non-blank cells (known on the host: no wait). */
int t3d_tileset_draw(t3d_tileset* t, const float position[3], const float color[4], void* device_dst, size_t dst_quads,
                     const t3d_sprite_vertex** device_ptr, uint32_t* n_quads);

/* Font::drawText(text, x, y, color, size) (src/renderer/Font.cpp:242-390): one SpriteBatch quad per character that has a glyph,
 * in reading order, laid out by the reference's pen walk (' ' and '\t' advance by the first glyph's advance, '\r' / '\n'
 * start a new line `size` pixels down, a glyph advances by floor(advance * scale + spacing); bytes >= 128 are skipped: `char`
 * is signed there).  A t3d_font is ONE size of a font (Font::create(family, style, size, glyphs, glyphCount, texture,
 * format), Font.cpp:103-161); choosing among several sizes (findClosestSize, :221-239) stays with the caller. */
typedef struct t3d_glyph {      /* Font::Glyph, include/renderer/Font.h:296-323 */
    uint32_t code;
    uint32_t width;             /* pixels */
    int32_t  bearing_x;
    uint32_t advance;
    float    uvs[4];            /* u1, v1, u2, v2 */
} t3d_glyph;
typedef struct t3d_font t3d_font;
int t3d_font_create(t3d_ctx* ctx, uint32_t size, const t3d_glyph* glyphs, uint32_t glyph_count, t3d_font** out);
int t3d_font_destroy(t3d_font* f);
int t3d_font_set_character_spacing(t3d_font* f, float spacing);                                   /* Font.h:246 */
#define T3D_TEXT_ON_DEVICE 1u   /* `text` is a device pointer (16-byte aligned); nothing is uploaded */
/* size = 0: the font's own size (Font.cpp:253-256).  right_to_left: the reference's other walk (Font.cpp:283-327) -- the
 * blanks a line starts with forwards, then the rest of the line from its last character backwards; the text then ends at
 * its first NUL (c_str()), which device-resident text must not contain.  Destination as in t3d_ctx_draw_sprites; *n_quads =
 * characters drawn (counted on the host for host text; read back from the device, with a wait, for T3D_TEXT_ON_DEVICE). */
int t3d_font_draw_text(t3d_font* f, const char* text, size_t length, uint32_t text_flags, int x, int y, const float color[4],
                       uint32_t size, int right_to_left, void* device_dst, size_t dst_quads, const t3d_sprite_vertex** device_ptr,
                       uint32_t* n_quads);
int t3d_ctx_download_quads(t3d_ctx* ctx, const t3d_sprite_vertex* device_ptr, uint32_t n_quads, float* out); /* tests */

/* ---- pure host helpers (no GPU needed): the scalar logic of the path ----------------------- */
/* Raw 31-bit draw j of particle `serial` for T3D_RNG_DEVICE (same function the spawn kernel evaluates). */
int32_t t3d_device_rng_draw(uint64_t seed, uint64_t serial, uint32_t j);
/* PE.cpp:720-725: the process-wide update rate cap.  Returns 1 and sets *elapsed_ms when the update proceeds. */
int t3d_host_rate_cap(double* running_time, float elapsed_time, float* elapsed_ms);
/* The runtime's launch policy (not in the reference): should update() of a set of emitters wait for the draw() that
 * follows, to go out as one fused launch?  particles = their live particles (upper bound), spawned = particles asked to
 * spawn since their previous update (stands in for deaths).  Always 1 with the current kernels (the fused frame wins at
 * every measured churn level); kept as the one place the decision lives. */
int t3d_host_fuse_policy(uint64_t particles, uint64_t spawned);
/* PE.cpp:732-743: advances *emit_time and returns how many particles to emit this frame. */
uint32_t t3d_host_emission_count(float* emit_time, float time_per_emission, float elapsed_ms);
/* PE.cpp:526-582: uv table for a grid of width x height frames on a tex_w x tex_h texture. Returns frames written. */
uint32_t t3d_host_sprite_frame_coords(uint32_t frame_count, int width, int height, float tex_w, float tex_h, float* tex_coords);
/* Number of draws one spawned particle consumes starting at draws[0] (PE.cpp:312-364); 0 if the
 * stream ends first.  Used to cut a recorded stream into per-particle offsets. */
size_t t3d_host_particle_draws(const int32_t* draws, size_t n, int ellipsoid, uint32_t frame_random_offset);
/* ui Sprite::draw (src/ui/Sprite.cpp:495-574): what a Sprite node makes of its node, the active camera's node and its own
 * offset / anchor / flip state before its ONE SpriteBatch::draw(position, frame, scale, color, anchor, angle) call
 * (SB.cpp:202-216 -> :235-289).  t3d_host_ui_sprite_record fills the t3d_sprite2d that call amounts to (T3D_SPRITE_ROTATED);
 * a scene's sprites go out together as one t3d_ctx_draw_sprites list.  Scalar per-node logic: host, no GPU. */
typedef struct t3d_ui_sprite {
    float    node_translation[3];   /* Node::getTranslationWorld() */
    float    camera_translation[2]; /* x, y of the active camera node's world translation; 0 when the scene has no camera */
    float    width, height;         /* Sprite::getWidth / getHeight */
    uint32_t offset;                /* Sprite::Offset bits (include/graphics/Sprite.h:64-82): 0x02 hcenter, 0x04 right, 0x10 top,
                                       0x20 vcenter, 0x80 anchor */
    float    anchor[2];
    float    node_rotation[4];      /* Node::getRotation(): x, y, z, w */
    float    node_scale[2];         /* Node::getScaleX / getScaleY */
    uint32_t flip;                  /* Sprite::FlipFlags: 1 vertical, 2 horizontal */
    float    frame[4];              /* the current frame's source rectangle in texels: x, y, width, height */
    float    color[4];
    float    opacity;
} t3d_ui_sprite;
void t3d_host_ui_sprite_record(const t3d_ui_sprite* sprite, uint32_t texture_width, uint32_t texture_height, t3d_sprite2d* out);
/* Font::measureText(text, size, &width, &height) (src/renderer/Font.cpp:675-731 with getTokenWidth, :1624-1657) for one size of a
 * font: the widest line and `size` per line -- what a caller needs to place a text before t3d_font_draw_text.  Lines end at
 * '\n' only ('\r' is skipped here, unlike in drawText), the text at its first NUL; size = 0: the font's own.  Scalar walk over
 * the bytes: host logic, no GPU. */
void t3d_host_font_measure_text(const t3d_glyph* glyphs, uint32_t glyph_count, uint32_t font_size, float character_spacing,
                                const char* text, size_t length, uint32_t size, uint32_t* width, uint32_t* height);
/* Width and height of a PNG from its header: all ParticleEmitter ever asks of its texture (PE.cpp:244-247). */
int t3d_host_png_size(const char* path, uint32_t* width, uint32_t* height);
/* ---- Properties: the slice of tractor::Properties (include/scene/Properties.h, src/scene/Properties.cpp) that
 * ParticleEmitter::create reads -- nested namespaces `name [id] { ... }` holding `key = value` strings, typed getters with
 * the reference's conversions.  A handle names one namespace; the root returned by _load / _new owns the whole tree and
 * is the only handle to free.  This is how a caller that already holds a parsed tree (SceneLoader.cpp:318-324) hands a
 * `particle` namespace over without going through a file. */
int t3d_properties_load(const char* url, t3d_properties** out);                       /* Properties::create(url) */
int t3d_properties_new(const char* name_space, const char* id, t3d_properties** out); /* an empty root namespace */
void t3d_properties_free(t3d_properties* root);
int t3d_properties_set(t3d_properties* p, const char* key, const char* value);        /* Properties::setString */
int t3d_properties_add_namespace(t3d_properties* p, const char* name_space, const char* id, t3d_properties** child);
const char* t3d_properties_namespace(const t3d_properties* p);                        /* Properties::getNamespace() */
const char* t3d_properties_id(const t3d_properties* p);                               /* Properties::getId() */
size_t t3d_properties_namespace_count(const t3d_properties* p);
t3d_properties* t3d_properties_get_namespace(t3d_properties* p, size_t index);        /* the walk of getNextNamespace() */
size_t t3d_properties_count(const t3d_properties* p);
const char* t3d_properties_key(const t3d_properties* p, size_t index);                /* the walk of getNextProperty() */
const char* t3d_properties_value(const t3d_properties* p, size_t index);
const char* t3d_properties_get(const t3d_properties* p, const char* key);             /* getString; NULL when absent */
int t3d_properties_get_bool(const t3d_properties* p, const char* key, int default_value); /* Properties.cpp:940-944 */
int t3d_properties_get_int(const t3d_properties* p, const char* key);                 /* :947-962 */
long long t3d_properties_get_long(const t3d_properties* p, const char* key);          /* :981-994 */
float t3d_properties_get_float(const t3d_properties* p, const char* key);             /* :965-978 */
/* getVector3 / getVector4 (:1214-1251): dim = 3 or 4; returns 1 when the key held a well-formed vector, else 0 and zeros. */
int t3d_properties_get_vector(const t3d_properties* p, const char* key, float* out, int dim);
/* getPath: the value as it stands when that file exists, else relative to the directory of the file the tree was loaded
 * from; T3D_ERR_IO when neither exists. */
int t3d_properties_get_path(const t3d_properties* p, const char* key, char* out, size_t out_len);
/* PE.cpp:92-184 on a `particle` namespace: params + sprite sheet description. */
int t3d_particle_properties_parse(const t3d_properties* particle, t3d_emitter_params* params, uint32_t* frame_count,
                                  int* sprite_width, int* sprite_height, char* texture_path, size_t texture_path_len);
/* Parses a `.particle` file into params + sprite sheet description (PE.cpp:92-184). */
int t3d_particle_file_parse(const char* url, t3d_emitter_params* params, uint32_t* frame_count,
                            int* sprite_width, int* sprite_height, char* texture_path, size_t texture_path_len);
/* Writes a `.particle` file in the sample's format (samples/browser/src/ParticlesSample.cpp:340-424). */
int t3d_particle_file_save(const char* url, const char* name, const t3d_emitter_params* params,
                           uint32_t frame_count, int sprite_width, int sprite_height, const char* texture_path);

#ifdef __cplusplus
}
#endif
#endif /* T3D_PARTICLES_H */
