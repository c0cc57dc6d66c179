/*
 * particle_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement, in plain C, of the reference's ParticleEmitter emit/update/draw path
 * (Tractor3Dlib/src/graphics/ParticleEmitter.cpp + SpriteBatch.cpp:292-363 + math/Matrix.cpp:348-406).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product (libt3d_particles.so) never links or calls it.
 *
 * Pinning: the reference has no tests or golden vectors for this path, so the oracle is
 * pinned against the reference ITSELF: oracle/_ref/libt3dref.so is the unmodified reference compiled
 * headless (oracle/Makefile), tests/test_oracle_vs_ref.py compares the two bit-for-bit on shared random
 * streams, and tests/golden/ holds vectors generated from oracle/_ref by tests/golden/make_golden.py.
 * T3D_ROT_PER_PARTICLE (a mode the reference cannot exhibit, SB.cpp:339) is restated, not reference-pinned.
 */
#ifndef PARTICLE_ORACLE_H
#define PARTICLE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/t3d_particles.h" /* t3d_emitter_params, T3D_COL_*, T3D_ROT_* (types only) */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_emitter orc_emitter;

/* ---- the draw source standing in for libc rand() (process-wide, like the reference's) ---- */
void orc_rand_seed(uint64_t seed);                        /* sequential generator, same as oracle/_ref's */
void orc_rand_replay(const int32_t* draws, size_t n);     /* NULL: back to the generator */
size_t orc_rand_replay_remaining(void);
void orc_rand_record(int on);
size_t orc_rand_recorded(int32_t* out, size_t max);
uint64_t orc_rand_calls(void);
/* The counter-based generator of T3D_RNG_DEVICE, restated. */
int32_t orc_device_rng_draw(uint64_t seed, uint64_t serial, uint32_t j);

/* ---- process-wide statics of the reference (PE.cpp:720, SB.cpp:337-339) ---- */
void orc_reset_statics(void);
void orc_set_rotation_mode(int mode);                     /* T3D_ROT_FROZEN (default) / T3D_ROT_PER_PARTICLE */
void orc_set_frozen_rotation(const float u[3], float angle);
int orc_get_frozen_rotation(float m[16]);                 /* returns 1 if latched */

/* ---- emitter ---- */
orc_emitter* orc_emitter_create(const t3d_emitter_params* p);
void orc_emitter_destroy(orc_emitter* e);
void orc_emitter_set_params(orc_emitter* e, const t3d_emitter_params* p);
void orc_emitter_get_params(orc_emitter* e, t3d_emitter_params* p);
void orc_emitter_set_emission_rate(orc_emitter* e, uint32_t rate);
void orc_emitter_set_sprite_tex_coords(orc_emitter* e, uint32_t frame_count, const float* tex_coords);
void orc_emitter_set_sprite_frame_rects(orc_emitter* e, uint32_t frame_count, const float* rects);
void orc_emitter_set_sprite_frame_coords(orc_emitter* e, uint32_t frame_count, int width, int height);
uint32_t orc_emitter_get_sprite_tex_coords(orc_emitter* e, float* out, size_t max_frames);
void orc_emitter_set_node_world(orc_emitter* e, const float m[16]);
void orc_emitter_set_camera_world(orc_emitter* e, const float m[16]);
/* seed != 0 or enable: draws for this emitter come from orc_device_rng_draw(seed, serial, j) */
void orc_emitter_set_device_rng(orc_emitter* e, int enable, uint64_t seed);
float orc_emitter_get_emit_time(orc_emitter* e);

void orc_emitter_start(orc_emitter* e);
void orc_emitter_stop(orc_emitter* e);
int orc_emitter_is_started(orc_emitter* e);
int orc_emitter_is_active(orc_emitter* e);
void orc_emitter_emit_once(orc_emitter* e, uint32_t n);
void orc_emitter_update(orc_emitter* e, float elapsed_ms);
uint32_t orc_emitter_draw(orc_emitter* e);
uint32_t orc_emitter_get_count(orc_emitter* e);
uint32_t orc_emitter_get_state(orc_emitter* e, uint32_t* out, size_t stride);
int orc_emitter_set_state(orc_emitter* e, const uint32_t* in, size_t stride, uint32_t n);
uint32_t orc_emitter_get_vertices(orc_emitter* e, float* out, size_t max_quads);
/* T3D_VERTEX_CLIP restated: the captured vertices with sprite.vert:19 applied (10 floats per vertex). */
uint32_t orc_emitter_get_clip_vertices(orc_emitter* e, const float vp[16], float* out, size_t max_quads);
/* t3d_emitter_bounds restated: min / max of position -+ size over the live particles; returns 0 when none is live. */
int orc_emitter_bounds(orc_emitter* e, float lo[3], float hi[3]);

/* MeshBatch::add's index stitching for n quads appended one by one (MB.cpp:87-148, TRIANGLE_STRIP, indexed), with
 * 32-bit indices.  Returns the number of indices written (6n-2).  Restated, not reference-pinned (MeshBatch.cpp needs GL). */
uint64_t orc_strip_indices(uint32_t n_quads, uint32_t* out);

/* t3d_emitter_depth_sorted_indices restated (no reference counterpart): 6 indices per live particle, farthest first. */
uint64_t orc_emitter_depth_sorted_indices(orc_emitter* e, const float view_dir[3], uint32_t* out, uint64_t max_indices);
/* one SpriteBatch::draw 2-D call per t3d_sprite2d (SB.cpp:235-289, :366-413, :475-506); 36 floats per quad. */
uint32_t orc_sprites_2d(const t3d_sprite2d* sprites, size_t n, float* out, size_t max_quads);
/* TileSet::draw (TileSet.cpp:207-236) from the position it has computed by line 205. */
uint32_t orc_tileset_draw(const float* sources, uint32_t rows, uint32_t cols, float tile_w, float tile_h, float tex_w, float tex_h,
                          const float position[3], const float color[4], float* out, size_t max_quads);
/* Font::drawText(text, x, y, color, size, rightToLeft), Font.cpp:242-390, one font size; 36 floats per quad. */
uint32_t orc_font_draw_text(const t3d_glyph* glyphs, uint32_t glyph_count, uint32_t font_size, float character_spacing,
                            const char* text, size_t length, int x, int y, const float color[4], uint32_t size, int right_to_left,
                            float* out, size_t max_quads);

/* ---- timing loops (seconds) for bench.py's cpu_baseline "port" leg ---- */
double orc_time_update(orc_emitter* e, float elapsed_ms, int frames);
double orc_time_draw(orc_emitter* e, int frames);
double orc_time_frames(orc_emitter* e, float elapsed_ms, int frames);

#ifdef __cplusplus
}
#endif
#endif
