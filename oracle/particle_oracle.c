/*
 * particle_oracle.c -- TEST INFRASTRUCTURE, not product code (see particle_oracle.h).
 *
 * Plain-C restatement of the reference's particle path.  Every function cites the reference lines it
 * follows:  PE.cpp = Tractor3Dlib/src/graphics/ParticleEmitter.cpp, PE.h = the matching header,
 * SB.cpp = src/graphics/SpriteBatch.cpp, Matrix.cpp = src/math/Matrix.cpp,
 * MathUtil.h = include/math/MathUtil.h, pch.h = include/pch.h.
 *
 * Floating point: compile with -O2 -ffp-contract=off (oracle/Makefile) so mul and add stay separate
 * roundings, as in the reference build and in the --fmad=false kernels.  Expression order is kept
 * exactly as the reference writes it (left-to-right sums, "+ w*m[12]" terms included) so that results
 * are bit-identical to oracle/_ref, signed zeros included.
 */
#define _GNU_SOURCE
#include "particle_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- per-particle record: same field order as ParticleEmitter::Particle, PE.h:792-822 (144 B on LP64) ---- */
typedef struct orc_particle {
    float color_start[4], color_end[4], color[4];
    float position[3], velocity[3], acceleration[3], rotation_axis[3];
    long energy_start, energy;
    float size_start, size_end, size;
    float rotation_per_particle_speed, rotation_speed, angle, time_on_current_frame;
    unsigned int frame;
} orc_particle;

struct orc_emitter {
    t3d_emitter_params p;              /* PE.h:824-877 */
    orc_particle* particles;           /* PE.cpp:31 */
    unsigned int alloc_max;            /* particleCountMax at create: the size of `particles` */
    unsigned int particle_count;
    int started;
    int has_node;
    float node_world[16];
    float camera_world[16];
    float* tex_coords;                 /* frame_count * 4: u1, v1, u2, v2 */
    unsigned int frame_count;
    float percent_per_frame;           /* 1 / frame_count, PE.cpp:518, 532 */
    float frame_duration_secs;         /* PE.cpp:494 */
    float tex_w_ratio, tex_h_ratio;    /* PE.cpp:246-247 */
    float time_per_emission;           /* PE.cpp:266 */
    float emit_time;                   /* PE.h:876 */
    int device_rng;
    uint64_t rng_seed, rng_serial;
    uint32_t rng_j;
    float* vertices;                   /* capture of the last draw: 36 floats per quad */
    size_t vertex_quads, vertex_cap;
};

/* =====================================================================================
 * Draw source: stands in for libc rand() (PE.cpp:359; pch.h:151-152).
 * ===================================================================================== */
static uint64_t g_seq_state = 0x1234ULL;
static int32_t* g_replay = NULL;
static size_t g_replay_n = 0, g_replay_pos = 0;
static int g_replaying = 0, g_recording = 0;
static int32_t* g_recorded = NULL;
static size_t g_recorded_n = 0, g_recorded_cap = 0;
static uint64_t g_rand_calls = 0;

static uint64_t splitmix64_next(uint64_t* s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* Specification of T3D_RNG_DEVICE, restated: a 32-bit hash of (seed, particle serial, draw index).  The seed and the
 * serial are folded into a per-particle key with odd multipliers; draw j is key + (j+1)*0x165667B1 pushed through the
 * "lowbias32" finaliser (xor-shift 16, *0x7FEB352D, xor-shift 15, *0x846CA68B, xor-shift 16); the top 31 bits are the draw. */
int32_t orc_device_rng_draw(uint64_t seed, uint64_t serial, uint32_t j)
{
    uint32_t seed_lo = (uint32_t)(seed & 0xFFFFFFFFu), seed_hi = (uint32_t)(seed >> 32);
    uint32_t ser_lo = (uint32_t)(serial & 0xFFFFFFFFu), ser_hi = (uint32_t)(serial >> 32);
    uint32_t key = (seed_lo * 2654435761u) ^ (seed_hi * 2246822519u) ^ (ser_lo * 3266489917u) ^ (ser_hi * 668265263u);
    uint32_t x = key + (j + 1u) * 374761393u;
    x ^= x >> 16;
    x *= 2146121005u;
    x ^= x >> 15;
    x *= 2221713035u;
    x ^= x >> 16;
    return (int32_t)(x >> 1);
}

static void record(int32_t r)
{
    if (!g_recording) return;
    if (g_recorded_n == g_recorded_cap) {
        g_recorded_cap = g_recorded_cap ? g_recorded_cap * 2 : 4096;
        g_recorded = (int32_t*)realloc(g_recorded, g_recorded_cap * sizeof(int32_t));
    }
    g_recorded[g_recorded_n++] = r;
}

static int32_t next_draw(orc_emitter* e)
{
    int32_t r;
    if (e && e->device_rng) {
        r = orc_device_rng_draw(e->rng_seed, e->rng_serial, e->rng_j++);
    } else if (g_replaying) {
        if (g_replay_pos >= g_replay_n) {
            fprintf(stderr, "particle_oracle: replay stream exhausted after %zu draws\n", g_replay_pos);
            abort();
        }
        r = g_replay[g_replay_pos++];
    } else {
        r = (int32_t)(splitmix64_next(&g_seq_state) >> 33);
    }
    ++g_rand_calls;
    record(r);
    return r;
}

void orc_rand_seed(uint64_t seed) { g_seq_state = seed; g_replaying = 0; }
void orc_rand_replay(const int32_t* draws, size_t n)
{
    free(g_replay); g_replay = NULL; g_replay_n = g_replay_pos = 0; g_replaying = 0;
    if (!draws) return;
    g_replay = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    memcpy(g_replay, draws, n * sizeof(int32_t));
    g_replay_n = n; g_replaying = 1;
}
size_t orc_rand_replay_remaining(void) { return g_replaying ? g_replay_n - g_replay_pos : 0; }
void orc_rand_record(int on) { g_recording = on != 0; g_recorded_n = 0; }
size_t orc_rand_recorded(int32_t* out, size_t max)
{
    size_t n = g_recorded_n < max ? g_recorded_n : max;
    if (out && n) memcpy(out, g_recorded, n * sizeof(int32_t));
    return g_recorded_n;
}
uint64_t orc_rand_calls(void) { return g_rand_calls; }

/* pch.h:151-152.  RAND_MAX is the int 2^31-1 on glibc; `(float)rand()/RAND_MAX` converts it to float,
 * which rounds to 2^31, so the quotient is exactly r * 2^-31 (and can reach 1.0f). */
static float random_0_1(orc_emitter* e) { return (float)next_draw(e) / (float)2147483647; }
static float random_minus1_1(orc_emitter* e) { return (2.0f * ((float)next_draw(e) / (float)2147483647)) - 1.0f; }

/* =====================================================================================
 * Math used by the path.
 * ===================================================================================== */
/* MathUtil.h:195-200 (transformVector4, scalar form): left-to-right sum of four products. */
static void transform_vector4(const float* m, float x, float y, float z, float w, float* dst)
{
    dst[0] = x * m[0] + y * m[4] + z * m[8] + w * m[12];
    dst[1] = x * m[1] + y * m[5] + z * m[9] + w * m[13];
    dst[2] = x * m[2] + y * m[6] + z * m[10] + w * m[14];
}
/* Matrix.cpp:965-968: transformPoint = w 1.  Arguments are taken by value, so dst may alias v. */
static void transform_point(const float* m, const float* v, float* dst) { transform_vector4(m, v[0], v[1], v[2], 1.0f, dst); }
/* Matrix.cpp:971-982: transformVector = w 0 (what `Vector3 *= Matrix` does, Matrix.h:1029-1033). */
static void transform_vector(const float* m, float* v) { transform_vector4(m, v[0], v[1], v[2], 0.0f, v); }

/* Matrix.cpp:348-406 (createRotation from axis + angle).  The reference build resolves cos/sin on a float
 * to the float overloads, which gcc pairs into one sincosf call (checked in oracle/_ref's objects). */
static void create_rotation(const float* axis, float angle, float* m)
{
    float x = axis[0], y = axis[1], z = axis[2];
    float n = x * x + y * y + z * z;
    if (n != 1.0f) {
        n = sqrtf(n);
        if (n > 0.000001f) {
            n = 1.0f / n;
            x *= n; y *= n; z *= n;
        }
    }
    float s, c;
    sincosf(angle, &s, &c);
    float t = 1.0f - c;
    float tx = t * x, ty = t * y, tz = t * z;
    float txy = tx * y, txz = tx * z, tyz = ty * z;
    float sx = s * x, sy = s * y, sz = s * z;
    m[0] = c + tx * x; m[1] = txy + sz;   m[2] = txz - sy;   m[3] = 0.0f;
    m[4] = txy - sz;   m[5] = c + ty * y; m[6] = tyz + sx;   m[7] = 0.0f;
    m[8] = txz + sy;   m[9] = tyz - sx;   m[10] = c + tz * z; m[11] = 0.0f;
    m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
}

/* =====================================================================================
 * Process-wide statics of the reference.
 * ===================================================================================== */
static double g_running_time = 0;       /* PE.cpp:720 */
static int g_rot_latched = 0;           /* SB.cpp:339: function-local static, initialised once */
static float g_rot_matrix[16];
static int g_rot_mode = T3D_ROT_FROZEN;

void orc_reset_statics(void) { g_running_time = 0; g_rot_latched = 0; }
void orc_set_rotation_mode(int mode) { g_rot_mode = mode; }
void orc_set_frozen_rotation(const float u[3], float angle) { create_rotation(u, angle, g_rot_matrix); g_rot_latched = 1; }
int orc_get_frozen_rotation(float m[16]) { if (g_rot_latched && m) memcpy(m, g_rot_matrix, sizeof(g_rot_matrix)); return g_rot_latched; }

/* =====================================================================================
 * Emitter configuration.
 * ===================================================================================== */
static void identity(float* m) { memset(m, 0, 16 * sizeof(float)); m[0] = m[5] = m[10] = m[15] = 1.0f; }

void orc_emitter_set_emission_rate(orc_emitter* e, uint32_t rate)
{
    /* PE.cpp:262-267 */
    e->p.emission_rate = rate;
    e->time_per_emission = 1000.0f / (float)rate;
}

void orc_emitter_set_sprite_tex_coords(orc_emitter* e, uint32_t frame_count, const float* tex_coords)
{
    /* PE.cpp:512-523 */
    e->frame_count = frame_count;
    e->percent_per_frame = 1.0f / (float)frame_count;
    float* nc = (float*)malloc(sizeof(float) * 4 * frame_count);
    memcpy(nc, tex_coords, sizeof(float) * 4 * frame_count);
    free(e->tex_coords);
    e->tex_coords = nc;
}

void orc_emitter_set_sprite_frame_rects(orc_emitter* e, uint32_t frame_count, const float* rects)
{
    /* PE.cpp:526-547: rects are {x, y, width, height} in texels. */
    e->frame_count = frame_count;
    e->percent_per_frame = 1.0f / (float)frame_count;
    free(e->tex_coords);
    e->tex_coords = (float*)malloc(sizeof(float) * 4 * frame_count);
    for (uint32_t i = 0; i < frame_count; ++i) {
        const float* r = rects + 4 * i;
        float* c = e->tex_coords + 4 * i;
        c[0] = e->tex_w_ratio * r[0];
        c[1] = 1.0f - e->tex_h_ratio * r[1];
        c[2] = c[0] + e->tex_w_ratio * r[2];
        c[3] = c[1] - e->tex_h_ratio * r[3];
    }
}

void orc_emitter_set_sprite_frame_coords(orc_emitter* e, uint32_t frame_count, int width, int height)
{
    /* PE.cpp:550-582: walk the sheet left-to-right, top-to-bottom; the rect array has frame_count entries
     * and entries the walk never reaches keep Rectangle's default (all zero). */
    float* rects = (float*)calloc((size_t)frame_count * 4, sizeof(float));
    unsigned int cols = (unsigned int)(e->p.texture_width / (float)width);
    unsigned int rows = (unsigned int)(e->p.texture_height / (float)height);
    unsigned int n = 0;
    for (unsigned int i = 0; i < rows && n != frame_count; ++i) {
        int y = (int)i * height;
        for (unsigned int j = 0; j < cols; ++j) {
            int x = (int)j * width;
            float* r = rects + 4 * ((size_t)i * cols + j);
            r[0] = (float)x; r[1] = (float)y; r[2] = (float)width; r[3] = (float)height;
            if (++n == frame_count) break;
        }
    }
    orc_emitter_set_sprite_frame_rects(e, frame_count, rects);
    free(rects);
}

uint32_t orc_emitter_get_sprite_tex_coords(orc_emitter* e, float* out, size_t max_frames)
{
    size_t n = e->frame_count < max_frames ? e->frame_count : max_frames;
    if (out) memcpy(out, e->tex_coords, sizeof(float) * 4 * n);
    return e->frame_count;
}

void orc_emitter_set_params(orc_emitter* e, const t3d_emitter_params* p)
{
    /* setParticleCountMax (PE.h:237) only rewrites the field; the array keeps the size it got at create (PE.cpp:31), so
     * the field may be lowered -- the clamp of PE.cpp:293-296 then uses the lower value -- and raised back, never beyond */
    uint32_t cap = (p->particle_count_max && p->particle_count_max <= e->alloc_max) ? p->particle_count_max : e->p.particle_count_max;
    float tw = e->p.texture_width, th = e->p.texture_height;
    e->p = *p;
    e->p.particle_count_max = cap;
    e->p.texture_width = tw; e->p.texture_height = th;
    orc_emitter_set_emission_rate(e, p->emission_rate);
    e->frame_duration_secs = (float)p->sprite_frame_duration / 1000.0f; /* PE.cpp:491-495 */
}

void orc_emitter_get_params(orc_emitter* e, t3d_emitter_params* p) { *p = e->p; }

orc_emitter* orc_emitter_create(const t3d_emitter_params* p)
{
    orc_emitter* e = (orc_emitter*)calloc(1, sizeof(orc_emitter));
    e->p = *p;
    e->alloc_max = p->particle_count_max;
    e->particles = (orc_particle*)calloc(p->particle_count_max ? p->particle_count_max : 1, sizeof(orc_particle));
    identity(e->node_world);
    identity(e->camera_world);
    e->has_node = 1;
    /* setTexture, PE.cpp:244-251: cache the texture size and its reciprocals, install one full-texture frame. */
    e->tex_w_ratio = 1.0f / (float)(unsigned int)p->texture_width;
    e->tex_h_ratio = 1.0f / (float)(unsigned int)p->texture_height;
    float full[4] = { 0.0f, 0.0f, p->texture_width, p->texture_height };
    orc_emitter_set_sprite_frame_rects(e, 1, full);
    orc_emitter_set_params(e, p);
    return e;
}

void orc_emitter_destroy(orc_emitter* e)
{
    if (!e) return;
    free(e->particles); free(e->tex_coords); free(e->vertices); free(e);
}

void orc_emitter_set_node_world(orc_emitter* e, const float m[16]) { memcpy(e->node_world, m, 64); e->has_node = 1; }
void orc_emitter_set_camera_world(orc_emitter* e, const float m[16]) { memcpy(e->camera_world, m, 64); }
void orc_emitter_set_device_rng(orc_emitter* e, int enable, uint64_t seed) { e->device_rng = enable; e->rng_seed = seed; e->rng_serial = 0; }
float orc_emitter_get_emit_time(orc_emitter* e) { return e->emit_time; }

void orc_emitter_start(orc_emitter* e) { e->started = 1; }   /* PE.cpp:270-274 */
void orc_emitter_stop(orc_emitter* e) { e->started = 0; }    /* PE.h:272 */
int orc_emitter_is_started(orc_emitter* e) { return e->started; }
int orc_emitter_is_active(orc_emitter* e)
{
    /* PE.cpp:277-284 */
    if (e->started) return 1;
    if (!e->has_node) return 0;
    return e->particle_count > 0;
}
uint32_t orc_emitter_get_count(orc_emitter* e) { return e->particle_count; }

/* =====================================================================================
 * Spawn.
 * ===================================================================================== */
static float generate_scalar(orc_emitter* e, float min, float max) { return min + (max - min) * random_0_1(e); } /* PE.cpp:611-614 */

static void generate_color(orc_emitter* e, const float* base, const float* var, float* dst)
{
    /* PE.cpp:669-679: x, y, z, w in that order, one draw each (consumed even when the variance is 0). */
    dst[0] = base[0] + var[0] * random_minus1_1(e);
    dst[1] = base[1] + var[1] * random_minus1_1(e);
    dst[2] = base[2] + var[2] * random_minus1_1(e);
    dst[3] = base[3] + var[3] * random_minus1_1(e);
}

static void generate_vector(orc_emitter* e, const float* base, const float* var, float* dst, int ellipsoid)
{
    if (ellipsoid) {
        /* PE.cpp:629-650: rejection-sample the unit ball, then scale and translate. */
        do {
            dst[0] = random_minus1_1(e);
            dst[1] = random_minus1_1(e);
            dst[2] = random_minus1_1(e);
        } while (sqrtf(dst[0] * dst[0] + dst[1] * dst[1] + dst[2] * dst[2]) > 1.0f); /* Vector3::length, Vector3.h:259 */
        dst[0] *= var[0]; dst[1] *= var[1]; dst[2] *= var[2];
        dst[0] += base[0]; dst[1] += base[1]; dst[2] += base[2];
    } else {
        /* PE.cpp:617-626 */
        dst[0] = base[0] + var[0] * random_minus1_1(e);
        dst[1] = base[1] + var[1] * random_minus1_1(e);
        dst[2] = base[2] + var[2] * random_minus1_1(e);
    }
}

void orc_emitter_emit_once(orc_emitter* e, uint32_t n)
{
    /* PE.cpp:287-369 */
    const t3d_emitter_params* P = &e->p;
    if (n + e->particle_count > P->particle_count_max) n = P->particle_count_max - e->particle_count; /* :293-296 */

    float world[16];
    memcpy(world, e->node_world, sizeof(world));
    float translation[3] = { world[12], world[13], world[14] }; /* Matrix::getTranslation -> decompose, Matrix.cpp:517-523 */
    world[12] = 0.0f; world[13] = 0.0f; world[14] = 0.0f;       /* :303-305 */

    for (uint32_t i = 0; i < n; ++i) {
        orc_particle* p = &e->particles[e->particle_count];
        e->rng_j = 0;

        generate_color(e, P->color_start, P->color_start_var, p->color_start);
        generate_color(e, P->color_end, P->color_end_var, p->color_end);
        memcpy(p->color, p->color_start, sizeof(p->color));

        /* :316 -- the bounds are float members (PE.h:840-841), so the float overload is chosen and the
         * result is truncated into the long fields. */
        p->energy = p->energy_start = (long)generate_scalar(e, P->energy_min, P->energy_max);
        p->size = p->size_start = generate_scalar(e, P->size_start_min, P->size_start_max);
        p->size_end = generate_scalar(e, P->size_end_min, P->size_end_max);
        p->rotation_per_particle_speed = generate_scalar(e, P->rotation_per_particle_speed_min, P->rotation_per_particle_speed_max);
        p->angle = generate_scalar(e, 0.0f, p->rotation_per_particle_speed);
        p->rotation_speed = generate_scalar(e, P->rotation_speed_min, P->rotation_speed_max);

        generate_vector(e, P->position, P->position_var, p->position, P->ellipsoid); /* :325 */
        generate_vector(e, P->velocity, P->velocity_var, p->velocity, 0);
        generate_vector(e, P->acceleration, P->acceleration_var, p->acceleration, 0);
        generate_vector(e, P->rotation_axis, P->rotation_axis_var, p->rotation_axis, 0);

        if (P->orbit_position) transform_point(world, p->position, p->position);             /* :332-335 */
        if (P->orbit_velocity) transform_point(world, p->velocity, p->velocity);             /* :337-340 */
        if (P->orbit_acceleration) transform_point(world, p->acceleration, p->acceleration); /* :342-345 */
        if (p->rotation_speed != 0.0f &&
            !(p->rotation_axis[0] == 0.0f && p->rotation_axis[1] == 0.0f && p->rotation_axis[2] == 0.0f))
            transform_point(world, p->rotation_axis, p->rotation_axis);                     /* :348-351 */

        p->position[0] += translation[0]; p->position[1] += translation[1]; p->position[2] += translation[2]; /* :354 */

        if (P->sprite_frame_random_offset > 0)
            p->frame = (unsigned int)next_draw(e) % P->sprite_frame_random_offset;          /* :357-360 */
        else
            p->frame = 0;
        p->time_on_current_frame = 0.0f;

        ++e->particle_count;
        ++e->rng_serial;
    }
}

/* =====================================================================================
 * Update.
 * ===================================================================================== */
void orc_emitter_update(orc_emitter* e, float elapsed_time)
{
    /* PE.cpp:713-832 */
    if (!orc_emitter_is_active(e)) return;

    g_running_time += elapsed_time;              /* :720-725, shared by every emitter of the process */
    if (g_running_time < 8) return;
    float elapsed_ms = (float)g_running_time;
    g_running_time = 0;

    float elapsed_secs = elapsed_ms * 0.001f;

    if (e->started && e->p.emission_rate) {
        e->emit_time += elapsed_ms;                                             /* :732 */
        unsigned int emit_count = (unsigned int)(e->emit_time / e->time_per_emission);
        if (emit_count) {
            if ((int)e->time_per_emission > 0)                                 /* :740-743 */
                e->emit_time = (float)fmod((double)e->emit_time, (double)e->time_per_emission);
            orc_emitter_emit_once(e, emit_count);
        }
    }

    const int animated = e->p.sprite_animated, looped = e->p.sprite_looped;
    for (size_t i = 0; i < e->particle_count; ++i) {
        orc_particle* p = &e->particles[i];
        p->energy = (long)((float)p->energy - elapsed_ms);                      /* :753  long -= float */

        if (p->energy > 0L) {
            if (p->rotation_speed != 0.0f &&
                !(p->rotation_axis[0] == 0.0f && p->rotation_axis[1] == 0.0f && p->rotation_axis[2] == 0.0f)) {
                float rot[16];                                                  /* :757-763 */
                create_rotation(p->rotation_axis, p->rotation_speed * elapsed_secs, rot);
                transform_point(rot, p->velocity, p->velocity);
                transform_point(rot, p->acceleration, p->acceleration);
            }
            p->velocity[0] += p->acceleration[0] * elapsed_secs;                /* :766-772 */
            p->velocity[1] += p->acceleration[1] * elapsed_secs;
            p->velocity[2] += p->acceleration[2] * elapsed_secs;
            p->position[0] += p->velocity[0] * elapsed_secs;
            p->position[1] += p->velocity[1] * elapsed_secs;
            p->position[2] += p->velocity[2] * elapsed_secs;
            p->angle += p->rotation_per_particle_speed * elapsed_secs;          /* :774 */

            float percent = 1.0f - ((float)p->energy / (float)p->energy_start); /* :777 */
            for (int k = 0; k < 4; ++k)
                p->color[k] = p->color_start[k] + (p->color_end[k] - p->color_start[k]) * percent; /* :779-782 */
            p->size = p->size_start + (p->size_end - p->size_start) * percent;  /* :784 */

            if (animated) {
                if (!looped) {                                                  /* :789-803 */
                    float percent_spent = 0.0f;
                    for (size_t k = 0; k < p->frame; ++k) percent_spent += e->percent_per_frame;
                    p->time_on_current_frame = percent - percent_spent;
                    if (p->frame < e->frame_count - 1 && p->time_on_current_frame >= e->percent_per_frame) ++p->frame;
                } else {                                                        /* :804-818 */
                    p->time_on_current_frame += elapsed_secs;
                    if (p->time_on_current_frame >= e->frame_duration_secs) {
                        p->time_on_current_frame -= e->frame_duration_secs;
                        ++p->frame;
                        if (p->frame == e->frame_count) p->frame = 0;
                    }
                }
            }
        } else {
            /* :821-830 swap-with-last removal; the moved particle is not examined this frame. */
            if (i != e->particle_count - 1) e->particles[i] = e->particles[e->particle_count - 1];
            --e->particle_count;
        }
    }
}

/* =====================================================================================
 * Draw.
 * ===================================================================================== */
static void billboard(const float* pos, const float* right, const float* up, float width, float height,
                      float u1, float v1, float u2, float v2, const float* color, float angle, float* out)
{
    /* SB.cpp:292-363, with rotationPoint = (0.5, 0.5) as ParticleEmitter::draw passes it (PE.cpp:855). */
    float tr[3] = { right[0] * (width * 0.5f), right[1] * (width * 0.5f), right[2] * (width * 0.5f) };
    float tf[3] = { up[0] * (height * 0.5f), up[1] * (height * 0.5f), up[2] * (height * 0.5f) };
    float p0[3], p1[3], p2[3], p3[3];
    for (int k = 0; k < 3; ++k) { p0[k] = pos[k]; p0[k] -= tr[k]; p0[k] -= tf[k]; }
    for (int k = 0; k < 3; ++k) { p1[k] = pos[k]; p1[k] += tr[k]; p1[k] -= tf[k]; }
    for (int k = 0; k < 3; ++k) tf[k] = up[k] * height;
    for (int k = 0; k < 3; ++k) { p2[k] = p0[k] + tf[k]; p3[k] = p1[k] + tf[k]; }

    if (angle != 0) {
        float rp[3];
        for (int k = 0; k < 3; ++k) {
            tr[k] = right[k] * (width * 0.5f);
            tf[k] *= 0.5f;
            rp[k] = p0[k];
            rp[k] += tr[k];
            rp[k] += tf[k];
        }
        float u[3] = { (right[1] * up[2]) - (right[2] * up[1]),   /* MathUtil.h:216-225 */
                       (right[2] * up[0]) - (right[0] * up[2]),
                       (right[0] * up[1]) - (right[1] * up[0]) };
        float local[16];
        const float* rot;
        if (g_rot_mode == T3D_ROT_FROZEN) {
            /* SB.cpp:339: `static Matrix rotation = createRotation(u, rotationAngle)` -- computed by the first
             * rotated quad the process ever draws and reused by every later one. */
            if (!g_rot_latched) { create_rotation(u, angle, g_rot_matrix); g_rot_latched = 1; }
            rot = g_rot_matrix;
        } else {
            create_rotation(u, angle, local);
            rot = local;
        }
        float* ps[4] = { p0, p1, p2, p3 };
        for (int q = 0; q < 4; ++q) {
            float* p = ps[q];
            p[0] -= rp[0]; p[1] -= rp[1]; p[2] -= rp[2];
            transform_vector(rot, p);
            p[0] += rp[0]; p[1] += rp[1]; p[2] += rp[2];
        }
    }
    /* SB.cpp:355-359 */
    const float* ps[4] = { p0, p1, p2, p3 };
    const float us[4] = { u1, u2, u1, u2 };
    const float vs[4] = { v1, v1, v2, v2 };
    for (int q = 0; q < 4; ++q) {
        float* v = out + 9 * q;
        v[0] = ps[q][0]; v[1] = ps[q][1]; v[2] = ps[q][2];
        v[3] = us[q]; v[4] = vs[q];
        v[5] = color[0]; v[6] = color[1]; v[7] = color[2]; v[8] = color[3];
    }
}

uint32_t orc_emitter_draw(orc_emitter* e)
{
    /* PE.cpp:835-888 */
    e->vertex_quads = 0; /* nothing is submitted unless the batch is started and finished below */
    if (!orc_emitter_is_active(e)) return 0;
    if (e->particle_count > 0) {
        if (e->vertex_cap < e->particle_count) {
            e->vertex_cap = e->particle_count;
            e->vertices = (float*)realloc(e->vertices, e->vertex_cap * T3D_FLOATS_PER_QUAD * sizeof(float));
        }
        e->vertex_quads = 0; /* SpriteBatch::start() */
        const float* cw = e->camera_world;
        float right[3] = { cw[0], cw[1], cw[2] }; /* Matrix.cpp:692-701 */
        float up[3] = { cw[4], cw[5], cw[6] };    /* Matrix.cpp:656-665 */
        for (size_t i = 0; i < e->particle_count; ++i) {
            const orc_particle* p = &e->particles[i];
            const float* uv = e->tex_coords + (size_t)p->frame * 4;
            billboard(p->position, right, up, p->size, p->size, uv[0], uv[1], uv[2], uv[3], p->color, p->angle,
                      e->vertices + e->vertex_quads * T3D_FLOATS_PER_QUAD);
            ++e->vertex_quads;
        }
    }
    return 1;
}

uint32_t orc_emitter_get_vertices(orc_emitter* e, float* out, size_t max_quads)
{
    size_t n = e->vertex_quads < max_quads ? e->vertex_quads : max_quads;
    if (out && n) memcpy(out, e->vertices, n * T3D_FLOATS_PER_QUAD * sizeof(float));
    return (uint32_t)e->vertex_quads;
}

/* The vertex shader's one line applied on the CPU (res/shaders/sprite.vert:19: gl_Position = u_projectionMatrix *
 * vec4(a_position, 1), with the matrix ParticleEmitter::draw hands the batch at PE.cpp:846-849): every captured vertex
 * becomes {x, y, z, w, u, v, r, g, b, a}.  RESTATED, not reference-pinned (the reference leaves this to the GPU); each
 * component is summed in MathUtil::transformVector4's order (MathUtil.h:195-200). */
uint32_t orc_emitter_get_clip_vertices(orc_emitter* e, const float vp[16], float* out, size_t max_quads)
{
    size_t n = e->vertex_quads < max_quads ? e->vertex_quads : max_quads;
    for (size_t q = 0; out && q < n; ++q)
        for (int k = 0; k < 4; ++k) {
            const float* v = e->vertices + (q * 4 + (size_t)k) * 9;
            float* o = out + (q * 4 + (size_t)k) * 10;
            transform_vector4(vp, v[0], v[1], v[2], 1.0f, o); /* writes o[0..2] */
            o[3] = v[0] * vp[3] + v[1] * vp[7] + v[2] * vp[11] + 1.0f * vp[15];
            memcpy(o + 4, v + 3, 6 * sizeof(float));
        }
    return (uint32_t)e->vertex_quads;
}

/* Bounding box of the live particles: min / max over position -+ size per axis (the engine has a TODO here,
 * src/scene/Node.cpp:777-780).  RESTATED definition of t3d_emitter_bounds.  Returns 0 when no particle is live. */
int orc_emitter_bounds(orc_emitter* e, float lo[3], float hi[3])
{
    if (!e->particle_count) return 0;
    for (int k = 0; k < 3; ++k) { lo[k] = 3.402823466e38f; hi[k] = -3.402823466e38f; }
    for (size_t i = 0; i < e->particle_count; ++i) {
        const orc_particle* p = &e->particles[i];
        for (int k = 0; k < 3; ++k) {
            const float a = p->position[k] - p->size, b = p->position[k] + p->size;
            if (a < lo[k]) lo[k] = a;
            if (b > hi[k]) hi[k] = b;
        }
    }
    return 1;
}

/* =====================================================================================
 * State exchange in the column layout of include/t3d_particles.h.
 * ===================================================================================== */
static void put_f(uint32_t* out, size_t stride, int col, size_t i, float v) { memcpy(&out[(size_t)col * stride + i], &v, 4); }
static float get_f(const uint32_t* in, size_t stride, int col, size_t i) { float v; memcpy(&v, &in[(size_t)col * stride + i], 4); return v; }

uint32_t orc_emitter_get_state(orc_emitter* e, uint32_t* out, size_t stride)
{
    uint32_t n = e->particle_count;
    if (!out) return n;
    for (uint32_t i = 0; i < n; ++i) {
        const orc_particle* p = &e->particles[i];
        for (int k = 0; k < 4; ++k) {
            put_f(out, stride, T3D_COL_COLOR_START + k, i, p->color_start[k]);
            put_f(out, stride, T3D_COL_COLOR_END + k, i, p->color_end[k]);
            put_f(out, stride, T3D_COL_COLOR + k, i, p->color[k]);
        }
        for (int k = 0; k < 3; ++k) {
            put_f(out, stride, T3D_COL_POSITION + k, i, p->position[k]);
            put_f(out, stride, T3D_COL_VELOCITY + k, i, p->velocity[k]);
            put_f(out, stride, T3D_COL_ACCELERATION + k, i, p->acceleration[k]);
            put_f(out, stride, T3D_COL_ROTATION_AXIS + k, i, p->rotation_axis[k]);
        }
        out[(size_t)T3D_COL_ENERGY_START * stride + i] = (uint32_t)(int32_t)p->energy_start;
        out[(size_t)T3D_COL_ENERGY * stride + i] = (uint32_t)(int32_t)p->energy;
        put_f(out, stride, T3D_COL_SIZE_START, i, p->size_start);
        put_f(out, stride, T3D_COL_SIZE_END, i, p->size_end);
        put_f(out, stride, T3D_COL_SIZE, i, p->size);
        put_f(out, stride, T3D_COL_ROT_PER_PARTICLE_SPEED, i, p->rotation_per_particle_speed);
        put_f(out, stride, T3D_COL_ROTATION_SPEED, i, p->rotation_speed);
        put_f(out, stride, T3D_COL_ANGLE, i, p->angle);
        put_f(out, stride, T3D_COL_TIME_ON_FRAME, i, p->time_on_current_frame);
        out[(size_t)T3D_COL_FRAME * stride + i] = p->frame;
    }
    return n;
}

int orc_emitter_set_state(orc_emitter* e, const uint32_t* in, size_t stride, uint32_t n)
{
    if (n > e->p.particle_count_max) return 1;
    for (uint32_t i = 0; i < n; ++i) {
        orc_particle* p = &e->particles[i];
        for (int k = 0; k < 4; ++k) {
            p->color_start[k] = get_f(in, stride, T3D_COL_COLOR_START + k, i);
            p->color_end[k] = get_f(in, stride, T3D_COL_COLOR_END + k, i);
            p->color[k] = get_f(in, stride, T3D_COL_COLOR + k, i);
        }
        for (int k = 0; k < 3; ++k) {
            p->position[k] = get_f(in, stride, T3D_COL_POSITION + k, i);
            p->velocity[k] = get_f(in, stride, T3D_COL_VELOCITY + k, i);
            p->acceleration[k] = get_f(in, stride, T3D_COL_ACCELERATION + k, i);
            p->rotation_axis[k] = get_f(in, stride, T3D_COL_ROTATION_AXIS + k, i);
        }
        p->energy_start = (int32_t)in[(size_t)T3D_COL_ENERGY_START * stride + i];
        p->energy = (int32_t)in[(size_t)T3D_COL_ENERGY * stride + i];
        p->size_start = get_f(in, stride, T3D_COL_SIZE_START, i);
        p->size_end = get_f(in, stride, T3D_COL_SIZE_END, i);
        p->size = get_f(in, stride, T3D_COL_SIZE, i);
        p->rotation_per_particle_speed = get_f(in, stride, T3D_COL_ROT_PER_PARTICLE_SPEED, i);
        p->rotation_speed = get_f(in, stride, T3D_COL_ROTATION_SPEED, i);
        p->angle = get_f(in, stride, T3D_COL_ANGLE, i);
        p->time_on_current_frame = get_f(in, stride, T3D_COL_TIME_ON_FRAME, i);
        p->frame = in[(size_t)T3D_COL_FRAME * stride + i];
    }
    e->particle_count = n;
    return 0;
}

/* =====================================================================================
 * MeshBatch index stitching (MB.cpp:87-148): every SpriteBatch::draw adds 4 vertices with indices {0,1,2,3}.
 * ===================================================================================== */
uint64_t orc_strip_indices(uint32_t n_quads, uint32_t* out)
{
    static const uint32_t quad[4] = { 0, 1, 2, 3 }; /* SB.cpp:361 */
    uint64_t index_count = 0;
    uint32_t vertex_count = 0;
    for (uint32_t q = 0; q < n_quads; ++q) {
        if (vertex_count == 0) {
            for (int i = 0; i < 4; ++i) out[index_count + i] = quad[i];        /* MB.cpp:118-122 */
        } else {
            out[index_count] = out[index_count - 1];                           /* MB.cpp:126-133: degenerate triangle */
            out[index_count + 1] = vertex_count;
            index_count += 2;
            for (int i = 0; i < 4; ++i) out[index_count + i] = quad[i] + vertex_count; /* MB.cpp:137-141 */
        }
        index_count += 4;
        vertex_count += 4;
    }
    return index_count;
}

/* =====================================================================================
 * SURVEY 8(f4), restated (the reference has no depth sort): triangle-list indices that walk the live particles by
 * decreasing position . view_dir, equal depths in live-array order.  The order is DEFINED on the integer key
 * ~order(depth) (order() maps a float onto an unsigned integer monotonically, -0 below +0), sorted ascending and stably.
 * ===================================================================================== */
static uint32_t float_order_u(float f)
{
    uint32_t b;
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
typedef struct { uint32_t key, index; } depth_key;
static int depth_key_cmp(const void* a, const void* b)
{
    const depth_key* x = (const depth_key*)a; const depth_key* y = (const depth_key*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->index < y->index ? -1 : (x->index > y->index ? 1 : 0);   /* stable */
}
uint64_t orc_emitter_depth_sorted_indices(orc_emitter* e, const float view_dir[3], uint32_t* out, uint64_t max_indices)
{
    const uint32_t n = e->particle_count;
    if (!out || !n) return 6ull * n;
    depth_key* keys = (depth_key*)malloc(sizeof(depth_key) * n);
    for (uint32_t i = 0; i < n; ++i) {
        const orc_particle* p = &e->particles[i];
        const float d = p->position[0] * view_dir[0] + p->position[1] * view_dir[1] + p->position[2] * view_dir[2];
        keys[i].key = ~float_order_u(d);
        keys[i].index = i;
    }
    qsort(keys, n, sizeof(depth_key), depth_key_cmp);
    static const uint32_t pat[6] = { 0, 1, 2, 2, 1, 3 };
    for (uint64_t j = 0; j < 6ull * n && j < max_indices; ++j) out[j] = 4u * keys[j / 6].index + pat[j % 6];
    free(keys);
    return 6ull * n;
}

/* =====================================================================================
 * SpriteBatch's 2-D overloads (SB.cpp:235-289, :366-413, :475-506, clipSprite :523-597) and TileSet::draw
 * (src/graphics/TileSet.cpp:179-236).  Pinned bit for bit to the reference's own code by tests/test_oracle_vs_ref.py.
 * ===================================================================================== */
static void sprite_vertex(float* v, float x, float y, float z, float u, float t, const float* c)
{
    v[0] = x; v[1] = y; v[2] = z; v[3] = u; v[4] = t; v[5] = c[0]; v[6] = c[1]; v[7] = c[2]; v[8] = c[3]; /* SB.cpp:21-31 */
}

/* Vector2::rotate, src/math/Vector2.cpp:181-200.  `double sinAngle = sin(angle)` with a float argument: in C++ that is
 * the float overload (sinf) widened afterwards; the products are then formed in double and rounded to float where the
 * reference assigns to its float members. */
static void rotate2(float* x, float* y, float px, float py, float angle)
{
    const double s = (double)sinf(angle), c = (double)cosf(angle);
    if (px == 0.0f && py == 0.0f) {
        const float tx = (float)(*x * c - *y * s);
        *y = (float)(*y * c + *x * s);
        *x = tx;
    } else {
        const float tx = *x - px, ty = *y - py;
        *x = (float)(tx * c - ty * s + px);
        *y = (float)(ty * c + tx * s + py);
    }
}

/* SB.cpp:523-597 */
static int clip_sprite(const float* clip, float* x, float* y, float* w, float* h, float* u1, float* v1, float* u2, float* v2)
{
    if (*x + *w < clip[0] || *x > clip[0] + clip[2] || *y + *h < clip[1] || *y > clip[1] + clip[3]) return 0;
    float uvw = *u2 - *u1, uvh = *v2 - *v1;
    if (*x < clip[0]) {
        const float percent = (clip[0] - *x) / *w, dx = clip[0] - *x, du = uvw * percent;
        *x = clip[0]; *w -= dx; *u1 += du; uvw -= du;
    }
    if (*y < clip[1]) {
        const float percent = (clip[1] - *y) / *h, dy = clip[1] - *y, dv = uvh * percent;
        *y = clip[1]; *h -= dy; *v1 += dv; uvh -= dv;
    }
    const float cx2 = clip[0] + clip[2];
    const float x2 = *x + *w;
    if (x2 > cx2) {
        const float percent = (x2 - cx2) / *w;
        *w = cx2 - *x;
        *u2 -= uvw * percent;
    }
    const float cy2 = clip[1] + clip[3];
    const float y2 = *y + *h;
    if (y2 > cy2) {
        const float percent = (y2 - cy2) / *h;
        *h = cy2 - *y;
        *v2 -= uvh * percent;
    }
    return 1;
}

/* One sprite -> 0 or 1 quad (36 floats at q). */
static int sprite_quad(const t3d_sprite2d* sp, float* q)
{
    float x = sp->x, y = sp->y, w = sp->width, h = sp->height, u1 = sp->u1, v1 = sp->v1, u2 = sp->u2, v2 = sp->v2;
    const float z = sp->z;
    if (sp->flags & T3D_SPRITE_CLIPPED) {                       /* SB.cpp:394-413 */
        if (!clip_sprite(sp->clip, &x, &y, &w, &h, &u1, &v1, &u2, &v2)) return 0;
    } else if (sp->flags & T3D_SPRITE_CENTERED) {               /* SB.cpp:248-252, :487-491 */
        x -= 0.5f * w;
        y -= 0.5f * h;
    }
    const float x2 = x + w, y2 = y + h;
    if ((sp->flags & T3D_SPRITE_ROTATED) && !(sp->flags & T3D_SPRITE_CLIPPED)) {  /* SB.cpp:254-288 */
        float ulx = x, uly = y, urx = x2, ury = y, dlx = x, dly = y2, drx = x2, dry = y2;
        if (sp->rotation_angle != 0) {
            float px = sp->rotation_point[0], py = sp->rotation_point[1];
            px *= w; py *= h; px += x; py += y;
            rotate2(&ulx, &uly, px, py, sp->rotation_angle);
            rotate2(&urx, &ury, px, py, sp->rotation_angle);
            rotate2(&dlx, &dly, px, py, sp->rotation_angle);
            rotate2(&drx, &dry, px, py, sp->rotation_angle);
        }
        sprite_vertex(q + 0, dlx, dly, z, u1, v1, sp->color);
        sprite_vertex(q + 9, ulx, uly, z, u1, v2, sp->color);
        sprite_vertex(q + 18, drx, dry, z, u2, v1, sp->color);
        sprite_vertex(q + 27, urx, ury, z, u2, v2, sp->color);
    } else {                                                    /* SB.cpp:493-505 */
        sprite_vertex(q + 0, x, y, z, u1, v1, sp->color);
        sprite_vertex(q + 9, x, y2, z, u1, v2, sp->color);
        sprite_vertex(q + 18, x2, y, z, u2, v1, sp->color);
        sprite_vertex(q + 27, x2, y2, z, u2, v2, sp->color);
    }
    return 1;
}

uint32_t orc_sprites_2d(const t3d_sprite2d* sp, size_t n, float* out, size_t max_quads)
{
    uint32_t quads = 0;
    float scratch[36];
    for (size_t i = 0; i < n; ++i)
        quads += (uint32_t)sprite_quad(&sp[i], (out && quads < max_quads) ? out + (size_t)quads * 36 : scratch);
    return quads;
}

/* TileSet::draw, TileSet.cpp:207-236, from the position the function has arrived at by line 205 (node translation minus
 * the camera's x and y).  sources = rows * cols * {x, y}, negative = blank. */
uint32_t orc_tileset_draw(const float* sources, uint32_t rows, uint32_t cols, float tile_w, float tile_h, float tex_w, float tex_h,
                          const float position[3], const float color[4], float* out, size_t max_quads)
{
    const float wr = 1.0f / tex_w, hr = 1.0f / tex_h;          /* SB.cpp:155-156 */
    float px = position[0], py = position[1];
    py += tile_h * (float)(rows - 1);                            /* :208 */
    const float x_start = px;
    uint32_t quads = 0;
    float scratch[36];
    for (uint32_t row = 0; row < rows; ++row) {
        for (uint32_t col = 0; col < cols; ++col) {
            const float sx = sources[2 * ((size_t)row * cols + col)], sy = sources[2 * ((size_t)row * cols + col) + 1];
            if (sx >= 0 && sy >= 0) {                            /* :217 */
                t3d_sprite2d sp;
                memset(&sp, 0, sizeof(sp));
                sp.x = px; sp.y = py; sp.z = position[2]; sp.width = tile_w; sp.height = tile_h;
                sp.u1 = wr * sx;                                 /* SB.cpp:208-211 */
                sp.v1 = 1.0f - hr * sy;
                sp.u2 = sp.u1 + wr * tile_w;
                sp.v2 = sp.v1 - hr * tile_h;
                memcpy(sp.color, color, 16);
                sp.rotation_point[0] = sp.rotation_point[1] = 0.5f;
                sp.flags = T3D_SPRITE_ROTATED;                   /* :223-228: the rotationPoint overload, angle 0 */
                quads += (uint32_t)sprite_quad(&sp, (out && quads < max_quads) ? out + (size_t)quads * 36 : scratch);
            }
            px += tile_w;                                        /* :231 */
        }
        px = x_start;                                            /* :233-234 */
        py -= tile_h;
    }
    return quads;
}

/* Font::drawText(text, x, y, color, size, rightToLeft), src/renderer/Font.cpp:242-390, for ONE size of a font (size
 * delegation, :258-264, is the caller's): the pen walk, one SpriteBatch::draw(x, y, w, h, u1, v1, u2, v2, color) per
 * character with a glyph (SB.cpp:366-378 -> :475-506, z = 0, position not centred).  36 floats per quad. */
typedef struct {
    const t3d_glyph* glyphs; uint32_t glyph_count; float scale; int spacing; uint32_t size; int x0;
    int x_pos, y_pos; const float* color; float* out; size_t max_quads; uint32_t quads;
} font_walk;
static void font_walk_char(font_walk* w, signed char c)                                /* :340: `char`, signed on the reference's platforms */
{
    float scratch[36];
    switch (c) {                                                                       /* :342-384 */
    case ' ':  w->x_pos = (int)((unsigned)w->x_pos + w->glyphs[0].advance); break;
    case '\r':
    case '\n': w->y_pos = (int)((unsigned)w->y_pos + w->size); w->x_pos = w->x0; break;
    case '\t': w->x_pos = (int)((unsigned)w->x_pos + w->glyphs[0].advance * 4u); break;
    default: {
        const int index = c - 32;                                                      /* :354 */
        if (index >= 0 && index < (int)w->glyph_count) {
            const t3d_glyph* g = &w->glyphs[index];
            t3d_sprite2d sp;
            memset(&sp, 0, sizeof(sp));
            sp.x = (float)(w->x_pos + (int)((float)g->bearing_x * w->scale));          /* :369 */
            sp.y = (float)w->y_pos;
            sp.width = (float)g->width * w->scale;                                     /* :371 */
            sp.height = (float)w->size;
            sp.u1 = g->uvs[0]; sp.v1 = g->uvs[1]; sp.u2 = g->uvs[2]; sp.v2 = g->uvs[3];
            memcpy(sp.color, w->color, 16);
            w->quads += (uint32_t)sprite_quad(&sp, (w->out && w->quads < w->max_quads) ? w->out + (size_t)w->quads * 36 : scratch);
            /* :378: `xPos += floor(...)` -- the floor is taken in float, the sum in double (exact), back to int */
            w->x_pos = (int)((double)w->x_pos + (double)floorf((float)g->advance * w->scale + (float)w->spacing));
        }
        break;
    }
    }
}
uint32_t orc_font_draw_text(const t3d_glyph* glyphs, uint32_t glyph_count, uint32_t font_size, float character_spacing,
                            const char* text, size_t length, int x, int y, const float color[4], uint32_t size, int right_to_left,
                            float* out, size_t max_quads)
{
    if (size == 0) size = font_size;                                  /* :253-256 */
    font_walk w;
    w.glyphs = glyphs; w.glyph_count = glyph_count; w.size = size; w.x0 = x;
    w.scale = (float)size / (float)font_size;                         /* :268 */
    w.spacing = (int)((float)size * character_spacing);               /* :269 */
    w.x_pos = x; w.y_pos = y;                                         /* :277 */
    w.color = color; w.out = out; w.max_quads = max_quads; w.quads = 0;
    if (!right_to_left) {
        for (size_t i = 0; i < length; ++i) font_walk_char(&w, (signed char)text[i]);   /* :330 (text.length(): embedded NULs are walked over) */
        return w.quads;
    }
    /* :283-327, :386-389: per line, the delimiters it starts with going forwards, then the rest of it from its last character
     * backwards; the walk is over text.c_str(), so it ends at the first NUL */
    size_t end = 0;
    while (end < length && text[end] != 0) ++end;
    size_t cursor = 0;
    for (;;) {
        int done = 0;
        for (;;) {                                                    /* :287-315 */
            if (cursor >= end) { done = 1; break; }
            const char d = text[cursor];
            if (!(d == ' ' || d == '\t' || d == '\r' || d == '\n')) break;
            font_walk_char(&w, (signed char)d);                       /* (the same three cases of the switch) */
            ++cursor;
        }
        if (done) break;
        size_t len = 0;                                               /* :317 strcspn(cursor, "\r\n") */
        while (cursor + len < end && text[cursor + len] != '\r' && text[cursor + len] != '\n') ++len;
        for (size_t i = len; i-- > 0;) font_walk_char(&w, (signed char)text[cursor + i]); /* :318-319, :330 */
        cursor += len;                                                /* :387 */
    }
    return w.quads;
}

/* =====================================================================================
 * Timing loops.
 * ===================================================================================== */
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }
double orc_time_update(orc_emitter* e, float elapsed_ms, int frames)
{
    double t0 = now_s();
    for (int i = 0; i < frames; ++i) orc_emitter_update(e, elapsed_ms);
    return now_s() - t0;
}
double orc_time_draw(orc_emitter* e, int frames)
{
    double t0 = now_s();
    for (int i = 0; i < frames; ++i) orc_emitter_draw(e);
    return now_s() - t0;
}
double orc_time_frames(orc_emitter* e, float elapsed_ms, int frames)
{
    double t0 = now_s();
    for (int i = 0; i < frames; ++i) { orc_emitter_update(e, elapsed_ms); orc_emitter_draw(e); }
    return now_s() - t0;
}
