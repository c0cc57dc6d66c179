// ref_capi.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A C-callable harness around the UNMODIFIED reference implementation of the particle path.  It is
// compiled together with the reference's own translation units where they lie under
// /root/reference (see oracle/Makefile; nothing from the reference is copied into this repo) into
// oracle/_ref/libt3dref.so.  It exists so that tests/, bench.py's cpu_baseline/reference legs and
// the golden-vector generator can drive tractor::ParticleEmitter headless:
//   * rand() is defined here (the reference path calls libc rand(): PE.cpp:359, pch.h:151-152), so the
//     draw stream can be generated deterministically, recorded, and replayed;
//   * Node / Scene / Camera / Texture / MeshBatch are the overlay stubs in oracle/ref_build/overlay;
//   * private members are read with -fno-access-control (this TU only).
//
// "PE.cpp" below = Tractor3Dlib/src/graphics/ParticleEmitter.cpp of the reference.
#include "pch.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "framework/FileSystem.h"
#include "framework/Platform.h"
#include "graphics/ParticleEmitter.h"
#include "animation/AnimationTarget.h"
#include "animation/AnimationValue.h"
#include "graphics/Sprite.h"
#include "graphics/SpriteBatch.h"
#include "graphics/TileSet.h"
#include "renderer/Font.h"
#include "scene/Node.h"
#include "scene/Scene.h"

#include "t3d_particles.h" // only for t3d_emitter_params and the T3D_COL_* column order

// ---- symbols the reference expects its platform layer to provide ---------------------------
GLenum __gl_error_code = 0;
ALenum __al_error_code = 0;

namespace tractor
{
static int g_log_quiet = 0;
void Logger::log(Level level, const char* message, ...)
{
    if (g_log_quiet && level != LEVEL_ERROR) return;
    va_list args;
    va_start(args, message);
    vfprintf(stderr, message, args);
    va_end(args);
}
void print(const char* format, ...)
{
    va_list args;
    va_start(args, format);
    vfprintf(stderr, format, args);
    va_end(args);
}
// ui/Sprite.cpp is an AnimationTarget; the animation system is not part of the harness: the few symbols it needs, inert
AnimationTarget::~AnimationTarget() {}
int AnimationTarget::getPropertyId(TargetType, const std::string&) { return -1; }
void AnimationTarget::cloneInto(AnimationTarget*, NodeCloneContext&) const {}
float AnimationValue::getFloat(unsigned int) const { return 0.0f; }
void AnimationValue::setFloat(unsigned int, float) {}
float Curve::lerp(float t, float from, float to) { return from + (to - from) * t; }
std::string Platform::displayFileDialog(size_t, const std::string&, const std::string&, const std::string&,
                                        const std::string&)
{
    return std::string();
}
void Node::setDrawable(Drawable* drawable)
{
    _drawable = drawable;
    if (drawable) drawable->setNode(this);
}

// Texture::create without GL: "@<W>x<H>" gives a synthetic texture; otherwise the PNG IHDR is read
// (bytes 16..23 = big-endian width, height).  Only the size is ever used (PE.cpp:244-247).
Texture* Texture::create(const std::string& path, bool)
{
    unsigned int w = 0, h = 0;
    if (!path.empty() && path[0] == '@')
    {
        if (sscanf(path.c_str(), "@%ux%u", &w, &h) != 2) return nullptr;
    }
    else
    {
        std::string full = path;
        FILE* f = fopen(full.c_str(), "rb");
        if (!f)
        {
            full = FileSystem::getResourcePath() + path;
            f = fopen(full.c_str(), "rb");
        }
        if (!f) return nullptr;
        unsigned char hdr[24];
        size_t got = fread(hdr, 1, sizeof(hdr), f);
        fclose(f);
        if (got != sizeof(hdr) || memcmp(hdr + 1, "PNG", 3) != 0) return nullptr;
        w = (hdr[16] << 24) | (hdr[17] << 16) | (hdr[18] << 8) | hdr[19];
        h = (hdr[20] << 24) | (hdr[21] << 16) | (hdr[22] << 8) | hdr[23];
    }
    if (!w || !h) return nullptr;
    Texture* t = new Texture();
    t->_width = w;
    t->_height = h;
    return t;
}
} // namespace tractor

// ---- rand(): generator / replay / record ----------------------------------------------------
namespace
{
uint64_t g_seq_state = 0x1234ULL;
std::vector<int32_t> g_replay;
size_t g_replay_pos = 0;
bool g_replaying = false;
bool g_recording = false;
std::vector<int32_t> g_recorded;
uint64_t g_rand_calls = 0;

inline uint64_t splitmix64(uint64_t& s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
} // namespace

// Bound inside this library with -Wl,-Bsymbolic so the reference objects call it, not libc's.
extern "C" int rand(void) noexcept
{
    int32_t r;
    if (g_replaying)
    {
        if (g_replay_pos >= g_replay.size())
        {
            fprintf(stderr, "t3dref: replay stream exhausted after %zu draws\n", g_replay_pos);
            abort();
        }
        r = g_replay[g_replay_pos++];
    }
    else
    {
        r = (int32_t)(splitmix64(g_seq_state) >> 33); // 31 bits: [0, RAND_MAX] with glibc's RAND_MAX = 2^31-1
    }
    ++g_rand_calls;
    if (g_recording) g_recorded.push_back(r);
    return r;
}

using namespace tractor;

namespace
{
struct RefEmitter
{
    ParticleEmitter* pe{ nullptr };
    unsigned int alloc_max{ 0 }; // _particleCountMax at create = the size of the _particles array (PE.cpp:31)
    Node node;       // the emitter's node (identity by default)
    Node cameraNode; // the active camera's node
    Camera camera;
    Scene scene;
};

// SpriteBatch.cpp keeps its sprite effect in a file-level static and drops it with the last batch (SB.cpp:49-65) -- but
// the batch's material still holds a reference when that destructor runs, so the static is left dangling once the last
// batch of the process is gone and the next SpriteBatch::create() would touch freed memory.  A harness that creates and
// destroys batches all day keeps one alive for good.
void keep_sprite_effect_alive()
{
    static SpriteBatch* keeper = nullptr;
    if (keeper) return;
    Texture* tex = Texture::create(std::string("@1x1"));
    keeper = SpriteBatch::create(tex);
    tex->release();
}

void wire(RefEmitter* r)
{
    r->alloc_max = r->pe->_particleCountMax;
    r->camera._node = &r->cameraNode;
    r->scene._camera = &r->camera;
    r->node._scene = &r->scene;
    r->node.setDrawable(r->pe);
}

void apply_params(ParticleEmitter* e, const t3d_emitter_params* p)
{
    // Same order as ParticleEmitter::create(Properties*) applies them, PE.cpp:193-208.
    e->setEmissionRate(p->emission_rate);
    e->setEllipsoid(p->ellipsoid != 0);
    e->setSize(p->size_start_min, p->size_start_max, p->size_end_min, p->size_end_max);
    e->_energyMin = p->energy_min; // setEnergy takes long (PE.cpp:381); write the float members directly
    e->_energyMax = p->energy_max;
    e->setColor(Vector4(p->color_start), Vector4(p->color_start_var), Vector4(p->color_end), Vector4(p->color_end_var));
    e->setPosition(Vector3(p->position), Vector3(p->position_var));
    e->setVelocity(Vector3(p->velocity), Vector3(p->velocity_var));
    e->setAcceleration(Vector3(p->acceleration), Vector3(p->acceleration_var));
    e->setRotationPerParticle(p->rotation_per_particle_speed_min, p->rotation_per_particle_speed_max);
    e->setRotation(p->rotation_speed_min, p->rotation_speed_max, Vector3(p->rotation_axis), Vector3(p->rotation_axis_var));
    e->setSpriteAnimated(p->sprite_animated != 0);
    e->setSpriteLooped(p->sprite_looped != 0);
    e->setSpriteFrameRandomOffset((int)p->sprite_frame_random_offset);
    e->setSpriteFrameDuration((long)p->sprite_frame_duration);
    e->setOrbit(p->orbit_position != 0, p->orbit_velocity != 0, p->orbit_acceleration != 0);
    e->setBlendMode((ParticleEmitter::BlendMode)p->blend_mode);
}
} // namespace

extern "C" {

// ---- rand control ----
void t3dref_rand_seed(uint64_t seed) { g_seq_state = seed; g_replaying = false; }
void t3dref_rand_replay(const int32_t* draws, size_t n)
{
    if (!draws) { g_replaying = false; g_replay.clear(); g_replay_pos = 0; return; }
    g_replay.assign(draws, draws + n);
    g_replay_pos = 0;
    g_replaying = true;
}
size_t t3dref_rand_replay_remaining(void) { return g_replaying ? g_replay.size() - g_replay_pos : 0; }
void t3dref_rand_record(int on) { g_recording = on != 0; g_recorded.clear(); }
size_t t3dref_rand_recorded(int32_t* out, size_t max)
{
    size_t n = g_recorded.size() < max ? g_recorded.size() : max;
    if (out && n) memcpy(out, g_recorded.data(), n * sizeof(int32_t));
    return g_recorded.size();
}
uint64_t t3dref_rand_calls(void) { return g_rand_calls; }
void t3dref_set_quiet(int q) { tractor::g_log_quiet = q; }
void t3dref_set_resource_path(const char* path) { FileSystem::setResourcePath(path); }
size_t t3dref_sizeof_particle(void) { return sizeof(ParticleEmitter::Particle); }

// ---- lifecycle ----
void* t3dref_emitter_create(const t3d_emitter_params* p)
{
    keep_sprite_effect_alive();
    char tex[64];
    snprintf(tex, sizeof(tex), "@%ux%u", (unsigned)p->texture_width, (unsigned)p->texture_height);
    ParticleEmitter* e = ParticleEmitter::create(std::string(tex), (ParticleEmitter::BlendMode)p->blend_mode,
                                                 p->particle_count_max);
    if (!e) return nullptr;
    apply_params(e, p);
    RefEmitter* r = new RefEmitter();
    r->pe = e;
    wire(r);
    return r;
}

void* t3dref_emitter_create_from_file(const char* url)
{
    keep_sprite_effect_alive();
    ParticleEmitter* e = ParticleEmitter::create(std::string(url));
    if (!e) return nullptr;
    RefEmitter* r = new RefEmitter();
    r->pe = e;
    wire(r);
    return r;
}

void* t3dref_emitter_clone(void* h)
{
    RefEmitter* src = (RefEmitter*)h;
    NodeCloneContext ctx;
    ParticleEmitter* e = static_cast<ParticleEmitter*>(src->pe->clone(ctx));
    RefEmitter* r = new RefEmitter();
    r->pe = e;
    wire(r);
    return r;
}

void t3dref_emitter_destroy(void* h)
{
    RefEmitter* r = (RefEmitter*)h;
    if (!r) return;
    r->node.setDrawable(nullptr);
    if (r->pe) r->pe->release();
    delete r;
}

void t3dref_emitter_set_params(void* h, const t3d_emitter_params* p)
{
    RefEmitter* r = (RefEmitter*)h;
    apply_params(r->pe, p);
    // setParticleCountMax (PE.h:237) only rewrites the field: going beyond the array allocated at create would overrun it
    if (p->particle_count_max && p->particle_count_max <= r->alloc_max) r->pe->setParticleCountMax(p->particle_count_max);
}

void t3dref_emitter_get_params(void* h, t3d_emitter_params* p)
{
    ParticleEmitter* e = ((RefEmitter*)h)->pe;
    memset(p, 0, sizeof(*p));
    p->particle_count_max = e->_particleCountMax;
    p->emission_rate = e->_emissionRate;
    p->ellipsoid = e->_ellipsoid;
    p->size_start_min = e->_sizeStartMin; p->size_start_max = e->_sizeStartMax;
    p->size_end_min = e->_sizeEndMin; p->size_end_max = e->_sizeEndMax;
    p->energy_min = e->_energyMin; p->energy_max = e->_energyMax;
    memcpy(p->color_start, &e->_colorStart.x, 16); memcpy(p->color_start_var, &e->_colorStartVar.x, 16);
    memcpy(p->color_end, &e->_colorEnd.x, 16); memcpy(p->color_end_var, &e->_colorEndVar.x, 16);
    memcpy(p->position, &e->_position.x, 12); memcpy(p->position_var, &e->_positionVar.x, 12);
    memcpy(p->velocity, &e->_velocity.x, 12); memcpy(p->velocity_var, &e->_velocityVar.x, 12);
    memcpy(p->acceleration, &e->_acceleration.x, 12); memcpy(p->acceleration_var, &e->_accelerationVar.x, 12);
    p->rotation_per_particle_speed_min = e->_rotationPerParticleSpeedMin;
    p->rotation_per_particle_speed_max = e->_rotationPerParticleSpeedMax;
    p->rotation_speed_min = e->_rotationSpeedMin; p->rotation_speed_max = e->_rotationSpeedMax;
    memcpy(p->rotation_axis, &e->_rotationAxis.x, 12); memcpy(p->rotation_axis_var, &e->_rotationAxisVar.x, 12);
    p->orbit_position = e->_orbitPosition; p->orbit_velocity = e->_orbitVelocity; p->orbit_acceleration = e->_orbitAcceleration;
    p->sprite_animated = e->_spriteAnimated; p->sprite_looped = e->_spriteLooped;
    p->sprite_frame_random_offset = e->_spriteFrameRandomOffset;
    p->sprite_frame_duration = e->_spriteFrameDuration;
    p->texture_width = e->_spriteTextureWidth; p->texture_height = e->_spriteTextureHeight;
    p->blend_mode = (int32_t)e->_spriteBlendMode;
}

void t3dref_emitter_set_emission_rate(void* h, uint32_t rate) { ((RefEmitter*)h)->pe->setEmissionRate(rate); }
void t3dref_emitter_set_sprite_tex_coords(void* h, uint32_t n, const float* c)
{
    ((RefEmitter*)h)->pe->setSpriteTexCoords(n, const_cast<float*>(c));
}
void t3dref_emitter_set_sprite_frame_rects(void* h, uint32_t n, const float* rects)
{
    std::vector<Rectangle> r(n);
    for (uint32_t i = 0; i < n; ++i) r[i] = Rectangle(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]);
    ((RefEmitter*)h)->pe->setSpriteFrameCoords(n, r.data());
}
void t3dref_emitter_set_sprite_frame_coords(void* h, uint32_t n, int w, int hgt)
{
    ((RefEmitter*)h)->pe->setSpriteFrameCoords(n, w, hgt);
}
uint32_t t3dref_emitter_get_sprite_tex_coords(void* h, float* out, size_t max_frames)
{
    ParticleEmitter* e = ((RefEmitter*)h)->pe;
    uint32_t n = e->_spriteFrameCount;
    if (out) memcpy(out, e->_spriteTextureCoords, sizeof(float) * 4 * (n < max_frames ? n : max_frames));
    return n;
}
float t3dref_emitter_get_time_per_emission(void* h) { return ((RefEmitter*)h)->pe->_timePerEmission; }
float t3dref_emitter_get_emit_time(void* h) { return ((RefEmitter*)h)->pe->_emitTime; }
uint32_t t3dref_emitter_get_sprite_width(void* h) { return ((RefEmitter*)h)->pe->getSpriteWidth(); }
uint32_t t3dref_emitter_get_sprite_height(void* h) { return ((RefEmitter*)h)->pe->getSpriteHeight(); }

void t3dref_emitter_set_node_world(void* h, const float m[16]) { ((RefEmitter*)h)->node._world.set(m); }
void t3dref_emitter_set_camera_world(void* h, const float m[16]) { ((RefEmitter*)h)->cameraNode._world.set(m); }

// ---- the path ----
void t3dref_emitter_start(void* h) { ((RefEmitter*)h)->pe->start(); }
void t3dref_emitter_stop(void* h) { ((RefEmitter*)h)->pe->stop(); }
int t3dref_emitter_is_started(void* h) { return ((RefEmitter*)h)->pe->isStarted(); }
int t3dref_emitter_is_active(void* h) { return ((RefEmitter*)h)->pe->isActive(); }
void t3dref_emitter_emit_once(void* h, uint32_t n) { ((RefEmitter*)h)->pe->emitOnce(n); }
void t3dref_emitter_update(void* h, float elapsed_ms) { ((RefEmitter*)h)->pe->update(elapsed_ms); }
uint32_t t3dref_emitter_draw(void* h)
{
    ParticleEmitter* e = ((RefEmitter*)h)->pe;
    e->_spriteBatch->_batch->submitted = false; // an inactive or empty emitter submits nothing (PE.cpp:837-839)
    return e->draw(false);
}
uint32_t t3dref_emitter_get_count(void* h) { return ((RefEmitter*)h)->pe->getParticlesCount(); }

// Live particles as T3D_NUM_COLUMNS columns of `stride` words (same layout as t3d_emitter_download_state).
uint32_t t3dref_emitter_get_state(void* h, uint32_t* out, size_t stride)
{
    ParticleEmitter* e = ((RefEmitter*)h)->pe;
    uint32_t n = e->_particleCount;
    if (!out) return n;
    auto F = [&](int col, size_t i, float v) { memcpy(&out[(size_t)col * stride + i], &v, 4); };
    for (uint32_t i = 0; i < n; ++i)
    {
        const ParticleEmitter::Particle& p = e->_particles[i];
        const float* cs = &p._colorStart.x; const float* ce = &p._colorEnd.x; const float* c = &p._color.x;
        for (int k = 0; k < 4; ++k) { F(T3D_COL_COLOR_START + k, i, cs[k]); F(T3D_COL_COLOR_END + k, i, ce[k]); F(T3D_COL_COLOR + k, i, c[k]); }
        const float* ps = &p._position.x; const float* v = &p._velocity.x; const float* a = &p._acceleration.x; const float* ax = &p._rotationAxis.x;
        for (int k = 0; k < 3; ++k) { F(T3D_COL_POSITION + k, i, ps[k]); F(T3D_COL_VELOCITY + k, i, v[k]); F(T3D_COL_ACCELERATION + k, i, a[k]); F(T3D_COL_ROTATION_AXIS + k, i, ax[k]); }
        out[(size_t)T3D_COL_ENERGY_START * stride + i] = (uint32_t)(int32_t)p._energyStart;
        out[(size_t)T3D_COL_ENERGY * stride + i] = (uint32_t)(int32_t)p._energy;
        F(T3D_COL_SIZE_START, i, p._sizeStart); F(T3D_COL_SIZE_END, i, p._sizeEnd); F(T3D_COL_SIZE, i, p._size);
        F(T3D_COL_ROT_PER_PARTICLE_SPEED, i, p._rotationPerParticleSpeed); F(T3D_COL_ROTATION_SPEED, i, p._rotationSpeed);
        F(T3D_COL_ANGLE, i, p._angle); F(T3D_COL_TIME_ON_FRAME, i, p._timeOnCurrentFrame);
        out[(size_t)T3D_COL_FRAME * stride + i] = p._frame;
    }
    return n;
}

// Overwrites the live particles (used to build large synthetic populations without spawning).
int t3dref_emitter_set_state(void* h, const uint32_t* in, size_t stride, uint32_t n)
{
    ParticleEmitter* e = ((RefEmitter*)h)->pe;
    if (n > e->_particleCountMax) return 1;
    auto F = [&](int col, size_t i) { float v; memcpy(&v, &in[(size_t)col * stride + i], 4); return v; };
    for (uint32_t i = 0; i < n; ++i)
    {
        ParticleEmitter::Particle& p = e->_particles[i];
        p._colorStart.set(F(0, i), F(1, i), F(2, i), F(3, i));
        p._colorEnd.set(F(4, i), F(5, i), F(6, i), F(7, i));
        p._color.set(F(8, i), F(9, i), F(10, i), F(11, i));
        p._position.set(F(12, i), F(13, i), F(14, i));
        p._velocity.set(F(15, i), F(16, i), F(17, i));
        p._acceleration.set(F(18, i), F(19, i), F(20, i));
        p._rotationAxis.set(F(21, i), F(22, i), F(23, i));
        p._energyStart = (int32_t)in[(size_t)T3D_COL_ENERGY_START * stride + i];
        p._energy = (int32_t)in[(size_t)T3D_COL_ENERGY * stride + i];
        p._sizeStart = F(26, i); p._sizeEnd = F(27, i); p._size = F(28, i);
        p._rotationPerParticleSpeed = F(29, i); p._rotationSpeed = F(30, i); p._angle = F(31, i);
        p._timeOnCurrentFrame = F(32, i);
        p._frame = in[(size_t)T3D_COL_FRAME * stride + i];
    }
    e->_particleCount = n;
    return 0;
}

// Vertex stream captured from the last draw(): 36 floats per quad (4 x SpriteVertex, SB.h:360-380).
uint32_t t3dref_emitter_get_vertices(void* h, float* out, size_t max_quads)
{
    ParticleEmitter* e = ((RefEmitter*)h)->pe;
    MeshBatch* mb = e->_spriteBatch->_batch.get();
    size_t quads = mb->submitted ? mb->capturedVertexCount() / 4 : 0; // what the last draw() handed to GL
    if (out)
    {
        size_t n = quads < max_quads ? quads : max_quads;
        size_t bytes = n * 4 * sizeof(SpriteBatch::SpriteVertex);
        if (bytes > mb->capturedBytes().size()) bytes = mb->capturedBytes().size();
        memcpy(out, mb->capturedBytes().data(), bytes);
    }
    return (uint32_t)quads;
}
void t3dref_emitter_set_discard_vertices(void* h, int discard)
{
    ((RefEmitter*)h)->pe->_spriteBatch->_batch->discard = discard != 0;
}

// ---- timing loops (steady_clock around the reference's own update()/draw(); seconds) ----
double t3dref_time_update(void* h, float elapsed_ms, int frames)
{
    ParticleEmitter* e = ((RefEmitter*)h)->pe;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; ++i) e->update(elapsed_ms);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
double t3dref_time_draw(void* h, int frames)
{
    ParticleEmitter* e = ((RefEmitter*)h)->pe;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; ++i) e->draw(false);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
double t3dref_time_frames(void* h, float elapsed_ms, int frames)
{
    ParticleEmitter* e = ((RefEmitter*)h)->pe;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; ++i) { e->update(elapsed_ms); e->draw(false); }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- SpriteBatch's 2-D overloads and TileSet, driven headless ----
static uint32_t copy_captured(MeshBatch* mb, float* out, size_t max_quads)
{
    const size_t quads = mb->capturedVertexCount() / 4;
    if (out)
    {
        size_t bytes = (quads < max_quads ? quads : max_quads) * 4 * sizeof(SpriteBatch::SpriteVertex);
        if (bytes > mb->capturedBytes().size()) bytes = mb->capturedBytes().size();
        memcpy(out, mb->capturedBytes().data(), bytes);
    }
    return (uint32_t)quads;
}

// One SpriteBatch::draw call per t3d_sprite2d, the overload the flags name; returns the quads the batch received.
// time_s (optional): seconds spent between start() and finish().
uint32_t t3dref_sprites_2d(const t3d_sprite2d* sp, size_t n, uint32_t tex_w, uint32_t tex_h, float* out, size_t max_quads, double* time_s)
{
    keep_sprite_effect_alive();
    char name[64];
    snprintf(name, sizeof(name), "@%ux%u", tex_w, tex_h);
    Texture* tex = Texture::create(std::string(name));
    if (!tex) return 0;
    SpriteBatch* b = SpriteBatch::create(tex);
    tex->release();
    auto t0 = std::chrono::steady_clock::now();
    b->start();
    for (size_t i = 0; i < n; ++i)
    {
        const t3d_sprite2d& s = sp[i];
        const Vector4 color(s.color[0], s.color[1], s.color[2], s.color[3]);
        if (s.flags & T3D_SPRITE_CLIPPED)
            b->draw(s.x, s.y, s.z, s.width, s.height, s.u1, s.v1, s.u2, s.v2, color, Rectangle(s.clip[0], s.clip[1], s.clip[2], s.clip[3]));
        else if (s.flags & T3D_SPRITE_ROTATED)
            b->draw(s.x, s.y, s.z, s.width, s.height, s.u1, s.v1, s.u2, s.v2, color, Vector2(s.rotation_point[0], s.rotation_point[1]),
                    s.rotation_angle, (s.flags & T3D_SPRITE_CENTERED) != 0);
        else
            b->draw(s.x, s.y, s.z, s.width, s.height, s.u1, s.v1, s.u2, s.v2, color, (s.flags & T3D_SPRITE_CENTERED) != 0);
    }
    b->finish();
    if (time_s) *time_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const uint32_t quads = copy_captured(b->_batch.get(), out, max_quads);
    delete b;
    return quads;
}

namespace
{
struct RefTileSet
{
    TileSet* ts{ nullptr };
    Node node, cameraNode;
    Camera camera;
    Scene scene;
};
} // namespace

void* t3dref_tileset_create(uint32_t tex_w, uint32_t tex_h, float tile_w, float tile_h, uint32_t rows, uint32_t cols)
{
    keep_sprite_effect_alive();
    char name[64];
    snprintf(name, sizeof(name), "@%ux%u", tex_w, tex_h);
    TileSet* ts = TileSet::create(std::string(name), tile_w, tile_h, rows, cols);
    if (!ts) return nullptr;
    RefTileSet* r = new RefTileSet();
    r->ts = ts;
    r->camera._node = &r->cameraNode;
    r->scene._camera = &r->camera;
    r->node._scene = &r->scene;
    r->node.setDrawable(ts);
    return r;
}
void t3dref_tileset_destroy(void* h)
{
    RefTileSet* r = (RefTileSet*)h;
    if (!r) return;
    r->node.setDrawable(nullptr);
    r->ts->release();
    delete r;
}
void t3dref_tileset_set_tile_source(void* h, uint32_t col, uint32_t row, float x, float y)
{
    ((RefTileSet*)h)->ts->setTileSource(col, row, Vector2(x, y));
}
// TileSet::draw with the node at node_t and the active camera's node at camera_t (world translations).
uint32_t t3dref_tileset_draw(void* h, const float node_t[3], const float camera_t[3], const float color[4], float opacity, float* out,
                             size_t max_quads)
{
    RefTileSet* r = (RefTileSet*)h;
    for (int k = 0; k < 3; ++k)
    {
        r->node._world.m[12 + k] = node_t[k];
        r->cameraNode._world.m[12 + k] = camera_t[k];
    }
    r->ts->setColor(Vector4(color[0], color[1], color[2], color[3]));
    r->ts->setOpacity(opacity);
    r->ts->draw(false);
    return copy_captured(r->ts->_batch->_batch.get(), out, max_quads);
}

// ---- Font::drawText (src/renderer/Font.cpp:242-390), driven headless ----
static_assert(sizeof(Font::Glyph) == sizeof(t3d_glyph), "t3d_glyph mirrors Font::Glyph (Font.h:296-323)");
void* t3dref_font_create(uint32_t size, const t3d_glyph* glyphs, uint32_t glyph_count, uint32_t tex_w, uint32_t tex_h)
{
    keep_sprite_effect_alive();
    char name[64];
    snprintf(name, sizeof(name), "@%ux%u", tex_w, tex_h);
    Texture* tex = Texture::create(std::string(name));
    if (!tex) return nullptr;
    std::vector<Font::Glyph> g(glyph_count);
    memcpy(g.data(), glyphs, glyph_count * sizeof(Font::Glyph));
    Font* f = Font::create(std::string("t3d"), Font::PLAIN, size, g.data(), (int)glyph_count, tex, Font::BITMAP); // Font.cpp:103-161
    tex->release();
    return f;
}
void t3dref_font_destroy(void* h)
{
    if (h) ((Font*)h)->release();
}
void t3dref_font_set_character_spacing(void* h, float spacing) { ((Font*)h)->setCharacterSpacing(spacing); }
// start(), drawText(text, x, y, color, size, rightToLeft), finish(); returns the quads the font's batch received.
uint32_t t3dref_font_draw_text(void* h, const char* text, size_t length, int x, int y, const float color[4], uint32_t size, int right_to_left,
                               float* out, size_t max_quads, double* time_s)
{
    Font* f = (Font*)h;
    const std::string s(text, length);
    auto t0 = std::chrono::steady_clock::now();
    f->start();
    f->drawText(s, x, y, Vector4(color[0], color[1], color[2], color[3]), size, right_to_left != 0);
    f->finish();
    if (time_s) *time_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return copy_captured(f->_batch->_batch.get(), out, max_quads);
}

void t3dref_font_measure_text(void* h, const char* text, size_t length, uint32_t size, uint32_t* width, uint32_t* height)
{
    unsigned int w = 0, hh = 0;
    ((Font*)h)->measureText(std::string(text, length), size, &w, &hh); // Font.cpp:675-731
    *width = w;
    *height = hh;
}

// ---- ui Sprite::draw (src/ui/Sprite.cpp:495-574), driven headless: a Sprite on a node under a scene with an active camera ----
uint32_t t3dref_ui_sprite_draw(const t3d_ui_sprite* in, uint32_t tex_w, uint32_t tex_h, float* out, size_t max_quads)
{
    keep_sprite_effect_alive();
    char name[64];
    snprintf(name, sizeof(name), "@%ux%u", tex_w, tex_h);
    Sprite* sp = Sprite::create(std::string(name), in->width, in->height, Rectangle(in->frame[0], in->frame[1], in->frame[2], in->frame[3]), 1, nullptr);
    if (!sp) return 0;
    sp->setOffset((Sprite::Offset)in->offset);
    sp->setAnchor(Vector2(in->anchor[0], in->anchor[1]));
    sp->setFlip((int)in->flip);
    sp->setColor(Vector4(in->color[0], in->color[1], in->color[2], in->color[3]));
    sp->setOpacity(in->opacity);
    Node node, cameraNode;
    Camera camera;
    Scene scene;
    camera._node = &cameraNode;
    scene._camera = &camera;
    node._scene = &scene;
    for (int k = 0; k < 3; ++k) node._world.m[12 + k] = in->node_translation[k];
    cameraNode._world.m[12] = in->camera_translation[0];
    cameraNode._world.m[13] = in->camera_translation[1];
    node._rotation = Quaternion(in->node_rotation[0], in->node_rotation[1], in->node_rotation[2], in->node_rotation[3]);
    node._scaleX = in->node_scale[0];
    node._scaleY = in->node_scale[1];
    sp->setNode(&node);
    sp->draw(false);
    const uint32_t quads = copy_captured(sp->_batch->_batch.get(), out, max_quads);
    sp->setNode(nullptr);
    sp->release();
    return quads;
}

} // extern "C"
