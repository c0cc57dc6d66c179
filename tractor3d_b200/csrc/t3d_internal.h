// t3d_internal.h -- device-visible records shared by the kernels (t3d_kernels.cu) and the host runtime
// (t3d_runtime.cu).  Nothing here is part of the ABI.
#pragma once

#include <cstdint>

#include "../../include/t3d_particles.h"

namespace t3d
{

// Particles per tile of the update kernel = particles per thread x threads per CTA of the selected variant
// (t3d_update.cu).
int update_tile_size();
int draw_sub(uint32_t n_items);
int draw_tile_size(uint32_t n_items); // quads per draw CTA for a launch over n_items emitters
const char* update_variant_name();
// Spawn and draw: one particle per thread.
constexpr int kSpawnThreads = 256;
constexpr int kDrawThreads = 256;
constexpr int kMaxSmemFrames = 256; // sprite uv tables up to this many frames are staged in shared memory

// Segment alignment inside a slab, in particles (128 B per column).
constexpr uint32_t kSegmentAlign = 32;

// Live count and spawn serial of one emitter.  Two copies: a kernel reads cnt[parity] and writes
// cnt[parity ^ 1], so no CTA of a launch can observe the value another CTA of the same launch writes.
struct EmitterCounts
{
    uint32_t count;
    uint32_t drawn; // cnt[0].drawn: quads written by the last draw of this emitter
    uint64_t serial; // particles ever spawned (key of the device generator)
};

enum : uint32_t
{
    kFlagAnimated = 1u,   // _spriteAnimated
    kFlagLooped = 2u,     // _spriteLooped
    kFlagMayRotate = 4u,  // some live particle may have rotationSpeed != 0: axis/rotSpeed columns are read
    kFlagMaySpin = 8u,    // some live particle may have angle != 0 (used by the frozen-rotation latch)
};

// Spawn-time configuration: the float members of t3d_emitter_params the spawn kernel needs.
struct SpawnParams
{
    float color_start[4], color_start_var[4], color_end[4], color_end_var[4];
    float position[3], position_var[3], velocity[3], velocity_var[3], acceleration[3], acceleration_var[3];
    float rotation_axis[3], rotation_axis_var[3];
    float size_start_min, size_start_max, size_end_min, size_end_max;
    float energy_min, energy_max;
    float spin_min, spin_max;         // rotationPerParticleSpeed
    float rot_speed_min, rot_speed_max;
    uint32_t ellipsoid, orbit_position, orbit_velocity, orbit_acceleration;
    uint32_t frame_random_offset;
    uint32_t pad;
    uint64_t rng_seed;
};

// Host-written part of the per-emitter device record.
struct EmitterConst
{
    uint32_t* cols;   // column 0 of this emitter's segment; column c starts at cols + c * stride
    float* vtx;       // vertex stream of this emitter: 36 floats per quad
    const float* uv;  // frame_count * 4 floats (u1, v1, u2, v2)
    uint32_t stride;  // words between columns: the slab capacity (plain SoA) or the block size (blocked SoA)
    uint32_t block_shift; // 0: plain SoA.  k > 0: blocked SoA, 2^k particles x 34 columns contiguous per block
    uint32_t capacity;
    uint32_t flags;
    uint32_t frame_count;
    float percent_per_frame;
    float frame_duration_secs;
    SpawnParams sp;
};

struct EmitterDev
{
    EmitterCounts cnt[2]; // device-owned
    EmitterConst c;       // host-owned, re-uploaded when dirty
};

// One entry per emitter taking part in a launch.  tile_begin is the first global tile of the emitter in
// this launch; entries are sorted by it and a CTA finds its emitter by binary search.
struct UpdateItem
{
    EmitterDev* dev;
    uint32_t* cols;       // copied from the emitter's record so a CTA can start loading after ONE dependent fetch
    float elapsed_ms;
    uint32_t parity;      // which cnt[] copy is current on entry
    uint32_t tile_begin;
    uint32_t stride;
    uint32_t block_shift;
    uint32_t seg_capacity; // padded segment size: any slot below it lies inside the slab and may be read
    uint32_t flags;
    uint32_t frame_count;
    float percent_per_frame;
    float frame_duration_secs;
    uint32_t n_tiles;     // tiles of this launch that belong to the emitter (the kernel checks that they cover the live count)
    uint32_t pad;
};

struct SpawnItem
{
    EmitterDev* dev;
    const int32_t* draws;     // recorded draws for this emit (nullptr: device generator)
    const uint32_t* offsets;  // per-particle start offset into draws (nullptr: constant stride)
    uint32_t draw_stride;     // draws per particle when offsets == nullptr
    uint32_t request;         // particles asked for (clamped to free capacity on the device)
    uint32_t parity;
    uint32_t tile_begin;
    float world[16];          // node world matrix (PE.cpp:299)
};

struct DrawItem
{
    EmitterDev* dev;
    const uint32_t* cols; // copied from the emitter's record: one dependent fetch, then the loads can start
    float* vtx;
    const float* uv;
    uint32_t parity;
    uint32_t tile_begin;
    uint32_t stride;
    uint32_t block_shift;
    uint32_t seg_capacity;
    uint32_t frame_count;
    float right[3];
    float up[3];
    uint32_t n_tiles;     // as in UpdateItem
    uint32_t pad;
};

// Process-wide billboard rotation of the reference (SB.cpp:339), kept on the device.
struct FrozenRotation
{
    float m[16];
    int latched;
    int pad[3];
};

// ---- launch wrappers (t3d_kernels.cu) ----
void launch_spawn(void* stream, const SpawnItem* items, uint32_t n_items, uint32_t total_tiles);
// `fault`: device word the kernels OR into when a grid does not cover the live count it finds (kFaultUpdateGrid ...).
constexpr uint32_t kFaultUpdateGrid = 1u, kFaultDrawGrid = 2u;
void launch_update(void* stream, const UpdateItem* items, uint32_t n_items, uint32_t total_tiles,
                   unsigned long long* tile_status, uint32_t* ticket, uint32_t ticket_base, uint32_t epoch, uint32_t* fault);
void launch_draw(void* stream, const DrawItem* items, uint32_t n_items, uint32_t total_tiles, int rot_mode,
                 const FrozenRotation* frozen, uint32_t* fault);
int fused_tile_size();
int fused_mode(); // 1 always, 0 never, -1 automatic
void launch_update_draw(void* stream, const UpdateItem* items, const DrawItem* ditems, uint32_t n_items, uint32_t total_tiles,
                        unsigned long long* tile_status, uint32_t epoch, int rot_mode, const FrozenRotation* frozen,
                        uint32_t* fault);
void launch_latch(void* stream, const DrawItem* items, uint32_t n_items, FrozenRotation* frozen);
struct CountQuery
{
    const EmitterDev* dev;
    uint32_t parity;
    uint32_t pad;
};
// out[2*i] = live count, out[2*i+1] = quads written by the last draw, for each query.
// out[2*n] = the context's fault word.
void launch_gather_counts(void* stream, const CountQuery* q, uint32_t n, uint32_t* out, const uint32_t* fault);
// MeshBatch's triangle-strip index pattern for n_quads quads (MB.cpp:117-141), 6*n_quads-2 indices.
void launch_strip_indices(void* stream, uint32_t* out, uint64_t n_indices);
void launch_triangle_indices(void* stream, uint32_t* out, uint64_t n_indices);
// Device layout <-> the ABI's exchange layout (T3D_NUM_COLUMNS columns of `host_stride` words).
void launch_export_state(void* stream, const uint32_t* cols, uint32_t stride, uint32_t block_shift, uint32_t n, uint32_t* out,
                         uint32_t host_stride);
void launch_import_state(void* stream, uint32_t* cols, uint32_t stride, uint32_t block_shift, uint32_t n, const uint32_t* in,
                         uint32_t host_stride);

// The counter-based generator of T3D_RNG_DEVICE (host + device): 32-bit integer hash of (seed, particle serial,
// draw index).  key = the part that is constant per particle; one draw = key + (j+1)*C through the lowbias32
// finaliser (two multiplies, three xor-shifts), top 31 bits.  Cheap on purpose: a particle consumes ~30 draws and
// the spawn kernel is bound by exactly this arithmetic.
#if defined(__CUDACC__)
#define T3D_HD __host__ __device__
#else
#define T3D_HD
#endif
T3D_HD inline uint32_t device_rng_key(uint64_t seed, uint64_t serial)
{
    const uint32_t k = ((uint32_t)seed * 0x9E3779B1u) ^ ((uint32_t)(seed >> 32) * 0x85EBCA77u);
    return k ^ ((uint32_t)serial * 0xC2B2AE3Du) ^ ((uint32_t)(serial >> 32) * 0x27D4EB2Fu);
}
T3D_HD inline int32_t device_rng_from_key(uint32_t key, uint32_t j)
{
    uint32_t x = key + (j + 1u) * 0x165667B1u;
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return (int32_t)(x >> 1);
}
T3D_HD inline int32_t device_rng_draw(uint64_t seed, uint64_t serial, uint32_t j)
{
    return device_rng_from_key(device_rng_key(seed, serial), j);
}

} // namespace t3d
