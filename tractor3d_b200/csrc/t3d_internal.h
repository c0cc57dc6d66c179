// t3d_internal.h -- device-visible records shared by the kernels (t3d_kernels.cu) and the host runtime
// (t3d_runtime.cu).  Nothing here is part of the ABI.
#pragma once

#include <cstddef>
#include <cstdint>

#include "../../include/t3d_particles.h"

namespace t3d
{

// Particles per tile of the update kernel = particles per thread x threads per CTA x groups of the selected variant
// (t3d_update.cu).
int update_tile_size(int shape = 0);
int draw_sub(uint32_t n_items);
int draw_tile_size(uint32_t n_items); // quads per draw CTA for a launch over n_items emitters
const char* update_variant_name();
constexpr int kSpawnThreads = 256; // spawn and draw: one particle per thread
// A spawn CTA generates this many consecutive groups of kSpawnThreads particles.  Measured with 4: the descriptor fetch is paid a
// quarter as often, but 1 Mi particles are then 1 024 CTAs = 1.7 waves of the 592 the device holds, and the idle part of the
// last wave costs more than the fetches saved (50.1 against 47.6 us per 1 Mi particles); 1 = 6.9 waves.
constexpr int kSpawnSub = 1;
constexpr int kSpawnTile = kSpawnThreads * kSpawnSub;
constexpr int kDrawThreads = 256;
constexpr int kMaxSmemFrames = 256; // sprite uv tables up to this many frames are staged in shared memory (k_draw)
constexpr int kMaxFusedFrames = 64; // the fused update + draw pass has 1 KB for the table; larger sheets take the two launches

// Segment alignment inside a slab, in particles (a warp of the update kernel copies the life records of 64
// consecutive particles as one 4 KB run).
constexpr uint32_t kSegmentAlign = 64;

// ---- particle storage in HBM (private to the library; the ABI exchanges the 34 T3D_COL_* columns) ----------
// The 34 words of ParticleEmitter::Particle (PE.h:792-822) are split by how the frame touches them:
//   rw  : 15 columns of `stride` words, struct-of-arrays -- everything update() rewrites (or draw() reads) every frame.
//         A thread owns consecutive particles, so every access is a unit-stride vector access.
//   rec : one 64-byte LIFE RECORD per particle, array-of-structs -- the 15 words a particle is born with and keeps
//         (colour and size end points, acceleration, spin, initial energy).  The update streams them with warp-wide
//         contiguous copies; swap-with-last removal (PE.cpp:821-830) moves a record as two full 32-byte sectors
//         instead of 15 scattered words that each cost a sector fill and a write-back.
//   rot : one 16-byte record per particle: rotation axis and speed, only read by emitters that can rotate
//         (PE.cpp:757-763; no stock .particle file sets them).
enum : int
{
    RW_ENERGY = 0, RW_POS = 1, RW_VEL = 4, RW_ANGLE = 7, RW_COLOR = 8, RW_SIZE = 12, RW_TIME_ON_FRAME = 13, RW_FRAME = 14,
    RW_COUNT = 15
};
enum : int
{
    REC_COLOR_START = 0, REC_COLOR_END = 4, REC_ACC = 8, REC_SPIN = 11, REC_ENERGY_START = 12, REC_SIZE_START = 13,
    REC_SIZE_END = 14, REC_PAD = 15, REC_WORDS = 16
};
enum : int { ROT_AXIS = 0, ROT_SPEED = 3, ROT_WORDS = 4 };
constexpr uint32_t kBytesPerParticle = (RW_COUNT + REC_WORDS + ROT_WORDS) * 4; // 140

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
    kFlagMayRotate = 4u,  // some live particle may have rotationSpeed != 0: the rotation records are read
    kFlagMaySpin = 8u,    // some live particle may have angle != 0 (used by the frozen-rotation latch)
    kFlagBounds = 16u,    // the update also reduces the emitter's bounding box (t3d_emitter_bounds)
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

// Where one emitter's particles live: its segment of the slab's three arrays.
struct Storage
{
    uint32_t* rw;    // column 0 of the segment; column c starts at rw + c * stride
    uint32_t* rec;   // life record of particle i: the REC_WORDS words at rec + REC_WORDS * i (64-byte aligned)
    uint32_t* rot;   // rotation record of particle i: the ROT_WORDS words at rot + ROT_WORDS * i (16-byte aligned)
    uint32_t stride; // words between rw columns: the slab capacity
    uint32_t seg_capacity; // padded segment size: any particle index below it lies inside the slab and may be read
};

// Host-written part of the per-emitter device record.
struct EmitterConst
{
    Storage st;
    float* vtx;       // vertex stream of this emitter: 36 floats per quad
    const float* uv;  // frame_count * 4 floats (u1, v1, u2, v2)
    uint32_t capacity;
    uint32_t flags;
    uint32_t frame_count;
    float percent_per_frame;
    float frame_duration_secs;
    uint32_t pad;
    SpawnParams sp;
};

// Bounding box of the live particles after an update, reduced by the update kernel with 64-bit atomicMax on
// (launch epoch << 32 | order-preserving float bits): a newer launch's first contribution overrides whatever an older
// launch left, so nothing has to be reset between frames.  lo[] holds the complement of the encoding (max = min).
struct EmitterBounds
{
    unsigned long long lo[3], hi[3];
};

struct EmitterDev
{
    EmitterCounts cnt[2]; // device-owned
    EmitterBounds bounds; // device-owned
    EmitterConst c;       // host-owned, re-uploaded when dirty
};

// One entry per emitter taking part in a launch.  tile_begin is the first global tile of the emitter in
// this launch; entries are sorted by it and a CTA finds its emitter by binary search.
// 80 bytes, 16-byte aligned: a CTA requests the whole entry of the emitter it expects to own with five 128-bit loads at once.
struct alignas(16) UpdateItem
{
    EmitterDev* dev;
    Storage st;           // copied from the emitter's record so a CTA can start loading after ONE dependent fetch
    float elapsed_ms;
    uint32_t parity;      // which cnt[] copy is current on entry
    uint32_t tile_begin;
    uint32_t flags;
    uint32_t frame_count;
    float percent_per_frame;
    float frame_duration_secs;
    uint32_t n_tiles;     // tiles of this launch that belong to the emitter (the kernel checks that they cover the live count)
    uint32_t pad[2];
};
static_assert(sizeof(UpdateItem) == 80, "UpdateItem is fetched as five 128-bit words");

// 112 bytes, 16-byte aligned; a CTA requests the first 48 (everything but the matrix) of the entry it expects to own at once.
struct alignas(16) SpawnItem
{
    EmitterDev* dev;
    const int32_t* draws;     // recorded draws for this emit (nullptr: device generator)
    const uint32_t* offsets;  // per-particle start offset into draws (nullptr: constant stride)
    uint32_t draw_stride;     // draws per particle when offsets == nullptr
    uint32_t request;         // particles asked for (clamped to free capacity on the device)
    uint32_t parity;
    uint32_t tile_begin;
    uint32_t n_tiles;         // tiles of this launch that belong to the emitter
    uint32_t pad;
    float world[16];          // node world matrix (PE.cpp:299)
};
static_assert(sizeof(SpawnItem) == 112 && offsetof(SpawnItem, world) == 48, "the head of a SpawnItem is three 128-bit words");

// (the first 80 bytes -- everything but the matrix -- are requested at once by the CTA that expects to own the entry)
struct alignas(16) DrawItem
{
    EmitterDev* dev;
    const uint32_t* rw;   // copied from the emitter's record: one dependent fetch, then the loads can start
    float* vtx;           // where this draw's quads go: the emitter's own stream, or the caller's buffer (t3d_emitter_draw_to)
    const float* uv;
    uint32_t parity;
    uint32_t tile_begin;
    uint32_t stride;
    uint32_t seg_capacity;
    uint32_t frame_count;
    uint32_t n_tiles;     // as in UpdateItem
    float right[3];
    float up[3];
    float vp[16];         // view-projection matrix (PE.cpp:846-849); only read by the clip-space draw (T3D_VERTEX_CLIP)
};
static_assert(sizeof(DrawItem) == 144 && offsetof(DrawItem, vp) == 80, "the head of a DrawItem is five 128-bit words");

// Process-wide billboard rotation of the reference (SB.cpp:339), kept on the device.
struct FrozenRotation
{
    float m[16];
    int latched;
    int pad[3];
};

// ---- launch wrappers (t3d_kernels.cu, t3d_update.cu) ----
// host_draws: some item reads recorded draws (otherwise the launch is the device generator's specialisation)
void launch_spawn(void* stream, const SpawnItem* items, uint32_t n_items, uint32_t total_tiles, bool host_draws);
// `fault`: device word the kernels OR into when a grid does not cover the live count it finds (kFaultUpdateGrid ...).
constexpr uint32_t kFaultUpdateGrid = 1u, kFaultDrawGrid = 2u;
struct UpdateLaunch
{
    const UpdateItem* items;
    const DrawItem* ditems; // fused launches only
    uint32_t n_items, total_tiles;
    unsigned long long* tile_status;
    uint32_t* ticket;
    uint32_t ticket_base;
    uint32_t epoch;
    uint32_t* fault;
    int rot_mode;
    const FrozenRotation* frozen;
    bool bounds; // some emitter of the launch wants its bounding box
};
// shape: 0 = the variant's own tiles, 1 = half-size tiles (pick_tile_shape decides; narrow launches lose less to their tail)
void launch_update(void* stream, const UpdateLaunch& L, int shape);
void launch_update_draw(void* stream, const UpdateLaunch& L, int shape);
int pick_tile_shape(uint64_t tiles_big, uint64_t tiles_small, uint32_t slots);
// vertex_mode: T3D_VERTEX_WORLD (SpriteVertex, 36 floats per quad) or T3D_VERTEX_CLIP (40 floats per quad)
void launch_draw(void* stream, const DrawItem* items, uint32_t n_items, uint32_t total_tiles, int rot_mode,
                 const FrozenRotation* frozen, uint32_t* fault, int vertex_mode);
int fused_tile_size(int shape = 0);
int fused_mode(); // 1 always, 0 never, -1 automatic
void launch_latch(void* stream, const DrawItem* items, uint32_t n_items, FrozenRotation* frozen);
struct CountQuery
{
    const EmitterDev* dev;
    uint32_t parity;
    uint32_t pad;
};
// out[2*i] = live count, out[2*i+1] = quads written by the last draw, for each query.
// out[2*n] = the context's fault word.  boxes (optional): the six bounding-box words of each query's emitter.
void launch_gather_counts(void* stream, const CountQuery* q, uint32_t n, uint32_t* out, const uint32_t* fault,
                          unsigned long long* boxes);
// MeshBatch's triangle-strip index pattern for n_quads quads (MB.cpp:117-141), 6*n_quads-2 indices.
void launch_strip_indices(void* stream, uint32_t* out, uint64_t n_indices);
void launch_triangle_indices(void* stream, uint32_t* out, uint64_t n_indices);
// Device storage <-> the ABI's exchange layout (T3D_NUM_COLUMNS columns of `host_stride` words).
void launch_export_state(void* stream, const Storage& st, uint32_t n, uint32_t* out, uint32_t host_stride);
void launch_import_state(void* stream, const Storage& st, uint32_t n, const uint32_t* in, uint32_t host_stride);

// ---- 2-D sprites and tile sets (t3d_sprites.cu) ----
// Font::drawText (Font.cpp:242-390): the per-call constants of one text.
struct TextParams
{
    int x, y;               // where the pen starts
    uint32_t size;          // requested size (the font's own when the caller passed 0)
    float scale;            // (float)size / font size, Font.cpp:268
    int spacing;            // (int)(size * character spacing), :269
    uint32_t glyph_count;
    uint32_t space_advance; // _glyphs[0].advance: what ' ' and '\t' move the pen by (:344, :351)
    float color[4];
};
// Per-character-code tables of one text (a signed char that can have a glyph: 32 .. 127): what the code adds to the pen and
// what it does.  Built on the host (the glyph table and the per-call constants are there), handed to the kernels by value.
constexpr int kTextClasses = 128;
struct TextTables
{
    int adv[kTextClasses];            // ' ', '\t', glyph advances (Font.cpp:344, :351, :378); 0 for everything else
    unsigned char kind[kTextClasses]; // 0 nothing, 1 line break, 2 glyph, 3 blank (' ' or '\t')
};
uint32_t text_tiles(uint32_t n);          // tiles (one expansion CTA each)
size_t text_scratch_bytes(uint32_t n);    // device scratch a text of n bytes needs; its last 48 bytes end up holding {., line breaks, glyphs, ...}
// text: n bytes on the device, 16-byte aligned; glyphs: the font's table on the device; quads go to dst in reading order
void launch_text(void* stream, const char* text, uint32_t n, const t3d_glyph* glyphs, const TextTables& T, const TextParams& P, bool right_to_left,
                 void* scratch, float* dst);
uint32_t sprite_tiles(uint32_t n); // CTAs (tiles of 256 sprites) a list of n sprites takes
// tile_scratch: sprite_tiles(n) + 1 words when some sprite may be culled by its clip rectangle (ordered compaction; the
// last word receives the number of quads written), nullptr when sprite i simply becomes quad i.
void launch_sprites(void* stream, const t3d_sprite2d* sprites, uint32_t n, uint32_t* tile_scratch, float* dst);
void launch_tileset(void* stream, const uint32_t* cells, uint32_t n_quads, const float* tiles_xy, uint32_t cols, const float* xs,
                    const float* ys, float z, float tile_w, float tile_h, float wr, float hr, const float color[4], float* dst);

// ---- depth sort (t3d_sort.cu) ----
size_t depth_sort_temp_bytes(uint32_t n);
// scratch: 4 * n words; temp: depth_sort_temp_bytes(n) bytes; out: 6 * n indices.  Returns 0 or the cudaError_t of the sort.
int launch_depth_sort(void* stream, const uint32_t* rw, size_t stride, uint32_t n, const float view_dir[3], uint32_t* scratch, void* temp,
                      size_t temp_bytes, uint32_t* out);

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
