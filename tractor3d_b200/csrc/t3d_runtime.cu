// t3d_runtime.cu -- host runtime behind the C ABI of include/t3d_particles.h.
//
// What lives here (all host code; the kernels are in t3d_kernels.cu):
//   * the per-GPU context: slabs of struct-of-arrays particle storage in HBM, the per-emitter device
//     records, pinned staging for per-launch descriptors, tile-status words for the removal scan;
//   * the emitter object: the reference's configuration members and the scalar per-call logic that stays
//     on the host exactly as the reference computes it -- isActive (PE.cpp:277-284), the process-wide
//     update rate cap (PE.cpp:720-725), the emission count (PE.cpp:729-746), the capacity clamp
//     (PE.cpp:293-296), the sprite uv table (PE.cpp:512-582);
//   * a command queue that coalesces consecutive emit / update / draw calls on different emitters into one
//     launch per phase, so a game loop that calls emitter->update() on 8 K emitters costs two launches.
//
// PE.cpp = Tractor3Dlib/src/graphics/ParticleEmitter.cpp of the reference.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "t3d_device.cuh"
#include "t3d_host.h"
#include "t3d_internal.h"

using namespace t3d;

// --------------------------------------------------------------------------------------------------
// errors
// --------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define T3D_CUDA(call)                                                                                          \
    do                                                                                                          \
    {                                                                                                           \
        cudaError_t err__ = (call);                                                                             \
        if (err__ != cudaSuccess) return fail(T3D_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(err__)); \
    } while (0)

#define T3D_TRY(expr)                  \
    do                                 \
    {                                  \
        int rc__ = (expr);             \
        if (rc__ != T3D_OK) return rc__; \
    } while (0)

// --------------------------------------------------------------------------------------------------
// objects
// --------------------------------------------------------------------------------------------------
namespace
{

enum Phase { PHASE_NONE = 0, PHASE_EMIT = 1, PHASE_UPDATE = 2, PHASE_DRAW = 3 };
enum { KERNEL_SPAWN = 0, KERNEL_UPDATE = 1, KERNEL_DRAW = 2, KERNEL_UPDATE_DRAW = 3, KERNEL_KINDS = 4 };

struct Slab
{
    // one allocation: RW_COUNT columns of `capacity` words, then `capacity` 64-byte life records, then `capacity`
    // 16-byte rotation records (t3d_internal.h)
    uint32_t* base{ nullptr };
    float* vtx{ nullptr };     // capacity * vtx_floats floats, allocated at the first draw that needs it
    uint32_t vtx_floats{ 0 };  // floats per quad the vertex storage was sized for (36 world, 40 clip)
    uint32_t capacity{ 0 };
    uint32_t used{ 0 };        // high-water mark: [used, capacity) has never been handed out
    uint32_t live_emitters{ 0 };
    std::vector<std::pair<uint32_t, uint32_t>> free_ranges; // (offset, particles) returned by destroyed emitters, sorted, coalesced
    uint32_t* rw() const { return base; }
    uint32_t* rec() const { return base + (size_t)RW_COUNT * capacity; }
    uint32_t* rot() const { return base + (size_t)(RW_COUNT + REC_WORDS) * capacity; }
};

struct StagingSlot
{
    char* host{ nullptr };
    char* dev{ nullptr };
    size_t cap{ 0 };
    cudaEvent_t done{ nullptr };
    cudaEvent_t uploaded{ nullptr };
    bool in_flight{ false };
};

struct PendingOp
{
    t3d_emitter* e;
    float elapsed_ms;     // update
    uint32_t emit_n;      // particles to spawn before the update (or the whole op for PHASE_EMIT)
    size_t draws_begin;   // into ctx->pending_draws, when the draws come from the host
    size_t draws_count;
    uint32_t spawned;     // update: particles asked to spawn since the emitter's previous update (churn estimate)
    size_t offsets_begin; // into ctx->pending_offsets (ellipsoid), or SIZE_MAX for constant stride
    uint32_t draw_stride;
    bool host_draws;
    float* draw_dst;      // draw: the caller's device buffer (t3d_emitter_draw_to), or nullptr for the emitter's own stream
    // (No copies of the node / camera / view-projection matrices: the setters launch whatever an emitter still has queued
    //  before they change them, so the emitter's own copies are what the call saw when the descriptors are cut.)
};

} // namespace

// What the reference keeps in function-local statics, i.e. once per PROCESS: the update rate-cap accumulator
// (`static double runningTime`, PE.cpp:720) and the billboard rotation latched by the first rotated quad ever drawn
// (`static Matrix rotation`, SB.cpp:339).  By default every context of the process shares one copy, like every emitter of
// the reference process does -- a host that drives one context per GPU from its game thread gets the reference's results
// whichever way the emitters are split over the GPUs.  t3d_ctx_set_private_statics gives a context its own copy (for
// hosts that drive contexts from different threads: the shared copy is not synchronised).
struct ProcessStatics
{
    double running_time{ 0 };
    bool frozen_latched{ false };
    float frozen_m[16]{};
    uint64_t frozen_version{ 1 }; // bumped whenever the matrix or its latched state changes
};
static ProcessStatics g_statics;
static std::vector<t3d_ctx*> g_contexts; // every live context of the process

struct t3d_ctx
{
    int device{ 0 };
    cudaStream_t stream{ nullptr };
    bool own_stream{ false };

    std::vector<Slab> slabs;
    uint32_t next_slab_particles{ 1u << 20 };

    std::vector<EmitterDev*> dev_chunks; // arena of device records, kDevChunk per allocation
    uint32_t dev_used{ 0 };
    std::vector<EmitterDev*> dev_free;

    StagingSlot staging[8];
    int staging_next{ 0 };
    // Launch descriptors travel on a stream of their own, so that the upload for the next launch overlaps the kernel
    // of the current one; the launching stream waits for the slot's `uploaded` event.  (T3D_COPY_STREAM=0: same stream.)
    cudaStream_t copy_stream{ nullptr };

    unsigned long long* tile_status{ nullptr };
    size_t tile_status_cap{ 0 };
    uint32_t* ticket{ nullptr };
    uint32_t ticket_base{ 0 };
    uint32_t epoch{ 0 };

    FrozenRotation* frozen_dev{ nullptr };
    uint64_t frozen_dev_version{ 0 }; // the version of the statics' matrix the device copy holds (0: none yet)
    ProcessStatics own_statics;       // used instead of the process-wide copy when private_statics is set
    bool private_statics{ false };
    ProcessStatics& statics() { return private_statics ? own_statics : g_statics; }
    const ProcessStatics& statics() const { return private_statics ? own_statics : g_statics; }
    int rot_mode{ T3D_ROT_FROZEN };
    int vertex_mode{ T3D_VERTEX_WORLD };
    uint32_t cta_slots{ 296 }; // CTAs of the update kernel the device runs at once (2 per SM)

    int phase{ PHASE_NONE };
    std::vector<PendingOp> pending;
    // Updates whose launch is deferred because a draw phase was opened right behind them: if the draws turn out to
    // cover the same emitters in the same order, both become one fused launch (T3D_FUSED=1).
    std::vector<PendingOp> held;
    std::vector<int32_t> pending_draws;
    std::vector<uint32_t> pending_offsets;

    uint64_t launches{ 0 }, h2d_bytes{ 0 }, d2h_bytes{ 0 };
    // Host time spent inside the library, by what it was spent on (t3d_ctx_host_profile): seconds since creation.
    double host_s[T3D_HOST_PROFILE_SLOTS]{};
    double host_child_s{ 0 }; // time already attributed by scopes nested inside the one that is open
    // Per-kind launch timing: every launch is bracketed by a pair of events on the stream; pairs are
    // resolved (and recycled) lazily when somebody asks for the totals.
    struct EventPair { cudaEvent_t start, stop; };
    std::vector<EventPair> ev_pending[KERNEL_KINDS], ev_free;
    double kernel_ms_total[KERNEL_KINDS]{ 0, 0, 0, 0 };
    uint64_t kernel_launches[KERNEL_KINDS]{ 0, 0, 0, 0 };
    float kernel_ms_last[KERNEL_KINDS]{ 0, 0, 0, 0 };
    bool timing_enabled{ true };

    // Asynchronous count feedback: after every update launch the new live counts are copied to pinned memory
    // without waiting; when a copy has landed, it tightens the host's upper bounds (which size the next grids).
    struct Feedback
    {
        std::vector<t3d_emitter*> es;
        std::vector<uint64_t> emitted_mark; // emitter's emitted_total when the launch was enqueued
        std::vector<uint64_t> version_mark; // emitter's launch_version right after that launch
        std::vector<uint32_t> box_epoch;    // the epoch this launch's bounding box carries (0: the emitter does not ask for one)
        bool boxes{ false };                // the six box words of every emitter ride behind the counts
        uint32_t* host{ nullptr };
        uint32_t* dev{ nullptr };
        size_t cap{ 0 };
        cudaEvent_t done{ nullptr };
        bool in_flight{ false };
        uint64_t serial{ 0 };               // which read-back of the context this slot holds (0: none yet)
    };
    Feedback feedback[4];
    int feedback_next{ 0 };
    uint64_t feedback_serial{ 0 };          // read-backs submitted so far
    uint64_t pipelined_serial{ 0 };         // the read-back the last t3d_batch_frame_pipelined call left travelling (0: none)

    uint32_t* strip_indices{ nullptr }; // cached MeshBatch index pattern
    uint32_t strip_quads{ 0 };
    uint32_t* tri_indices{ nullptr };   // cached triangle-list pattern
    uint32_t tri_quads{ 0 };

    // Device word the kernels OR into when they find their grid too small for the live count (the host's upper bound was
    // wrong); it rides back with every count read-back and turns into T3D_ERR_INTERNAL on the next call.
    uint32_t* fault_dev{ nullptr };
    uint32_t fault{ 0 };

    // Downloads that overlap the next frame (t3d_ctx_copy_to_host_async): a stream of their own, one event per copy.
    cudaStream_t download_stream{ nullptr };
    cudaEvent_t download_fence{ nullptr };
    std::vector<cudaEvent_t> download_done; // ticket t (1-based) -> download_done[t - 1]
    uint64_t download_waited{ 0 };          // every ticket <= this has been waited for

    // 2-D sprites (t3d_ctx_draw_sprites, t3d_tileset_draw): two pinned/device buffer pairs for lists that arrive from the
    // host (the copy of list k + 1 is filled while list k is still in flight), the context-owned vertex buffer, the
    // per-tile scratch of the ordered compaction
    struct SpriteStage
    {
        char* host{ nullptr };
        char* dev{ nullptr };
        size_t cap{ 0 };
        cudaEvent_t done{ nullptr };
        bool in_flight{ false };
    };
    SpriteStage sprite_stage[2];
    int sprite_stage_next{ 0 };
    float* sprite_vtx{ nullptr };
    size_t sprite_vtx_quads{ 0 };
    uint32_t* sprite_scratch{ nullptr };
    size_t sprite_scratch_words{ 0 };
    std::vector<struct t3d_tileset*> tilesets;
    std::vector<struct t3d_font*> fonts;

    // depth sort (t3d_emitter_depth_sorted_indices): keys / values / library temp / sorted indices, grown on demand
    uint32_t* sort_scratch{ nullptr };
    void* sort_temp{ nullptr };
    uint32_t* sort_indices{ nullptr };
    size_t sort_cap{ 0 }, sort_temp_cap{ 0 };

    uint32_t* readback_host{ nullptr }; // pinned scratch for counts
    size_t readback_cap{ 0 };
    uint32_t* readback_dev{ nullptr };

    std::vector<t3d_emitter*> emitters;
};

struct t3d_emitter
{
    t3d_ctx* ctx{ nullptr };
    t3d_emitter_params p{};
    std::vector<float> tex_coords; // frame_count * 4
    uint32_t frame_count{ 1 };
    float percent_per_frame{ 1.0f };
    float frame_duration_secs{ 0.0f };
    float tex_w_ratio{ 0 }, tex_h_ratio{ 0 };
    float time_per_emission{ 100.0f };
    float emit_time{ 0 };
    bool started{ false };
    bool has_node{ false }, has_camera{ false };
    float node_world[16];
    float camera_world[16];

    int rng_mode{ T3D_RNG_LIBC };
    uint64_t rng_seed{ 0 };
    std::vector<int32_t> fed; // T3D_RNG_STREAM queue
    size_t fed_head{ 0 };

    EmitterDev* dev{ nullptr };
    int slab{ -1 };
    uint32_t seg_offset{ 0 }, seg_capacity{ 0 };
    uint32_t alloc_capacity{ 0 }; // particleCountMax at create: what the storage holds; p.particle_count_max may be set lower
    float* uv_dev{ nullptr };
    uint32_t uv_cap{ 0 };
    uint32_t parity{ 0 };
    uint64_t launch_version{ 0 }; // bumped by every launch that writes this emitter's counts (spawn, update, draw, upload)
    bool const_dirty{ true }, uv_dirty{ true };
    bool may_rotate{ false }, may_spin{ false };
    bool want_bounds{ false }, culling{ false };
    uint32_t bounds_epoch{ 0 }; // launch epoch of the last update that reduced this emitter's box (0: none yet)
    // the last box that has reached the host (count feedback or t3d_emitter_bounds): what culling tests against
    bool box_known{ false }, box_empty{ true };
    float box_lo[3], box_hi[3];
    bool has_vp{ false };
    float view_projection[16];
    uint32_t count_ub{ 0 };     // upper bound of the live count (exact when count_exact)
    bool count_exact{ true };
    bool queued{ false };
    // a queued spawn will read node_world / a queued draw will read camera_world and view_projection when its descriptors
    // are cut: the setters launch it first if they are about to change what it was called with
    bool node_in_use{ false }, camera_in_use{ false };
    bool held{ false }; // has an update in ctx->held (launch deferred, see flush)
    bool last_draw_empty{ true }; // the last draw() submitted no quads (inactive, no particles, or culled)
    float* last_draw_dst{ nullptr }; // where the last draw's quads went when not into the emitter's own stream (t3d_emitter_draw_to)
    uint64_t emitted_total{ 0 };  // particles ever asked to spawn (upper bound of those spawned)
    uint64_t emitted_at_update{ 0 }; // emitted_total when the previous update was queued
    // Milliseconds for which at least one particle is GUARANTEED to be alive: a particle that certainly spawned
    // starts with energy >= floor(min(energyMin, energyMax)) (PE.cpp:316) and every update takes at most
    // elapsed + 1 off it (PE.cpp:753 truncates).  While positive, isActive() of a stopped emitter needs no read-back.
    double alive_budget_ms{ 0 };
    EmitterConst hc{}; // host copy of the device constants as last uploaded (launch descriptors are cut from it)
};

constexpr uint32_t kDevChunk = 4096;

// Adds the wall time of a scope to one slot of the context's host profile.
struct HostTimer
{
    t3d_ctx* c;
    int which;
    double outer_children;
    std::chrono::steady_clock::time_point t0;
    HostTimer(t3d_ctx* ctx, int w) : c(ctx), which(w), outer_children(ctx ? ctx->host_child_s : 0.0), t0(std::chrono::steady_clock::now())
    {
        if (c) c->host_child_s = 0.0;
    }
    ~HostTimer()
    {
        if (!c) return;
        const double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        c->host_s[which] += total - c->host_child_s; // exclusive of the scopes that ran inside this one
        c->host_child_s = outer_children + total;
    }
};

// --------------------------------------------------------------------------------------------------
// context plumbing
// --------------------------------------------------------------------------------------------------
static int use_device(t3d_ctx* c)
{
    T3D_CUDA(cudaSetDevice(c->device));
    return T3D_OK;
}

static int staging_acquire(t3d_ctx* c, size_t bytes, StagingSlot** out)
{
    HostTimer ht(c, T3D_HOST_WAIT); // (waits only when all eight slots are still being read by launches in flight)
    StagingSlot& s = c->staging[c->staging_next];
    c->staging_next = (c->staging_next + 1) % 8;
    if (s.in_flight)
    {
        T3D_CUDA(cudaEventSynchronize(s.done));
        s.in_flight = false;
    }
    if (!s.done) T3D_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    if (!s.uploaded) T3D_CUDA(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
    if (s.cap < bytes)
    {
        // All slots grow together (one pause now instead of one per slot over the next frames): wait for the
        // launches that still read them, then reallocate the ring.
        T3D_CUDA(cudaStreamSynchronize(c->stream));
        if (c->copy_stream) T3D_CUDA(cudaStreamSynchronize(c->copy_stream));
        const size_t cap = std::max<size_t>(bytes + bytes / 2, 1 << 16);
        for (StagingSlot& o : c->staging)
        {
            if (o.host) cudaFreeHost(o.host);
            if (o.dev) cudaFree(o.dev);
            o.host = o.dev = nullptr;
            o.cap = 0;
            o.in_flight = false;
            T3D_CUDA(cudaMallocHost((void**)&o.host, cap));
            T3D_CUDA(cudaMalloc((void**)&o.dev, cap));
            o.cap = cap;
        }
    }
    *out = &s;
    return T3D_OK;
}

static int staging_submit(t3d_ctx* c, StagingSlot* s, size_t bytes)
{
    HostTimer ht(c, T3D_HOST_SUBMIT);
    if (c->copy_stream)
    {
        // The slot's device buffer is free: staging_acquire waited for the last launch that read it.
        T3D_CUDA(cudaMemcpyAsync(s->dev, s->host, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        T3D_CUDA(cudaEventRecord(s->uploaded, c->copy_stream));
        T3D_CUDA(cudaStreamWaitEvent(c->stream, s->uploaded, 0));
    }
    else
        T3D_CUDA(cudaMemcpyAsync(s->dev, s->host, bytes, cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += bytes;
    return T3D_OK;
}

static int staging_release(t3d_ctx* c, StagingSlot* s)
{
    HostTimer ht(c, T3D_HOST_SUBMIT);
    T3D_CUDA(cudaEventRecord(s->done, c->stream));
    s->in_flight = true;
    return T3D_OK;
}

static int alloc_segment(t3d_ctx* c, uint32_t capacity, int* slab_out, uint32_t* offset_out, uint32_t* padded_out)
{
    const uint32_t align = kSegmentAlign;
    uint32_t padded = (std::max(capacity, 1u) + align - 1) / align * align;
    *padded_out = padded;
    // first fit: a range a destroyed emitter gave back, else the untouched tail of a slab
    for (size_t i = 0; i < c->slabs.size(); ++i)
    {
        Slab& s = c->slabs[i];
        if (!s.base) continue;
        for (size_t r = 0; r < s.free_ranges.size(); ++r)
        {
            if (s.free_ranges[r].second < padded) continue;
            *slab_out = (int)i;
            *offset_out = s.free_ranges[r].first;
            s.free_ranges[r].first += padded;
            s.free_ranges[r].second -= padded;
            if (!s.free_ranges[r].second) s.free_ranges.erase(s.free_ranges.begin() + (ptrdiff_t)r);
            s.live_emitters++;
            return T3D_OK;
        }
        if (s.capacity - s.used >= padded)
        {
            *slab_out = (int)i;
            *offset_out = s.used;
            s.used += padded;
            s.live_emitters++;
            return T3D_OK;
        }
    }
    uint64_t want = std::max<uint64_t>(padded, c->next_slab_particles);
    want = (want + 1023) / 1024 * 1024;
    if (want > 0x7fffffffull) return fail(T3D_ERR_CAPACITY, "emitter capacity %u exceeds the 2^31 particle slab limit", capacity);
    Slab s;
    s.capacity = (uint32_t)want;
    T3D_CUDA(cudaMalloc((void**)&s.base, (size_t)kBytesPerParticle * s.capacity));
    s.used = padded;
    s.live_emitters = 1;
    // reuse the table entry of a slab that was given back
    size_t idx = c->slabs.size();
    for (size_t i = 0; i < c->slabs.size(); ++i)
        if (!c->slabs[i].base) idx = i;
    if (idx == c->slabs.size()) c->slabs.push_back(s);
    else c->slabs[idx] = s;
    if (c->next_slab_particles < (1u << 26)) c->next_slab_particles *= 2;
    *slab_out = (int)idx;
    *offset_out = 0;
    return T3D_OK;
}

// Returns an emitter's segment to its slab: merged with neighbouring free ranges; a range that reaches the high-water
// mark lowers it instead.  The last emitter of a slab gives the whole allocation back.
static void free_segment(t3d_ctx* c, int slab, uint32_t offset, uint32_t padded)
{
    Slab& s = c->slabs[slab];
    if (--s.live_emitters == 0)
    {
        if (s.base) cudaFree(s.base);
        if (s.vtx) cudaFree(s.vtx);
        s = Slab();
        return;
    }
    auto& fr = s.free_ranges;
    auto it = std::lower_bound(fr.begin(), fr.end(), std::make_pair(offset, 0u));
    it = fr.insert(it, std::make_pair(offset, padded));
    if (it + 1 != fr.end() && it->first + it->second == (it + 1)->first)
    {
        it->second += (it + 1)->second;
        fr.erase(it + 1);
    }
    if (it != fr.begin() && (it - 1)->first + (it - 1)->second == it->first)
    {
        (it - 1)->second += it->second;
        it = fr.erase(it) - 1;
    }
    if (it->first + it->second == s.used)
    {
        s.used = it->first;
        fr.erase(it);
    }
}

static int alloc_dev_record(t3d_ctx* c, EmitterDev** out)
{
    if (!c->dev_free.empty())
    {
        *out = c->dev_free.back();
        c->dev_free.pop_back();
    }
    else
    {
        if (c->dev_chunks.empty() || c->dev_used == kDevChunk)
        {
            EmitterDev* chunk = nullptr;
            T3D_CUDA(cudaMalloc((void**)&chunk, sizeof(EmitterDev) * kDevChunk));
            c->dev_chunks.push_back(chunk);
            c->dev_used = 0;
        }
        *out = c->dev_chunks.back() + c->dev_used++;
    }
    T3D_CUDA(cudaMemsetAsync(*out, 0, sizeof(EmitterDev), c->stream));
    return T3D_OK;
}

static void host_create_rotation(const float u[3], float angle, float* m)
{
    // Matrix.cpp:348-406 on the host, with the same libm sincosf the reference build calls.
    float x = u[0], y = u[1], z = u[2];
    float n = x * x + y * y + z * z;
    if (n != 1.0f)
    {
        n = sqrtf(n);
        if (n > 0.000001f)
        {
            n = 1.0f / n;
            x *= n;
            y *= n;
            z *= n;
        }
    }
    float s, co;
    sincosf(angle, &s, &co);
    float t = 1.0f - co;
    float tx = t * x, ty = t * y, tz = t * z;
    float txy = tx * y, txz = tx * z, tyz = ty * z;
    float sx = s * x, sy = s * y, sz = s * z;
    m[0] = co + tx * x; m[1] = txy + sz;    m[2] = txz - sy;    m[3] = 0.0f;
    m[4] = txy - sz;    m[5] = co + ty * y; m[6] = tyz + sx;    m[7] = 0.0f;
    m[8] = txz + sy;    m[9] = tyz - sx;    m[10] = co + tz * z; m[11] = 0.0f;
    m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
}

static int upload_frozen(t3d_ctx* c)
{
    FrozenRotation f;
    memset(&f, 0, sizeof(f));
    const ProcessStatics& st = c->statics();
    memcpy(f.m, st.frozen_m, sizeof(f.m));
    f.latched = st.frozen_latched ? 1 : 0;
    T3D_CUDA(cudaMemcpyAsync(c->frozen_dev, &f, sizeof(f), cudaMemcpyHostToDevice, c->stream));
    T3D_CUDA(cudaStreamSynchronize(c->stream)); // `f` is on the stack
    c->frozen_dev_version = st.frozen_version;
    return T3D_OK;
}

// Before a launch that draws: the device copy of the rotation must be the one the statics hold now (another context of
// the process may have latched it, or it may have been reset).
static int ensure_frozen_current(t3d_ctx* c)
{
    return c->frozen_dev_version == c->statics().frozen_version ? T3D_OK : upload_frozen(c);
}

// --------------------------------------------------------------------------------------------------
// per-emitter device constants
// --------------------------------------------------------------------------------------------------
static void fill_const(const t3d_emitter* e, EmitterConst* k)
{
    const t3d_ctx* c = e->ctx;
    const Slab& s = c->slabs[e->slab];
    memset(k, 0, sizeof(*k));
    k->st.rw = s.rw() + e->seg_offset;
    k->st.rec = s.rec() + (size_t)e->seg_offset * REC_WORDS;
    k->st.rot = s.rot() + (size_t)e->seg_offset * ROT_WORDS;
    k->st.stride = s.capacity;
    k->st.seg_capacity = e->seg_capacity;
    k->vtx = s.vtx ? s.vtx + (size_t)e->seg_offset * s.vtx_floats : nullptr;
    k->uv = e->uv_dev;
    k->capacity = e->p.particle_count_max;
    k->flags = (e->p.sprite_animated ? kFlagAnimated : 0u) | (e->p.sprite_looped ? kFlagLooped : 0u) |
               (e->may_rotate ? kFlagMayRotate : 0u) | (e->may_spin ? kFlagMaySpin : 0u) | (e->want_bounds ? kFlagBounds : 0u);
    k->frame_count = e->frame_count;
    k->percent_per_frame = e->percent_per_frame;
    k->frame_duration_secs = e->frame_duration_secs;
    SpawnParams& sp = k->sp;
    const t3d_emitter_params& p = e->p;
    memcpy(sp.color_start, p.color_start, 16); memcpy(sp.color_start_var, p.color_start_var, 16);
    memcpy(sp.color_end, p.color_end, 16); memcpy(sp.color_end_var, p.color_end_var, 16);
    memcpy(sp.position, p.position, 12); memcpy(sp.position_var, p.position_var, 12);
    memcpy(sp.velocity, p.velocity, 12); memcpy(sp.velocity_var, p.velocity_var, 12);
    memcpy(sp.acceleration, p.acceleration, 12); memcpy(sp.acceleration_var, p.acceleration_var, 12);
    memcpy(sp.rotation_axis, p.rotation_axis, 12); memcpy(sp.rotation_axis_var, p.rotation_axis_var, 12);
    sp.size_start_min = p.size_start_min; sp.size_start_max = p.size_start_max;
    sp.size_end_min = p.size_end_min; sp.size_end_max = p.size_end_max;
    sp.energy_min = p.energy_min; sp.energy_max = p.energy_max;
    sp.spin_min = p.rotation_per_particle_speed_min; sp.spin_max = p.rotation_per_particle_speed_max;
    sp.rot_speed_min = p.rotation_speed_min; sp.rot_speed_max = p.rotation_speed_max;
    sp.ellipsoid = p.ellipsoid != 0; sp.orbit_position = p.orbit_position != 0;
    sp.orbit_velocity = p.orbit_velocity != 0; sp.orbit_acceleration = p.orbit_acceleration != 0;
    sp.frame_random_offset = p.sprite_frame_random_offset;
    sp.rng_seed = e->rng_seed;
}

static int sync_const(t3d_emitter* e)
{
    if (!e->uv_dirty && !e->const_dirty) return T3D_OK;
    t3d_ctx* c = e->ctx;
    if (e->uv_dirty)
    {
        const uint32_t need = e->frame_count * 4;
        if (e->uv_cap < need)
        {
            // The old table may still be read by launches in flight on the stream.
            if (e->uv_dev)
            {
                T3D_CUDA(cudaStreamSynchronize(c->stream));
                cudaFree(e->uv_dev);
            }
            T3D_CUDA(cudaMalloc((void**)&e->uv_dev, sizeof(float) * need));
            e->uv_cap = need;
        }
        T3D_CUDA(cudaMemcpyAsync(e->uv_dev, e->tex_coords.data(), sizeof(float) * need, cudaMemcpyHostToDevice, c->stream));
        c->h2d_bytes += sizeof(float) * need;
        e->uv_dirty = false;
        e->const_dirty = true;
    }
    if (e->const_dirty)
    {
        fill_const(e, &e->hc);
        // Pageable source: the runtime stages it before returning.
        T3D_CUDA(cudaMemcpyAsync(&e->dev->c, &e->hc, sizeof(e->hc), cudaMemcpyHostToDevice, c->stream));
        c->h2d_bytes += sizeof(e->hc);
        e->const_dirty = false;
    }
    return T3D_OK;
}

// --------------------------------------------------------------------------------------------------
// flush: turn the queued ops of the current phase into launches
// --------------------------------------------------------------------------------------------------
static int resolve_timing(t3d_ctx* c, int k, bool wait)
{
    auto& v = c->ev_pending[k];
    size_t done = 0;
    for (; done < v.size(); ++done)
    {
        if (wait)
            T3D_CUDA(cudaEventSynchronize(v[done].stop));
        else if (cudaEventQuery(v[done].stop) != cudaSuccess)
            break;
        float ms = 0;
        T3D_CUDA(cudaEventElapsedTime(&ms, v[done].start, v[done].stop));
        c->kernel_ms_total[k] += ms;
        c->kernel_ms_last[k] = ms;
        c->kernel_launches[k]++;
        c->ev_free.push_back(v[done]);
    }
    v.erase(v.begin(), v.begin() + (ptrdiff_t)done);
    return T3D_OK;
}
static int time_begin(t3d_ctx* c, int k)
{
    HostTimer ht(c, T3D_HOST_SUBMIT);
    if (!c->timing_enabled) return T3D_OK;
    if (c->ev_pending[k].size() >= 64) T3D_TRY(resolve_timing(c, k, false)); // recycle finished pairs without waiting
    if (c->ev_pending[k].size() >= 1024) T3D_TRY(resolve_timing(c, k, true));
    t3d_ctx::EventPair p;
    if (!c->ev_free.empty())
    {
        p = c->ev_free.back();
        c->ev_free.pop_back();
    }
    else
    {
        T3D_CUDA(cudaEventCreate(&p.start));
        T3D_CUDA(cudaEventCreate(&p.stop));
    }
    T3D_CUDA(cudaEventRecord(p.start, c->stream));
    c->ev_pending[k].push_back(p);
    return T3D_OK;
}
static int time_end(t3d_ctx* c, int k)
{
    HostTimer ht(c, T3D_HOST_SUBMIT);
    if (!c->timing_enabled) return T3D_OK;
    T3D_CUDA(cudaEventRecord(c->ev_pending[k].back().stop, c->stream));
    return T3D_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int launch_spawns(t3d_ctx* c)
{
    // Gather the ops that spawn.
    uint32_t n_items = 0;
    for (const PendingOp& op : c->pending) n_items += op.emit_n > 0 ? 1u : 0u;
    if (!n_items) return T3D_OK;

    const size_t items_bytes = align_up(sizeof(SpawnItem) * n_items, 16);
    const size_t offs_bytes = align_up(sizeof(uint32_t) * c->pending_offsets.size(), 16);
    const size_t draws_bytes = sizeof(int32_t) * c->pending_draws.size();
    const size_t total = items_bytes + offs_bytes + draws_bytes;
    StagingSlot* slot;
    T3D_TRY(staging_acquire(c, total, &slot));
    SpawnItem* items = reinterpret_cast<SpawnItem*>(slot->host);
    const uint32_t* dev_offs = reinterpret_cast<const uint32_t*>(slot->dev + items_bytes);
    const int32_t* dev_draws = reinterpret_cast<const int32_t*>(slot->dev + items_bytes + offs_bytes);
    if (offs_bytes) memcpy(slot->host + items_bytes, c->pending_offsets.data(), sizeof(uint32_t) * c->pending_offsets.size());
    if (draws_bytes) memcpy(slot->host + items_bytes + offs_bytes, c->pending_draws.data(), draws_bytes);

    uint32_t tiles = 0, k = 0;
    bool any_host_draws = false;
    for (PendingOp& op : c->pending)
    {
        if (!op.emit_n) continue;
        t3d_emitter* e = op.e;
        T3D_TRY(sync_const(e));
        any_host_draws = any_host_draws || op.host_draws;
        SpawnItem& it = items[k++];
        memset(&it, 0, sizeof(it));
        it.dev = e->dev;
        if (op.host_draws)
        {
            it.draws = dev_draws + op.draws_begin;
            it.offsets = op.offsets_begin == SIZE_MAX ? nullptr : dev_offs + op.offsets_begin;
            it.draw_stride = op.draw_stride;
        }
        it.request = op.emit_n;
        it.parity = e->parity;
        e->parity ^= 1u;
        e->launch_version++;
        it.tile_begin = tiles;
        memcpy(it.world, e->node_world, sizeof(it.world)); // PE.cpp:299
        it.n_tiles = (op.emit_n + kSpawnTile - 1) / kSpawnTile;
        tiles += it.n_tiles;
    }
    T3D_TRY(staging_submit(c, slot, total));
    T3D_TRY(time_begin(c, KERNEL_SPAWN));
    launch_spawn(c->stream, reinterpret_cast<const SpawnItem*>(slot->dev), n_items, tiles, any_host_draws);
    T3D_CUDA(cudaGetLastError());
    T3D_TRY(time_end(c, KERNEL_SPAWN));
    c->launches++;
    T3D_TRY(staging_release(c, slot));
    return T3D_OK;
}

// Six words as the update kernel leaves them -> a box; false when no particle contributed in the launch of `epoch`.
static bool decode_box(const unsigned long long* w, uint32_t epoch, float lo[3], float hi[3])
{
    for (int k = 0; k < 3; ++k)
    {
        // a word of an older launch: no particle contributed on this axis in that update, i.e. none is live
        if ((uint32_t)(w[k] >> 32) != epoch || (uint32_t)(w[3 + k] >> 32) != epoch) return false;
        lo[k] = float_unorder(~(uint32_t)w[k]);
        hi[k] = float_unorder((uint32_t)w[3 + k]);
    }
    return true;
}

// A feedback copy has landed: tighten the bounds it speaks for.
static void absorb_feedback(t3d_ctx* c, t3d_ctx::Feedback& fb)
{
    {
        for (size_t i = 0; i < fb.es.size(); ++i)
        {
            t3d_emitter* e = fb.es[i];
            if (!e) continue;
            const uint64_t ub = (uint64_t)fb.host[2 * i] + (e->emitted_total - fb.emitted_mark[i]);
            if (i == 0) c->fault |= fb.host[2 * fb.es.size()];
            if (ub < e->count_ub) e->count_ub = (uint32_t)ub;
            if (fb.boxes && fb.box_epoch[i] && fb.box_epoch[i] == e->bounds_epoch) // (not superseded by a later update's box)
            {
                const unsigned long long* w = reinterpret_cast<const unsigned long long*>(fb.host + 2 * fb.cap + 2) + 6 * i;
                e->box_empty = !decode_box(w, fb.box_epoch[i], e->box_lo, e->box_hi);
                e->box_known = true;
            }
        }
        fb.in_flight = false;
    }
}

static void poll_feedback(t3d_ctx* c)
{
    for (t3d_ctx::Feedback& fb : c->feedback)
        if (fb.in_flight && cudaEventQuery(fb.done) == cudaSuccess) absorb_feedback(c, fb);
}

// Enqueues the read-back of the counts the update launch just produced; never waits.
static int submit_feedback(t3d_ctx* c, const std::vector<PendingOp>& ops, const CountQuery* dev_queries)
{
    t3d_ctx::Feedback& fb = c->feedback[c->feedback_next];
    if (fb.in_flight) return T3D_OK; // all slots busy: skip this frame's feedback
    c->feedback_next = (c->feedback_next + 1) % 4;
    fb.serial = ++c->feedback_serial;
    const size_t n = ops.size();
    // layout: counts (2 words per emitter) + the fault word, padded to `cap` emitters; then 6 x 64-bit box words per emitter
    auto words = [](size_t cap) { return 2 * cap + 2 + 12 * cap; };
    if (fb.cap < n)
    {
        if (fb.host) cudaFreeHost(fb.host);
        if (fb.dev) cudaFree(fb.dev);
        fb.host = fb.dev = nullptr;
        const size_t cap = std::max<size_t>(n * 2, 64);
        T3D_CUDA(cudaMallocHost((void**)&fb.host, words(cap) * sizeof(uint32_t)));
        T3D_CUDA(cudaMalloc((void**)&fb.dev, words(cap) * sizeof(uint32_t)));
        fb.cap = cap;
    }
    if (!fb.done) T3D_CUDA(cudaEventCreateWithFlags(&fb.done, cudaEventDisableTiming));
    fb.es.resize(n);
    fb.emitted_mark.resize(n);
    fb.version_mark.resize(n);
    fb.box_epoch.assign(n, 0);
    fb.boxes = false;
    for (size_t i = 0; i < n; ++i)
    {
        fb.es[i] = ops[i].e;
        fb.emitted_mark[i] = ops[i].e->emitted_total;
        fb.version_mark[i] = ops[i].e->launch_version;
        if (ops[i].e->want_bounds)
        {
            fb.box_epoch[i] = ops[i].e->bounds_epoch;
            fb.boxes = true;
        }
    }
    unsigned long long* dev_boxes = fb.boxes ? reinterpret_cast<unsigned long long*>(fb.dev + 2 * fb.cap + 2) : nullptr;
    launch_gather_counts(c->stream, dev_queries, (uint32_t)n, fb.dev, c->fault_dev, dev_boxes);
    T3D_CUDA(cudaGetLastError());
    c->launches++;
    T3D_CUDA(cudaMemcpyAsync(fb.host, fb.dev, (n * 2 + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += n * 2 * sizeof(uint32_t);
    if (fb.boxes)
    {
        T3D_CUDA(cudaMemcpyAsync(fb.host + 2 * fb.cap + 2, dev_boxes, n * 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        c->d2h_bytes += n * 6 * sizeof(unsigned long long);
    }
    T3D_CUDA(cudaEventRecord(fb.done, c->stream));
    fb.in_flight = true;
    return T3D_OK;
}

// T3D_DEBUG_SHRINK_GRID=1 (tests only): every update launch leaves the last tile of each emitter out.
static bool debug_shrink_grid()
{
    static int v = -1;
    if (v < 0)
    {
        const char* env = getenv("T3D_DEBUG_SHRINK_GRID");
        v = (env && env[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// Half-size tiles for a launch that is only a few waves wide (see pick_tile_shape).
static int choose_shape(const t3d_ctx* c, const std::vector<PendingOp>& ops, uint32_t tile_big, uint32_t tile_small)
{
    if (tile_small >= tile_big) return 0;
    uint64_t big = 0, small_ = 0;
    for (const PendingOp& op : ops)
    {
        big += std::max(1u, (op.e->count_ub + tile_big - 1) / tile_big);
        small_ += std::max(1u, (op.e->count_ub + tile_small - 1) / tile_small);
    }
    return pick_tile_shape(big, small_, c->cta_slots);
}

// Cuts the update descriptors (and the count queries of the feedback) of `ops`; advances each emitter's parity.
static void build_update_items(std::vector<PendingOp>& ops, UpdateItem* items, CountQuery* queries, uint32_t tile_size, uint32_t* tiles_out)
{
    uint32_t tiles = 0;
    for (size_t k = 0; k < ops.size(); ++k)
    {
        PendingOp& op = ops[k];
        t3d_emitter* e = op.e;
        UpdateItem& it = items[k];
        const EmitterConst& ec = e->hc; // current: sync_const ran just before
        it.dev = e->dev;
        it.st = ec.st;
        it.flags = ec.flags;
        it.frame_count = ec.frame_count;
        it.percent_per_frame = ec.percent_per_frame;
        it.frame_duration_secs = ec.frame_duration_secs;
        it.elapsed_ms = op.elapsed_ms;
        it.parity = e->parity;
        e->parity ^= 1u;
        e->launch_version++;
        it.tile_begin = tiles;
        queries[k].dev = e->dev;
        queries[k].parity = e->parity; // where this launch leaves the new count
        queries[k].pad = 0;
        // At least one tile per emitter: tile 0 republishes the count into the other parity slot.
        it.n_tiles = std::max(1u, (e->count_ub + tile_size - 1) / tile_size);
        if (debug_shrink_grid() && it.n_tiles > 1) it.n_tiles -= 1; // test hook: proves that the kernel's check fires
        tiles += it.n_tiles;
    }
    *tiles_out = tiles;
}

static int begin_scan_epoch(t3d_ctx* c, uint32_t tiles)
{
    if (c->tile_status_cap < tiles)
    {
        if (c->tile_status)
        {
            T3D_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->tile_status);
        }
        size_t cap = std::max<size_t>(tiles + tiles / 2, 4096);
        T3D_CUDA(cudaMalloc((void**)&c->tile_status, cap * sizeof(unsigned long long)));
        T3D_CUDA(cudaMemsetAsync(c->tile_status, 0, cap * sizeof(unsigned long long), c->stream));
        c->tile_status_cap = cap;
    }
    c->epoch++;
    if (c->epoch >= (1u << 30))
    {
        T3D_CUDA(cudaMemsetAsync(c->tile_status, 0, c->tile_status_cap * sizeof(unsigned long long), c->stream));
        c->epoch = 1;
    }
    return T3D_OK;
}

// Emitters of the launch that want their bounding box: remembers the epoch their box will carry.
static bool mark_bounds(t3d_ctx* c, std::vector<PendingOp>& ops)
{
    bool any = false;
    for (PendingOp& op : ops)
        if (op.e->want_bounds)
        {
            op.e->bounds_epoch = c->epoch;
            any = true;
        }
    return any;
}

static int launch_updates(t3d_ctx* c, std::vector<PendingOp>& ops)
{
    const uint32_t n_items = (uint32_t)ops.size();
    if (!n_items) return T3D_OK;
    for (PendingOp& op : ops) T3D_TRY(sync_const(op.e));
    StagingSlot* slot;
    const size_t items_bytes = align_up(sizeof(UpdateItem) * n_items, 16);
    const size_t total_bytes = items_bytes + sizeof(CountQuery) * n_items;
    T3D_TRY(staging_acquire(c, total_bytes, &slot));
    uint32_t tiles = 0;
    const int shape = choose_shape(c, ops, (uint32_t)update_tile_size(0), (uint32_t)update_tile_size(1));
    build_update_items(ops, reinterpret_cast<UpdateItem*>(slot->host), reinterpret_cast<CountQuery*>(slot->host + items_bytes),
                       (uint32_t)update_tile_size(shape), &tiles);
    T3D_TRY(begin_scan_epoch(c, tiles));
    T3D_TRY(staging_submit(c, slot, total_bytes));
    T3D_TRY(time_begin(c, KERNEL_UPDATE));
    UpdateLaunch L{};
    L.items = reinterpret_cast<const UpdateItem*>(slot->dev);
    L.n_items = n_items;
    L.total_tiles = tiles;
    L.tile_status = c->tile_status;
    L.ticket = c->ticket;
    L.ticket_base = c->ticket_base;
    L.epoch = c->epoch;
    L.fault = c->fault_dev;
    L.bounds = mark_bounds(c, ops);
    launch_update(c->stream, L, shape);
    T3D_CUDA(cudaGetLastError());
    T3D_TRY(time_end(c, KERNEL_UPDATE));
    c->ticket_base += tiles;
    c->launches++;
    T3D_TRY(submit_feedback(c, ops, reinterpret_cast<const CountQuery*>(slot->dev + items_bytes)));
    T3D_TRY(staging_release(c, slot));
    return T3D_OK;
}

// Vertex storage is allocated per slab the first time one of its emitters is drawn into it (draws into a caller's
// buffer need none); a later switch to the wider clip-space vertex reallocates it.
static int ensure_vertex_storage(t3d_ctx* c, std::vector<PendingOp>& ops)
{
    const uint32_t want = c->vertex_mode == T3D_VERTEX_CLIP ? T3D_FLOATS_PER_CLIP_QUAD : T3D_FLOATS_PER_QUAD;
    for (PendingOp& op : ops)
    {
        if (op.draw_dst) continue;
        Slab& s = c->slabs[op.e->slab];
        if (s.vtx && s.vtx_floats >= want) continue;
        if (s.vtx)
        {
            T3D_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(s.vtx);
            s.vtx = nullptr;
        }
        T3D_CUDA(cudaMalloc((void**)&s.vtx, (size_t)s.capacity * want * sizeof(float)));
        s.vtx_floats = want;
        for (t3d_emitter* o : c->emitters)
            if (o->slab == op.e->slab) o->const_dirty = true;
    }
    for (PendingOp& op : ops) T3D_TRY(sync_const(op.e));
    return T3D_OK;
}

static void build_draw_items(std::vector<PendingOp>& ops, DrawItem* items, uint32_t draw_tile, uint32_t* tiles_out, bool* any_spin_out,
                             bool clip)
{
    uint32_t tiles = 0;
    bool any_spin = false;
    for (size_t k = 0; k < ops.size(); ++k)
    {
        PendingOp& op = ops[k];
        t3d_emitter* e = op.e;
        DrawItem& it = items[k];
        const EmitterConst& ec = e->hc; // current: sync_const ran just before
        it.dev = e->dev;
        it.rw = ec.st.rw;
        it.vtx = op.draw_dst ? op.draw_dst : ec.vtx;
        it.uv = ec.uv;
        it.stride = ec.st.stride;
        it.seg_capacity = ec.st.seg_capacity;
        it.frame_count = ec.frame_count;
        if (clip) memcpy(it.vp, e->view_projection, sizeof(it.vp)); // PE.cpp:846-849 (only the clip-space draw reads it)
        e->launch_version++; // (a draw rewrites the emitter's quad count)
        it.parity = e->parity;
        it.tile_begin = tiles;
        // PE.cpp:860-864: right = (m0, m1, m2), up = (m4, m5, m6) of the camera node's world matrix.
        memcpy(it.right, e->camera_world, 12);
        memcpy(it.up, e->camera_world + 4, 12);
        it.n_tiles = std::max(1u, (e->count_ub + draw_tile - 1) / draw_tile);
        tiles += it.n_tiles;
        any_spin = any_spin || e->may_spin;
    }
    *tiles_out = tiles;
    *any_spin_out = any_spin;
}

static int launch_draws(t3d_ctx* c)
{
    const uint32_t n_items = (uint32_t)c->pending.size();
    T3D_TRY(ensure_vertex_storage(c, c->pending));
    T3D_TRY(ensure_frozen_current(c));
    StagingSlot* slot;
    T3D_TRY(staging_acquire(c, sizeof(DrawItem) * n_items, &slot));
    uint32_t tiles = 0;
    bool any_spin = false;
    build_draw_items(c->pending, reinterpret_cast<DrawItem*>(slot->host), (uint32_t)draw_tile_size(n_items), &tiles, &any_spin,
                     c->vertex_mode == T3D_VERTEX_CLIP);
    T3D_TRY(staging_submit(c, slot, sizeof(DrawItem) * n_items));
    if (c->rot_mode == T3D_ROT_FROZEN && !c->statics().frozen_latched && any_spin)
    {
        // SB.cpp:337-339: the first rotated quad of the process fixes the rotation matrix for good.  Find it on
        // the device, build the matrix on the host with the libm the reference uses, upload it.  Happens once.
        launch_latch(c->stream, reinterpret_cast<const DrawItem*>(slot->dev), n_items, c->frozen_dev);
        T3D_CUDA(cudaGetLastError());
        c->launches++;
        FrozenRotation f;
        T3D_CUDA(cudaMemcpyAsync(&f, c->frozen_dev, sizeof(f), cudaMemcpyDeviceToHost, c->stream));
        T3D_CUDA(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += sizeof(f);
        if (f.latched == 2)
        {
            const float u[3] = { f.m[0], f.m[1], f.m[2] };
            ProcessStatics& st = c->statics();
            host_create_rotation(u, f.m[3], st.frozen_m);
            st.frozen_latched = true;
            st.frozen_version++;
            T3D_TRY(upload_frozen(c));
        }
    }
    T3D_TRY(time_begin(c, KERNEL_DRAW));
    launch_draw(c->stream, reinterpret_cast<const DrawItem*>(slot->dev), n_items, tiles, c->rot_mode, c->frozen_dev, c->fault_dev,
                c->vertex_mode);
    T3D_CUDA(cudaGetLastError());
    T3D_TRY(time_end(c, KERNEL_DRAW));
    c->launches++;
    T3D_TRY(staging_release(c, slot));
    return T3D_OK;
}

// The held updates and the pending draws cover the same emitters in the same order, and no rotation matrix has to
// be latched between them: one launch does both (k_update<..., FUSED>).
static bool can_fuse(const t3d_ctx* c)
{
    if (c->held.empty() || c->held.size() != c->pending.size()) return false;
    if (c->vertex_mode != T3D_VERTEX_WORLD) return false; // the fused pass writes SpriteVertex
    bool any_spin = false;
    for (size_t k = 0; k < c->held.size(); ++k)
    {
        if (c->held[k].e != c->pending[k].e) return false;
        if (c->held[k].e->frame_count > (uint32_t)kMaxFusedFrames) return false; // the fused pass reads uv from shared memory only
        any_spin = any_spin || c->held[k].e->may_spin;
    }
    if (c->rot_mode == T3D_ROT_FROZEN && !c->statics().frozen_latched && any_spin) return false;
    return true;
}

static int launch_fused_frame(t3d_ctx* c)
{
    const uint32_t n_items = (uint32_t)c->held.size();
    T3D_TRY(ensure_vertex_storage(c, c->pending));
    T3D_TRY(ensure_frozen_current(c));
    StagingSlot* slot;
    const size_t items_bytes = align_up(sizeof(UpdateItem) * n_items, 16);
    const size_t queries_bytes = align_up(sizeof(CountQuery) * n_items, 16);
    const size_t total_bytes = items_bytes + queries_bytes + sizeof(DrawItem) * n_items;
    T3D_TRY(staging_acquire(c, total_bytes, &slot));
    uint32_t tiles = 0, draw_tiles = 0;
    bool any_spin = false;
    const int shape = choose_shape(c, c->held, (uint32_t)fused_tile_size(0), (uint32_t)fused_tile_size(1));
    const uint32_t tile_size = (uint32_t)fused_tile_size(shape);
    build_update_items(c->held, reinterpret_cast<UpdateItem*>(slot->host), reinterpret_cast<CountQuery*>(slot->host + items_bytes),
                       tile_size, &tiles);
    build_draw_items(c->pending, reinterpret_cast<DrawItem*>(slot->host + items_bytes + queries_bytes), tile_size, &draw_tiles, &any_spin, false);
    T3D_TRY(begin_scan_epoch(c, tiles));
    T3D_TRY(staging_submit(c, slot, total_bytes));
    T3D_TRY(time_begin(c, KERNEL_UPDATE_DRAW));
    UpdateLaunch L{};
    L.items = reinterpret_cast<const UpdateItem*>(slot->dev);
    L.ditems = reinterpret_cast<const DrawItem*>(slot->dev + items_bytes + queries_bytes);
    L.n_items = n_items;
    L.total_tiles = tiles;
    L.tile_status = c->tile_status;
    L.ticket = c->ticket;
    L.ticket_base = c->ticket_base;
    L.epoch = c->epoch;
    L.fault = c->fault_dev;
    L.rot_mode = c->rot_mode;
    L.frozen = c->frozen_dev;
    L.bounds = mark_bounds(c, c->held);
    launch_update_draw(c->stream, L, shape);
    T3D_CUDA(cudaGetLastError());
    c->ticket_base += tiles;
    T3D_TRY(time_end(c, KERNEL_UPDATE_DRAW));
    c->launches++;
    T3D_TRY(submit_feedback(c, c->held, reinterpret_cast<const CountQuery*>(slot->dev + items_bytes)));
    T3D_TRY(staging_release(c, slot));
    return T3D_OK;
}

// Should the pending updates wait for the draws that are about to be queued?  Automatic mode holds them while the
// churn is low.  A slot that receives a moved particle costs the fused pass an extra scattered 144-byte quad on top of the
// move, and its tiles drain their bulk stores before the move pass; measured on B200 with 38 M live particles
// (profiles/r1_churn_crossover.txt) the fused frame is 7.7 % faster with no particle replaced per frame, 2.9 % faster at
// 0.17 %, 1.8 % slower at 0.34 %, 8-12 % slower at 0.7-2.7 %.  Deaths are not known on the host; the particles spawned
// since the previous update stand in for them (equal in steady state).
static bool hold_for_fusion(const t3d_ctx* c)
{
    const int mode = fused_mode();
    if (mode >= 0) return mode == 1;
    // (the policy no longer looks at its arguments -- the fused frame wins at every churn level measured -- so the
    //  population is not summed up here; see t3d_host_fuse_policy)
    (void)c;
    return t3d_host_fuse_policy(0, 0) != 0;
}

static int check_fault(const t3d_ctx* c)
{
    if (c->fault)
        return fail(T3D_ERR_INTERNAL, "a kernel found its grid too small for the live count (%s%s): the host's upper bound was wrong; "
                                      "particles were left out of a frame", (c->fault & kFaultUpdateGrid) ? "update " : "",
                    (c->fault & kFaultDrawGrid) ? "draw" : "");
    return T3D_OK;
}

static int launch_held(t3d_ctx* c)
{
    if (c->held.empty()) return T3D_OK;
    const int rc = launch_updates(c, c->held);
    for (PendingOp& op : c->held) op.e->held = false;
    c->held.clear();
    return rc;
}

// Launches what is queued.  next_phase is the phase the caller is about to open: updates directly followed by draws
// are held back one step so that both can go out as one fused launch when they cover the same emitters.
static int flush(t3d_ctx* c, int next_phase = PHASE_NONE)
{
    HostTimer ht(c, T3D_HOST_BUILD);
    if (c->phase == PHASE_NONE || c->pending.empty())
    {
        c->phase = PHASE_NONE;
        if (next_phase != PHASE_DRAW && !c->held.empty())
        {
            T3D_TRY(use_device(c));
            return launch_held(c);
        }
        return T3D_OK;
    }
    T3D_TRY(use_device(c));
    poll_feedback(c);
    T3D_TRY(check_fault(c));
    int rc = T3D_OK;
    bool hold = false;
    if (c->phase == PHASE_EMIT)
    {
        rc = launch_held(c);
        if (rc == T3D_OK) rc = launch_spawns(c);
    }
    else if (c->phase == PHASE_UPDATE)
    {
        rc = launch_held(c);
        if (rc == T3D_OK) rc = launch_spawns(c);
        hold = rc == T3D_OK && next_phase == PHASE_DRAW && hold_for_fusion(c);
        if (rc == T3D_OK && !hold) rc = launch_updates(c, c->pending);
    }
    else if (c->phase == PHASE_DRAW)
    {
        if (can_fuse(c))
        {
            rc = launch_fused_frame(c);
            for (PendingOp& op : c->held) op.e->held = false;
            c->held.clear();
        }
        else
        {
            rc = launch_held(c);
            if (rc == T3D_OK) rc = launch_draws(c);
        }
    }
    for (PendingOp& op : c->pending)
    {
        op.e->queued = false;
        op.e->node_in_use = op.e->camera_in_use = false; // spawns and draws of this phase have been cut and launched
    }
    if (hold)
    {
        c->held.swap(c->pending);
        for (PendingOp& op : c->held) op.e->held = true;
    }
    c->pending.clear();
    c->pending_draws.clear();
    c->pending_offsets.clear();
    c->phase = PHASE_NONE;
    return rc;
}

static int enqueue(t3d_ctx* c, int phase, const PendingOp& op)
{
    if (c->phase != phase || op.e->queued) T3D_TRY(flush(c, phase));
    c->phase = phase;
    c->pending.push_back(op);
    op.e->queued = true;
    if (op.emit_n) op.e->node_in_use = true;
    if (phase == PHASE_DRAW) op.e->camera_in_use = true;
    return T3D_OK;
}

// Reads the live count (and the quad count of the last draw) back; makes the host mirror exact.
static int refresh_counts(t3d_ctx* c, t3d_emitter* const* es, size_t n, uint32_t* counts, uint32_t* drawn)
{
    T3D_TRY(flush(c));
    T3D_TRY(use_device(c));
    if (n == 0) return T3D_OK;
    {
        // The counts of the newest launch are already on their way back (submit_feedback): when nothing has touched these
        // emitters since and the question is about exactly that launch's emitters, wait for that copy instead of asking again.
        t3d_ctx::Feedback& fb = c->feedback[(c->feedback_next + 3) % 4];
        bool same = fb.in_flight && fb.es.size() == n;
        for (size_t i = 0; same && i < n; ++i) same = fb.es[i] == es[i] && fb.version_mark[i] == es[i]->launch_version;
        if (same)
        {
            {
                HostTimer ht(c, T3D_HOST_WAIT);
                T3D_CUDA(cudaEventSynchronize(fb.done));
            }
            poll_feedback(c); // older copies first (they can only loosen nothing), then this one
            if (fb.in_flight) absorb_feedback(c, fb);
            c->fault |= fb.host[2 * n];
            T3D_TRY(check_fault(c));
            for (size_t i = 0; i < n; ++i)
            {
                es[i]->count_ub = fb.host[2 * i];
                es[i]->count_exact = true;
                if (counts) counts[i] = fb.host[2 * i];
                if (drawn) drawn[i] = fb.host[2 * i + 1];
            }
            return T3D_OK;
        }
    }
    if (c->readback_cap < n)
    {
        if (c->readback_host) cudaFreeHost(c->readback_host);
        if (c->readback_dev) cudaFree(c->readback_dev);
        size_t cap = std::max<size_t>(n * 2, 256);
        T3D_CUDA(cudaMallocHost((void**)&c->readback_host, (cap * 2 + 1) * sizeof(uint32_t)));
        T3D_CUDA(cudaMalloc((void**)&c->readback_dev, (cap * 2 + 1) * sizeof(uint32_t)));
        c->readback_cap = cap;
    }
    StagingSlot* slot;
    T3D_TRY(staging_acquire(c, sizeof(CountQuery) * n, &slot));
    CountQuery* q = reinterpret_cast<CountQuery*>(slot->host);
    for (size_t i = 0; i < n; ++i)
    {
        q[i].dev = es[i]->dev;
        q[i].parity = es[i]->parity;
        q[i].pad = 0;
    }
    T3D_TRY(staging_submit(c, slot, sizeof(CountQuery) * n));
    launch_gather_counts(c->stream, reinterpret_cast<const CountQuery*>(slot->dev), (uint32_t)n, c->readback_dev, c->fault_dev, nullptr);
    T3D_CUDA(cudaGetLastError());
    c->launches++;
    T3D_CUDA(cudaMemcpyAsync(c->readback_host, c->readback_dev, (n * 2 + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    T3D_TRY(staging_release(c, slot));
    {
        HostTimer ht(c, T3D_HOST_WAIT);
        T3D_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->d2h_bytes += n * 2 * sizeof(uint32_t);
    c->fault |= c->readback_host[2 * n];
    T3D_TRY(check_fault(c));
    for (size_t i = 0; i < n; ++i)
    {
        es[i]->count_ub = c->readback_host[2 * i];
        es[i]->count_exact = true;
        if (counts) counts[i] = c->readback_host[2 * i];
        if (drawn) drawn[i] = c->readback_host[2 * i + 1];
    }
    return T3D_OK;
}

static int exact_count(t3d_emitter* e, uint32_t* out)
{
    if (!e->count_exact)
    {
        // One read-back serves every emitter whose count is stale: a loop over many stopped-but-live emitters
        // (isActive needs `count > 0`, PE.cpp:283) then costs one sync per phase, not one per emitter.
        std::vector<t3d_emitter*> stale;
        for (t3d_emitter* o : e->ctx->emitters)
            if (!o->count_exact) stale.push_back(o);
        T3D_TRY(refresh_counts(e->ctx, stale.data(), stale.size(), nullptr, nullptr));
    }
    *out = e->count_ub;
    return T3D_OK;
}

// --------------------------------------------------------------------------------------------------
// host-side random draws for T3D_RNG_LIBC / T3D_RNG_STREAM
// --------------------------------------------------------------------------------------------------
namespace
{
struct HostDraws
{
    t3d_emitter* e;
    bool underflow{ false };
    int32_t next()
    {
        if (e->rng_mode == T3D_RNG_LIBC) return (int32_t)rand();
        if (e->fed_head >= e->fed.size())
        {
            underflow = true;
            return 0;
        }
        return e->fed[e->fed_head++];
    }
};
} // namespace

// Pulls the draws n particles consume (PE.cpp:312-364 order) into the context's pending buffers.
static int collect_host_draws(t3d_emitter* e, uint32_t n, PendingOp* op)
{
    t3d_ctx* c = e->ctx;
    const bool ellipsoid = e->p.ellipsoid != 0;
    const uint32_t fro = e->p.sprite_frame_random_offset;
    const uint32_t stride = 26u + (fro > 0 ? 1u : 0u);
    HostDraws src{ e };
    const size_t fed_head0 = e->fed_head;
    const size_t d0 = c->pending_draws.size(), o0 = c->pending_offsets.size();
    op->host_draws = true;
    op->draws_begin = d0;
    op->draw_stride = stride;
    op->offsets_begin = ellipsoid ? o0 : SIZE_MAX;
    std::vector<int32_t>& out = c->pending_draws;
    if (!ellipsoid)
    {
        out.resize(d0 + (size_t)n * stride);
        for (size_t i = d0; i < out.size(); ++i) out[i] = src.next();
    }
    else
    {
        for (uint32_t i = 0; i < n; ++i)
        {
            c->pending_offsets.push_back((uint32_t)(out.size() - d0));
            for (int k = 0; k < 14; ++k) out.push_back(src.next());
            while (true)
            {
                const int32_t rx = src.next(), ry = src.next(), rz = src.next();
                out.push_back(rx);
                out.push_back(ry);
                out.push_back(rz);
                if (src.underflow || t3d_host_ellipsoid_accept(rx, ry, rz)) break; // PE.cpp:636-641
            }
            for (uint32_t k = 0; k < 9u + (fro > 0 ? 1u : 0u); ++k) out.push_back(src.next());
        }
    }
    if (src.underflow)
    {
        out.resize(d0);
        c->pending_offsets.resize(o0);
        e->fed_head = fed_head0;
        return fail(T3D_ERR_RNG_UNDERFLOW, "T3D_RNG_STREAM: %u particles need more draws than were fed", n);
    }
    op->draws_count = out.size() - d0;
    // Drop consumed draws from the front of the queue now and then.
    if (e->fed_head > (1u << 20) && e->fed_head * 2 > e->fed.size())
    {
        e->fed.erase(e->fed.begin(), e->fed.begin() + (ptrdiff_t)e->fed_head);
        e->fed_head = 0;
    }
    return T3D_OK;
}

// Common part of emitOnce, direct or from update (PE.cpp:287-306): clamp, capture the node matrix, source the draws.
static int prepare_emit(t3d_emitter* e, uint32_t n, PendingOp* op)
{
    if (!e->has_node) return fail(T3D_ERR_NO_NODE, "emitOnce: the emitter has no node world matrix (PE.cpp:289)");
    const uint32_t cap = e->p.particle_count_max;
    // Particles spawned from now on carry the current rotation settings.
    if (e->p.rotation_speed_min != 0.0f || e->p.rotation_speed_max != 0.0f)
    {
        if (!e->may_rotate) e->const_dirty = true;
        e->may_rotate = true;
    }
    if (e->p.rotation_per_particle_speed_min != 0.0f || e->p.rotation_per_particle_speed_max != 0.0f)
    {
        if (!e->may_spin) e->const_dirty = true;
        e->may_spin = true;
    }
    op->host_draws = false;
    op->draws_begin = op->draws_count = 0;
    op->offsets_begin = SIZE_MAX;
    op->draw_stride = 0;
    if (e->rng_mode == T3D_RNG_DEVICE)
    {
        // The kernel clamps against the live count it reads (PE.cpp:293-296); the host keeps an upper bound.
        op->emit_n = n;
        e->emitted_total += n;
        const uint64_t ub = (uint64_t)e->count_ub + n;
        if (n && ub <= cap) // all of them fit for sure
            e->alive_budget_ms = std::max(e->alive_budget_ms, (double)std::floor(std::min(e->p.energy_min, e->p.energy_max)));
        // (a cap lowered below the live count spawns nothing -- the reference's unsigned `max - count` would overrun its
        // array there -- and must not pull the bound under the particles that are already alive)
        e->count_ub = std::max(e->count_ub, (uint32_t)std::min<uint64_t>(ub, cap));
        return T3D_OK;
    }
    // Host draws: the reference consumes draws only for particles that fit, so the clamp must be exact here.
    if ((uint64_t)e->count_ub + n > cap)
    {
        uint32_t cnt;
        T3D_TRY(exact_count(e, &cnt));
        if ((uint64_t)cnt + n > cap) n = cap > cnt ? cap - cnt : 0;
    }
    op->emit_n = n;
    if (n)
    {
        T3D_TRY(collect_host_draws(e, n, op));
        e->count_ub += n;
        e->emitted_total += n;
        e->alive_budget_ms = std::max(e->alive_budget_ms, (double)std::floor(std::min(e->p.energy_min, e->p.energy_max)));
    }
    return T3D_OK;
}

// --------------------------------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------------------------------
extern "C" {

int t3d_abi_version(void) { return T3D_ABI_VERSION; }
const char* t3d_last_error(void) { return g_last_error.c_str(); }

int t3d_ctx_create(int device, void* cuda_stream, t3d_ctx** out)
{
    if (!out) return fail(T3D_ERR_INVALID, "t3d_ctx_create: out is null");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t err = cudaGetDeviceCount(&n_dev);
    if (err != cudaSuccess || n_dev == 0)
        return fail(T3D_ERR_CUDA, "no CUDA device available (%s); this backend has no CPU fallback",
                    err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
    if (device < 0 || device >= n_dev) return fail(T3D_ERR_INVALID, "t3d_ctx_create: device %d out of range (0..%d)", device, n_dev - 1);
    T3D_CUDA(cudaSetDevice(device));
    t3d_ctx* c = new t3d_ctx();
    c->device = device;
    if (cuda_stream)
        c->stream = (cudaStream_t)cuda_stream;
    else
    {
        cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e2 != cudaSuccess)
        {
            delete c;
            return fail(T3D_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e2));
        }
        c->own_stream = true;
    }
    if (cudaMalloc((void**)&c->ticket, sizeof(uint32_t)) != cudaSuccess ||
        cudaMemsetAsync(c->ticket, 0, sizeof(uint32_t), c->stream) != cudaSuccess ||
        cudaMalloc((void**)&c->frozen_dev, sizeof(FrozenRotation)) != cudaSuccess ||
        cudaMalloc((void**)&c->fault_dev, sizeof(uint32_t)) != cudaSuccess ||
        cudaMemsetAsync(c->fault_dev, 0, sizeof(uint32_t), c->stream) != cudaSuccess ||
        cudaMemsetAsync(c->frozen_dev, 0, sizeof(FrozenRotation), c->stream) != cudaSuccess)
    {
        delete c;
        return fail(T3D_ERR_CUDA, "t3d_ctx_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) c->cta_slots = 2u * (uint32_t)sms;
    }
    {
        const char* env = getenv("T3D_COPY_STREAM");
        if (!(env && env[0] == '0') && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
            c->copy_stream = nullptr; // descriptors then travel on the launching stream
    }
    g_contexts.push_back(c);
    *out = c;
    return T3D_OK;
}

int t3d_ctx_destroy(t3d_ctx* c)
{
    if (!c) return T3D_OK;
    g_contexts.erase(std::remove(g_contexts.begin(), g_contexts.end(), c), g_contexts.end());
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    std::vector<t3d_emitter*> es = c->emitters;
    for (t3d_emitter* e : es) t3d_emitter_destroy(e);
    for (Slab& s : c->slabs)
    {
        if (s.base) cudaFree(s.base);
        if (s.vtx) cudaFree(s.vtx);
    }
    if (c->download_stream) cudaStreamDestroy(c->download_stream);
    if (c->download_fence) cudaEventDestroy(c->download_fence);
    for (cudaEvent_t ev : c->download_done)
        if (ev) cudaEventDestroy(ev);
    for (EmitterDev* d : c->dev_chunks) cudaFree(d);
    for (StagingSlot& s : c->staging)
    {
        if (s.host) cudaFreeHost(s.host);
        if (s.dev) cudaFree(s.dev);
        if (s.done) cudaEventDestroy(s.done);
        if (s.uploaded) cudaEventDestroy(s.uploaded);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->tile_status) cudaFree(c->tile_status);
    if (c->ticket) cudaFree(c->ticket);
    if (c->frozen_dev) cudaFree(c->frozen_dev);
    if (c->fault_dev) cudaFree(c->fault_dev);
    if (c->readback_host) cudaFreeHost(c->readback_host);
    if (c->readback_dev) cudaFree(c->readback_dev);
    if (c->strip_indices) cudaFree(c->strip_indices);
    if (c->tri_indices) cudaFree(c->tri_indices);
    {
        std::vector<t3d_tileset*> ts = c->tilesets;
        for (t3d_tileset* t : ts) t3d_tileset_destroy(t);
        std::vector<t3d_font*> fs = c->fonts;
        for (t3d_font* f : fs) t3d_font_destroy(f);
    }
    for (t3d_ctx::SpriteStage& st : c->sprite_stage)
    {
        if (st.host) cudaFreeHost(st.host);
        if (st.dev) cudaFree(st.dev);
        if (st.done) cudaEventDestroy(st.done);
    }
    if (c->sort_scratch) cudaFree(c->sort_scratch);
    if (c->sort_temp) cudaFree(c->sort_temp);
    if (c->sort_indices) cudaFree(c->sort_indices);
    if (c->sprite_vtx) cudaFree(c->sprite_vtx);
    if (c->sprite_scratch) cudaFree(c->sprite_scratch);
    for (t3d_ctx::Feedback& fb : c->feedback)
    {
        if (fb.host) cudaFreeHost(fb.host);
        if (fb.dev) cudaFree(fb.dev);
        if (fb.done) cudaEventDestroy(fb.done);
    }
    for (int k = 0; k < KERNEL_KINDS; ++k)
        for (auto& p : c->ev_pending[k]) c->ev_free.push_back(p);
    for (auto& p : c->ev_free)
    {
        cudaEventDestroy(p.start);
        cudaEventDestroy(p.stop);
    }
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return T3D_OK;
}

int t3d_ctx_flush(t3d_ctx* c)
{
    if (!c) return fail(T3D_ERR_INVALID, "null context");
    return flush(c);
}

int t3d_ctx_sync(t3d_ctx* c)
{
    if (!c) return fail(T3D_ERR_INVALID, "null context");
    T3D_TRY(flush(c));
    T3D_TRY(use_device(c));
    HostTimer ht(c, T3D_HOST_WAIT);
    T3D_CUDA(cudaStreamSynchronize(c->stream));
    return T3D_OK;
}

int t3d_ctx_set_rotation_mode(t3d_ctx* c, int mode)
{
    if (!c || (mode != T3D_ROT_FROZEN && mode != T3D_ROT_PER_PARTICLE)) return fail(T3D_ERR_INVALID, "bad rotation mode %d", mode);
    T3D_TRY(flush(c));
    c->rot_mode = mode;
    return T3D_OK;
}

int t3d_ctx_set_frozen_rotation(t3d_ctx* c, const float u[3], float angle)
{
    if (!c || !u) return fail(T3D_ERR_INVALID, "null argument");
    T3D_TRY(flush(c));
    T3D_TRY(use_device(c));
    ProcessStatics& st = c->statics();
    host_create_rotation(u, angle, st.frozen_m);
    st.frozen_latched = true;
    st.frozen_version++;
    return upload_frozen(c);
}

int t3d_ctx_get_frozen_rotation(t3d_ctx* c, float m[16], int* latched)
{
    if (!c) return fail(T3D_ERR_INVALID, "null context");
    T3D_TRY(flush(c));
    if (m) memcpy(m, c->statics().frozen_m, sizeof(c->statics().frozen_m));
    if (latched) *latched = c->statics().frozen_latched ? 1 : 0;
    return T3D_OK;
}

int t3d_ctx_reset_statics(t3d_ctx* c)
{
    if (!c) return fail(T3D_ERR_INVALID, "null context");
    T3D_TRY(flush(c));
    T3D_TRY(use_device(c));
    // (work other contexts of the process have queued belongs to the "old process": it goes out first)
    if (!c->private_statics)
        for (t3d_ctx* o : g_contexts)
            if (o != c && !o->private_statics) T3D_TRY(flush(o));
    T3D_TRY(use_device(c));
    ProcessStatics& st = c->statics();
    st.running_time = 0;
    st.frozen_latched = false;
    memset(st.frozen_m, 0, sizeof(st.frozen_m));
    st.frozen_version++;
    return upload_frozen(c);
}

int t3d_ctx_set_private_statics(t3d_ctx* c, int on)
{
    if (!c) return fail(T3D_ERR_INVALID, "null context");
    T3D_TRY(flush(c));
    if ((on != 0) == c->private_statics) return T3D_OK;
    if (on) c->own_statics = ProcessStatics(); // a fresh copy: "a process of its own"
    c->private_statics = on != 0;
    c->frozen_dev_version = 0;
    return T3D_OK;
}

int t3d_ctx_last_kernel_ms(t3d_ctx* c, int which, float* ms)
{
    if (!c || which < 0 || which >= KERNEL_KINDS || !ms) return fail(T3D_ERR_INVALID, "bad argument");
    T3D_TRY(flush(c));
    T3D_TRY(resolve_timing(c, which, true));
    if (!c->kernel_launches[which]) return fail(T3D_ERR_INVALID, "no launch of kind %d has been recorded", which);
    *ms = c->kernel_ms_last[which];
    return T3D_OK;
}

int t3d_ctx_kernel_time(t3d_ctx* c, int which, double* total_ms, uint64_t* n_launches)
{
    if (!c || which < 0 || which >= KERNEL_KINDS) return fail(T3D_ERR_INVALID, "bad argument");
    T3D_TRY(flush(c));
    T3D_TRY(resolve_timing(c, which, true));
    if (total_ms) *total_ms = c->kernel_ms_total[which];
    if (n_launches) *n_launches = c->kernel_launches[which];
    return T3D_OK;
}

int t3d_ctx_launch_count(t3d_ctx* c, uint64_t* n)
{
    if (!c || !n) return fail(T3D_ERR_INVALID, "null argument");
    *n = c->launches;
    return T3D_OK;
}

int t3d_ctx_host_profile(t3d_ctx* c, double seconds[T3D_HOST_PROFILE_SLOTS])
{
    if (!c || !seconds) return fail(T3D_ERR_INVALID, "null argument");
    for (int k = 0; k < T3D_HOST_PROFILE_SLOTS; ++k) seconds[k] = c->host_s[k];
    return T3D_OK;
}

int t3d_ctx_transfer_bytes(t3d_ctx* c, uint64_t* h2d, uint64_t* d2h)
{
    if (!c) return fail(T3D_ERR_INVALID, "null context");
    if (h2d) *h2d = c->h2d_bytes;
    if (d2h) *d2h = c->d2h_bytes;
    return T3D_OK;
}

int t3d_ctx_set_vertex_mode(t3d_ctx* c, int mode)
{
    if (!c || (mode != T3D_VERTEX_WORLD && mode != T3D_VERTEX_CLIP)) return fail(T3D_ERR_INVALID, "bad vertex mode %d", mode);
    T3D_TRY(flush(c));
    c->vertex_mode = mode;
    return T3D_OK;
}

int t3d_ctx_copy_to_host_async(t3d_ctx* c, void* host_dst, const void* device_src, size_t bytes, uint64_t* ticket)
{
    if (!c || !host_dst || !device_src || !ticket) return fail(T3D_ERR_INVALID, "copy_to_host_async: null argument");
    T3D_TRY(flush(c));
    T3D_TRY(use_device(c));
    if (!c->download_stream) T3D_CUDA(cudaStreamCreateWithFlags(&c->download_stream, cudaStreamNonBlocking));
    if (!c->download_fence) T3D_CUDA(cudaEventCreateWithFlags(&c->download_fence, cudaEventDisableTiming));
    // after everything the context has enqueued so far, beside whatever it enqueues next
    T3D_CUDA(cudaEventRecord(c->download_fence, c->stream));
    T3D_CUDA(cudaStreamWaitEvent(c->download_stream, c->download_fence, 0));
    T3D_CUDA(cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, c->download_stream));
    cudaEvent_t done;
    T3D_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    T3D_CUDA(cudaEventRecord(done, c->download_stream));
    // The launching stream must not overwrite the source before the copy has read it: a later launch that reuses the
    // buffer is the caller's to order (alternate two buffers and wait for ticket k before redrawing into buffer k).
    c->download_done.push_back(done);
    c->d2h_bytes += bytes;
    *ticket = c->download_done.size();
    return T3D_OK;
}

int t3d_ctx_copy_wait(t3d_ctx* c, uint64_t ticket)
{
    if (!c || ticket == 0 || ticket > c->download_done.size()) return fail(T3D_ERR_INVALID, "copy_wait: unknown ticket");
    cudaEvent_t& ev = c->download_done[ticket - 1];
    if (ev)
    {
        T3D_CUDA(cudaEventSynchronize(ev));
        cudaEventDestroy(ev);
        ev = nullptr;
    }
    return T3D_OK;
}

// ---- lifecycle ----
static void apply_params(t3d_emitter* e, const t3d_emitter_params* p)
{
    // PE.h:237 only rewrites the field; the storage keeps the size it got at create, so the field may go down (the spawn
    // clamp of PE.cpp:293-296 then uses the lower value) and back up to that size, never beyond (t3d_emitter_set_params)
    const uint32_t cap = std::min(p->particle_count_max, e->alloc_capacity);
    const float tw = e->p.texture_width, th = e->p.texture_height;
    e->p = *p;
    e->p.particle_count_max = cap;
    e->p.texture_width = tw;
    e->p.texture_height = th;
    e->time_per_emission = 1000.0f / (float)p->emission_rate;                 // PE.cpp:262-267
    e->frame_duration_secs = (float)p->sprite_frame_duration / 1000.0f;       // PE.cpp:491-495
    e->const_dirty = true;
}

int t3d_emitter_create(t3d_ctx* c, const t3d_emitter_params* p, t3d_emitter** out)
{
    if (!c || !p || !out) return fail(T3D_ERR_INVALID, "t3d_emitter_create: null argument");
    *out = nullptr;
    if (p->particle_count_max == 0) return fail(T3D_ERR_INVALID, "particle_count_max must be > 0");
    if (p->emission_rate == 0) return fail(T3D_ERR_INVALID, "emission_rate must be > 0 (PE.cpp:264)");
    if (!(p->texture_width >= 1.0f) || !(p->texture_height >= 1.0f))
        return fail(T3D_ERR_INVALID, "texture_width/height must be >= 1 (PE.cpp:54-55)");
    T3D_TRY(use_device(c));
    t3d_emitter* e = new t3d_emitter();
    e->ctx = c;
    e->p = *p;
    e->alloc_capacity = p->particle_count_max;
    int rc = alloc_segment(c, p->particle_count_max, &e->slab, &e->seg_offset, &e->seg_capacity);
    if (rc == T3D_OK) rc = alloc_dev_record(c, &e->dev);
    if (rc != T3D_OK)
    {
        delete e;
        return rc;
    }
    // setTexture, PE.cpp:244-251
    e->tex_w_ratio = 1.0f / (float)(unsigned int)p->texture_width;
    e->tex_h_ratio = 1.0f / (float)(unsigned int)p->texture_height;
    const float full[4] = { 0.0f, 0.0f, p->texture_width, p->texture_height };
    e->tex_coords.resize(4);
    t3d_host_sprite_rect_coords(1, full, e->tex_w_ratio, e->tex_h_ratio, e->tex_coords.data());
    e->frame_count = 1;
    e->percent_per_frame = 1.0f;
    apply_params(e, p);
    memset(e->node_world, 0, sizeof(e->node_world));
    memset(e->camera_world, 0, sizeof(e->camera_world));
    memset(e->view_projection, 0, sizeof(e->view_projection));
    e->view_projection[0] = e->view_projection[5] = e->view_projection[10] = e->view_projection[15] = 1.0f;
    c->emitters.push_back(e);
    *out = e;
    return T3D_OK;
}

int t3d_emitter_create_from_file(t3d_ctx* c, const char* url, t3d_emitter** out)
{
    if (!c || !url || !out) return fail(T3D_ERR_INVALID, "t3d_emitter_create_from_file: null argument");
    t3d_emitter_params p;
    uint32_t frame_count = 0;
    int sw = 0, sh = 0;
    char tex[1024];
    int rc = t3d_particle_file_parse(url, &p, &frame_count, &sw, &sh, tex, sizeof(tex));
    if (rc != T3D_OK) return fail(rc, "%s", t3d_host_last_error());
    T3D_TRY(t3d_emitter_create(c, &p, out));
    // PE.cpp:207
    return t3d_emitter_set_sprite_frame_coords(*out, frame_count, sw, sh);
}

int t3d_emitter_create_from_properties(t3d_ctx* c, const t3d_properties* particle, t3d_emitter** out)
{
    if (!c || !particle || !out) return fail(T3D_ERR_INVALID, "t3d_emitter_create_from_properties: null argument");
    t3d_emitter_params p;
    uint32_t frame_count = 0;
    int sw = 0, sh = 0;
    char tex[1024];
    int rc = t3d_particle_properties_parse(particle, &p, &frame_count, &sw, &sh, tex, sizeof(tex));
    if (rc != T3D_OK) return fail(rc, "%s", t3d_host_last_error());
    T3D_TRY(t3d_emitter_create(c, &p, out));
    return t3d_emitter_set_sprite_frame_coords(*out, frame_count, sw, sh); // PE.cpp:207
}

int t3d_emitter_clone(t3d_emitter* e, t3d_emitter** out)
{
    if (!e || !out) return fail(T3D_ERR_INVALID, "null argument");
    T3D_TRY(t3d_emitter_create(e->ctx, &e->p, out)); // PE.cpp:894-921
    t3d_emitter* n = *out;
    T3D_TRY(t3d_emitter_set_sprite_tex_coords(n, e->frame_count, e->tex_coords.data())); // PE.cpp:922
    n->rng_mode = e->rng_mode;
    n->rng_seed = e->rng_seed;
    return T3D_OK;
}

int t3d_emitter_destroy(t3d_emitter* e)
{
    if (!e) return T3D_OK;
    t3d_ctx* c = e->ctx;
    if (e->queued || e->held) flush(c);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (t3d_ctx::Feedback& fb : c->feedback)
        for (t3d_emitter*& o : fb.es)
            if (o == e) o = nullptr;
    if (e->uv_dev) cudaFree(e->uv_dev);
    if (e->dev) c->dev_free.push_back(e->dev);
    free_segment(c, e->slab, e->seg_offset, e->seg_capacity);
    c->emitters.erase(std::remove(c->emitters.begin(), c->emitters.end(), e), c->emitters.end());
    delete e;
    return T3D_OK;
}

// ---- configuration ----
int t3d_emitter_set_params(t3d_emitter* e, const t3d_emitter_params* p)
{
    if (!e || !p) return fail(T3D_ERR_INVALID, "null argument");
    if (p->emission_rate == 0) return fail(T3D_ERR_INVALID, "emission_rate must be > 0 (PE.cpp:264)");
    if (p->particle_count_max > e->alloc_capacity)
        return fail(T3D_ERR_CAPACITY, "particle_count_max cannot grow past the %u the storage was sized for at create (asked %u)",
                    e->alloc_capacity, p->particle_count_max);
    if (p->particle_count_max == 0) return fail(T3D_ERR_INVALID, "particle_count_max must be > 0");
    if (e->queued || e->held) T3D_TRY(flush(e->ctx));
    apply_params(e, p);
    return T3D_OK;
}

int t3d_emitter_get_params(t3d_emitter* e, t3d_emitter_params* p)
{
    if (!e || !p) return fail(T3D_ERR_INVALID, "null argument");
    *p = e->p;
    return T3D_OK;
}

int t3d_emitter_set_emission_rate(t3d_emitter* e, uint32_t rate)
{
    if (!e || rate == 0) return fail(T3D_ERR_INVALID, "emission rate must be > 0 (PE.cpp:264)");
    e->p.emission_rate = rate;
    e->time_per_emission = 1000.0f / (float)rate;
    return T3D_OK;
}

int t3d_emitter_set_texture_size(t3d_emitter* e, uint32_t width, uint32_t height, int blend_mode)
{
    if (!e || !width || !height) return fail(T3D_ERR_INVALID, "setTexture: the texture must have a size (PE.cpp:244-247)");
    if (blend_mode < T3D_BLEND_NONE || blend_mode > T3D_BLEND_MULTIPLIED) return fail(T3D_ERR_INVALID, "Unsupported blend mode (%d).", blend_mode);
    if (e->queued || e->held) T3D_TRY(flush(e->ctx));
    e->p.blend_mode = blend_mode;           // PE.cpp:243
    e->p.texture_width = (float)width;      // PE.cpp:244-247
    e->p.texture_height = (float)height;
    e->tex_w_ratio = 1.0f / (float)width;
    e->tex_h_ratio = 1.0f / (float)height;
    const float full[4] = { 0.0f, 0.0f, (float)width, (float)height }; // PE.cpp:249-251: one frame, the whole texture
    return t3d_emitter_set_sprite_frame_rects(e, 1, full);
}

static int install_tex_coords(t3d_emitter* e, uint32_t frame_count)
{
    e->frame_count = frame_count;
    e->percent_per_frame = 1.0f / (float)frame_count; // PE.cpp:518, 532
    e->uv_dirty = true;
    return T3D_OK;
}

int t3d_emitter_set_sprite_tex_coords(t3d_emitter* e, uint32_t frame_count, const float* tex_coords)
{
    if (!e || !frame_count || !tex_coords) return fail(T3D_ERR_INVALID, "setSpriteTexCoords: bad argument (PE.cpp:514-515)");
    if (e->queued || e->held) T3D_TRY(flush(e->ctx));
    e->tex_coords.assign(tex_coords, tex_coords + (size_t)frame_count * 4);
    return install_tex_coords(e, frame_count);
}

int t3d_emitter_set_sprite_frame_rects(t3d_emitter* e, uint32_t frame_count, const float* rects)
{
    if (!e || !frame_count || !rects) return fail(T3D_ERR_INVALID, "setSpriteFrameCoords: bad argument (PE.cpp:528-529)");
    if (e->queued || e->held) T3D_TRY(flush(e->ctx));
    e->tex_coords.resize((size_t)frame_count * 4);
    t3d_host_sprite_rect_coords(frame_count, rects, e->tex_w_ratio, e->tex_h_ratio, e->tex_coords.data());
    return install_tex_coords(e, frame_count);
}

int t3d_emitter_set_sprite_frame_coords(t3d_emitter* e, uint32_t frame_count, int width, int height)
{
    if (!e || !frame_count || !width || !height) return fail(T3D_ERR_INVALID, "setSpriteFrameCoords: bad argument (PE.cpp:552-553)");
    if (e->queued || e->held) T3D_TRY(flush(e->ctx));
    e->tex_coords.resize((size_t)frame_count * 4);
    t3d_host_sprite_frame_coords(frame_count, width, height, e->p.texture_width, e->p.texture_height, e->tex_coords.data());
    return install_tex_coords(e, frame_count);
}

int t3d_emitter_get_sprite_tex_coords(t3d_emitter* e, uint32_t* frame_count, float* tex_coords, size_t max_frames)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    if (frame_count) *frame_count = e->frame_count;
    if (tex_coords) memcpy(tex_coords, e->tex_coords.data(), sizeof(float) * 4 * std::min<size_t>(max_frames, e->frame_count));
    return T3D_OK;
}

int t3d_emitter_set_rng(t3d_emitter* e, int mode, uint64_t seed)
{
    if (!e || mode < T3D_RNG_LIBC || mode > T3D_RNG_DEVICE) return fail(T3D_ERR_INVALID, "bad rng mode %d", mode);
    if (e->queued || e->held) T3D_TRY(flush(e->ctx));
    e->rng_mode = mode;
    e->rng_seed = seed;
    e->const_dirty = true;
    return T3D_OK;
}

int t3d_emitter_feed_draws(t3d_emitter* e, const int32_t* draws, size_t n)
{
    if (!e || (!draws && n)) return fail(T3D_ERR_INVALID, "null argument");
    e->fed.insert(e->fed.end(), draws, draws + n);
    return T3D_OK;
}

int t3d_emitter_pending_draws(t3d_emitter* e, size_t* n)
{
    if (!e || !n) return fail(T3D_ERR_INVALID, "null argument");
    *n = e->fed.size() - e->fed_head;
    return T3D_OK;
}

int t3d_emitter_set_node_world(t3d_emitter* e, const float world[16])
{
    if (!e || !world) return fail(T3D_ERR_INVALID, "null argument");
    if (e->node_in_use && memcmp(e->node_world, world, sizeof(e->node_world)) != 0) T3D_TRY(flush(e->ctx));
    memcpy(e->node_world, world, sizeof(e->node_world));
    e->has_node = true;
    return T3D_OK;
}

int t3d_emitter_set_camera_world(t3d_emitter* e, const float m[16])
{
    if (!e || !m) return fail(T3D_ERR_INVALID, "null argument");
    if (e->camera_in_use && memcmp(e->camera_world, m, sizeof(e->camera_world)) != 0) T3D_TRY(flush(e->ctx));
    memcpy(e->camera_world, m, sizeof(e->camera_world));
    e->has_camera = true;
    return T3D_OK;
}

int t3d_emitter_set_view_projection(t3d_emitter* e, const float m[16])
{
    if (!e || !m) return fail(T3D_ERR_INVALID, "null argument");
    if (e->camera_in_use && memcmp(e->view_projection, m, sizeof(e->view_projection)) != 0) T3D_TRY(flush(e->ctx));
    memcpy(e->view_projection, m, sizeof(e->view_projection));
    e->has_vp = true;
    return T3D_OK;
}

int t3d_emitter_enable_bounds(t3d_emitter* e, int on)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    if (e->queued || e->held) T3D_TRY(flush(e->ctx));
    if (e->want_bounds != (on != 0))
    {
        e->want_bounds = on != 0;
        e->const_dirty = true;
        if (!on)
        {
            e->bounds_epoch = 0;
            e->box_known = false;
        }
    }
    return T3D_OK;
}

static int read_bounds(t3d_emitter* e, float lo[3], float hi[3], int* valid)
{
    t3d_ctx* c = e->ctx;
    *valid = 0;
    if (!e->want_bounds) return T3D_OK;
    if (e->queued || e->held) T3D_TRY(flush(c));
    if (!e->bounds_epoch) return T3D_OK;
    T3D_TRY(use_device(c));
    EmitterBounds b;
    T3D_CUDA(cudaMemcpyAsync(&b, &e->dev->bounds, sizeof(b), cudaMemcpyDeviceToHost, c->stream));
    T3D_CUDA(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += sizeof(b);
    unsigned long long w[6] = { b.lo[0], b.lo[1], b.lo[2], b.hi[0], b.hi[1], b.hi[2] };
    e->box_empty = !decode_box(w, e->bounds_epoch, e->box_lo, e->box_hi);
    e->box_known = true;
    if (e->box_empty) return T3D_OK;
    memcpy(lo, e->box_lo, sizeof(e->box_lo));
    memcpy(hi, e->box_hi, sizeof(e->box_hi));
    *valid = 1;
    return T3D_OK;
}

int t3d_emitter_bounds(t3d_emitter* e, float lo[3], float hi[3], int* valid)
{
    if (!e || !lo || !hi || !valid) return fail(T3D_ERR_INVALID, "null argument");
    return read_bounds(e, lo, hi, valid);
}

int t3d_emitter_set_culling(t3d_emitter* e, int on)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    if (on) T3D_TRY(t3d_emitter_enable_bounds(e, 1));
    e->culling = on != 0;
    return T3D_OK;
}

// ---- the path ----
int t3d_emitter_start(t3d_emitter* e)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    e->started = true; // PE.cpp:270-274 (does not reset _emitTime)
    return T3D_OK;
}

int t3d_emitter_stop(t3d_emitter* e)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    e->started = false;
    return T3D_OK;
}

int t3d_emitter_is_started(t3d_emitter* e, int* out)
{
    if (!e || !out) return fail(T3D_ERR_INVALID, "null argument");
    *out = e->started ? 1 : 0;
    return T3D_OK;
}

// need_exact = false: the caller can live with "maybe" (answered as active) when only a read-back could tell.
static int is_active(t3d_emitter* e, bool* out, bool need_exact = true)
{
    // PE.cpp:277-284
    if (e->started)
    {
        *out = true;
        return T3D_OK;
    }
    if (!e->has_node || e->count_ub == 0)
    {
        *out = false;
        return T3D_OK;
    }
    if (e->alive_budget_ms > 0) // some particle cannot have died yet: `_particleCount > 0` without asking the device
    {
        *out = true;
        return T3D_OK;
    }
    if (!need_exact && !e->count_exact)
    {
        *out = true;
        return T3D_OK;
    }
    uint32_t cnt;
    T3D_TRY(exact_count(e, &cnt));
    *out = cnt > 0;
    return T3D_OK;
}

int t3d_emitter_is_active(t3d_emitter* e, int* out)
{
    if (!e || !out) return fail(T3D_ERR_INVALID, "null argument");
    bool a;
    T3D_TRY(is_active(e, &a));
    *out = a ? 1 : 0;
    return T3D_OK;
}

int t3d_emitter_emit_once(t3d_emitter* e, uint32_t particle_count)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    t3d_ctx* c = e->ctx;
    if (c->phase != PHASE_EMIT || e->queued) T3D_TRY(flush(c));
    PendingOp op{};
    op.e = e;
    T3D_TRY(prepare_emit(e, particle_count, &op));
    if (op.emit_n == 0) return T3D_OK;
    return enqueue(c, PHASE_EMIT, op);
}

int t3d_emitter_update(t3d_emitter* e, float elapsed_time)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    t3d_ctx* c = e->ctx;
    bool active;
    T3D_TRY(is_active(e, &active));
    if (!active) return T3D_OK; // PE.cpp:715
    float elapsed_ms;
    if (!t3d_host_rate_cap(&c->statics().running_time, elapsed_time, &elapsed_ms)) return T3D_OK; // PE.cpp:720-725 (process-wide)
    if (c->phase != PHASE_UPDATE || e->queued) T3D_TRY(flush(c));
    PendingOp op{};
    op.e = e;
    op.elapsed_ms = elapsed_ms;
    op.offsets_begin = SIZE_MAX;
    if (e->started && e->p.emission_rate)
    {
        const uint32_t emit_count = t3d_host_emission_count(&e->emit_time, e->time_per_emission, elapsed_ms); // PE.cpp:732-743
        if (emit_count) T3D_TRY(prepare_emit(e, emit_count, &op));                                            // PE.cpp:744
    }
    if (e->count_ub > 0) e->count_exact = false; // particles may die in the kernel; the bound stays valid
    e->alive_budget_ms -= (double)elapsed_ms + 1.0;
    op.spawned = (uint32_t)std::min<uint64_t>(e->emitted_total - e->emitted_at_update, UINT32_MAX);
    e->emitted_at_update = e->emitted_total;
    return enqueue(c, PHASE_UPDATE, op);
}

// Is the emitter's last bounding box entirely outside the clip volume of its view-projection matrix?  Conservative:
// only "all eight corners beyond the same plane" culls (clip-space test -w <= x, y, z <= w).
static bool outside_frustum(const float* vp, const float lo[3], const float hi[3])
{
    float cx[8][4];
    for (int k = 0; k < 8; ++k)
    {
        const float x = (k & 1) ? hi[0] : lo[0], y = (k & 2) ? hi[1] : lo[1], z = (k & 4) ? hi[2] : lo[2];
        for (int r = 0; r < 4; ++r) cx[k][r] = x * vp[r] + y * vp[4 + r] + z * vp[8 + r] + vp[12 + r];
    }
    for (int axis = 0; axis < 3; ++axis)
        for (int side = 0; side < 2; ++side)
        {
            bool all_out = true;
            for (int k = 0; k < 8 && all_out; ++k) all_out = side ? (cx[k][axis] > cx[k][3]) : (cx[k][axis] < -cx[k][3]);
            if (all_out) return true;
        }
    return false;
}

static int draw_common(t3d_emitter* e, float* dst, size_t dst_quads, uint32_t* draw_calls)
{
    t3d_ctx* c = e->ctx;
    if (draw_calls) *draw_calls = 0;
    // The return value (PE.cpp:837: 0 when inactive) is the only thing that needs the exact live count of a stopped
    // emitter; a caller that does not ask for it (the batch call) is not made to wait for the device -- a draw of an
    // emitter whose particles have all died writes no quads either way.
    bool active;
    T3D_TRY(is_active(e, &active, draw_calls != nullptr));
    e->last_draw_empty = true;
    e->last_draw_dst = nullptr;
    if (!active) return T3D_OK; // PE.cpp:837
    if (draw_calls) *draw_calls = 1;
    if (e->count_ub == 0) return T3D_OK; // PE.cpp:839
    if (!e->has_node || !e->has_camera)
        return fail(T3D_ERR_NO_NODE, "draw: the emitter needs a node and an active camera node (PE.cpp:858-859)");
    if (dst)
    {
        if (reinterpret_cast<uintptr_t>(dst) & 15u) return fail(T3D_ERR_INVALID, "draw_to: the destination must be 16-byte aligned");
        if ((size_t)e->count_ub > dst_quads)
            return fail(T3D_ERR_CAPACITY, "draw_to: the destination holds %zu quads but the emitter may have %u live particles", dst_quads,
                        e->count_ub);
    }
    // Culling tests the last box that has reached the host -- it rides back with the asynchronous count feedback, so it
    // is a frame or two old unless t3d_emitter_bounds was just called -- and never waits for the device.
    if (e->culling && e->box_known && !e->box_empty && outside_frustum(e->view_projection, e->box_lo, e->box_hi))
        return T3D_OK; // nothing of it can reach the screen
    // SB.cpp:339 latches the rotation of the first rotated quad the PROCESS ever draws.  Until that has happened, draws
    // reach the device in call order across the contexts that share the statics: whatever the others still hold goes first.
    if (!c->private_statics && c->rot_mode == T3D_ROT_FROZEN && !g_statics.frozen_latched && g_contexts.size() > 1)
        for (t3d_ctx* o : g_contexts)
            if (o != c && !o->private_statics && (!o->pending.empty() || !o->held.empty())) T3D_TRY(flush(o));
    PendingOp op{};
    op.e = e;
    op.offsets_begin = SIZE_MAX;
    op.draw_dst = dst;
    e->last_draw_empty = false;
    e->last_draw_dst = dst;
    return enqueue(c, PHASE_DRAW, op);
}

int t3d_emitter_draw(t3d_emitter* e, uint32_t* draw_calls)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    return draw_common(e, nullptr, 0, draw_calls);
}

int t3d_emitter_draw_to(t3d_emitter* e, void* device_dst, size_t dst_quads, uint32_t* draw_calls)
{
    if (!e || !device_dst) return fail(T3D_ERR_INVALID, "draw_to: null argument");
    return draw_common(e, static_cast<float*>(device_dst), dst_quads, draw_calls);
}

int t3d_emitter_get_count(t3d_emitter* e, uint32_t* count)
{
    if (!e || !count) return fail(T3D_ERR_INVALID, "null argument");
    return refresh_counts(e->ctx, &e, 1, count, nullptr);
}

// ---- batches ----
int t3d_batch_update(t3d_emitter* const* es, size_t n, float elapsed_ms)
{
    if (!es && n) return fail(T3D_ERR_INVALID, "null argument");
    HostTimer ht(n ? es[0]->ctx : nullptr, T3D_HOST_ENQUEUE);
    for (size_t i = 0; i < n; ++i) T3D_TRY(t3d_emitter_update(es[i], elapsed_ms));
    return T3D_OK;
}

int t3d_batch_draw(t3d_emitter* const* es, size_t n)
{
    if (!es && n) return fail(T3D_ERR_INVALID, "null argument");
    HostTimer ht(n ? es[0]->ctx : nullptr, T3D_HOST_ENQUEUE);
    for (size_t i = 0; i < n; ++i) T3D_TRY(t3d_emitter_draw(es[i], nullptr));
    return T3D_OK;
}

int t3d_batch_emit_once(t3d_emitter* const* es, size_t n, const uint32_t* counts)
{
    if ((!es || !counts) && n) return fail(T3D_ERR_INVALID, "null argument");
    HostTimer ht(n ? es[0]->ctx : nullptr, T3D_HOST_ENQUEUE);
    for (size_t i = 0; i < n; ++i) T3D_TRY(t3d_emitter_emit_once(es[i], counts[i]));
    return T3D_OK;
}

int t3d_batch_get_counts(t3d_emitter* const* es, size_t n, uint32_t* counts)
{
    if ((!es || !counts) && n) return fail(T3D_ERR_INVALID, "null argument");
    if (n == 0) return T3D_OK;
    t3d_ctx* c = es[0]->ctx;
    for (size_t i = 0; i < n; ++i)
        if (es[i]->ctx != c) return fail(T3D_ERR_INVALID, "t3d_batch_get_counts: emitters belong to different contexts");
    return refresh_counts(c, es, n, counts, nullptr);
}

int t3d_batch_set_transforms(t3d_emitter* const* es, size_t n, const float* node_worlds, const float* camera_world)
{
    if (!es && n) return fail(T3D_ERR_INVALID, "null argument");
    HostTimer ht(n ? es[0]->ctx : nullptr, T3D_HOST_ENQUEUE);
    for (size_t i = 0; i < n; ++i)
    {
        if (node_worlds) T3D_TRY(t3d_emitter_set_node_world(es[i], node_worlds + 16 * i));
        if (camera_world) T3D_TRY(t3d_emitter_set_camera_world(es[i], camera_world));
    }
    return T3D_OK;
}

int t3d_batch_frame(t3d_emitter* const* es, size_t n, const float* node_worlds, const float* camera_world, float elapsed_ms, uint32_t* counts)
{
    if (!es && n) return fail(T3D_ERR_INVALID, "null argument");
    if (node_worlds || camera_world) T3D_TRY(t3d_batch_set_transforms(es, n, node_worlds, camera_world));
    T3D_TRY(t3d_batch_update(es, n, elapsed_ms));
    T3D_TRY(t3d_batch_draw(es, n));
    if (counts) return t3d_batch_get_counts(es, n, counts);
    return T3D_OK;
}

int t3d_batch_frame_pipelined(t3d_emitter* const* es, size_t n, const float* node_worlds, const float* camera_world, float elapsed_ms,
                              uint32_t* previous_counts, int* have_previous)
{
    if (!es && n) return fail(T3D_ERR_INVALID, "null argument");
    if (have_previous) *have_previous = 0;
    if (n == 0) return T3D_OK;
    t3d_ctx* c = es[0]->ctx;
    for (size_t i = 0; i < n; ++i)
        if (es[i]->ctx != c) return fail(T3D_ERR_INVALID, "t3d_batch_frame_pipelined: emitters belong to different contexts");
    const uint64_t serial_before = c->feedback_serial;
    if (node_worlds || camera_world) T3D_TRY(t3d_batch_set_transforms(es, n, node_worlds, camera_world));
    T3D_TRY(t3d_batch_update(es, n, elapsed_ms));
    T3D_TRY(t3d_batch_draw(es, n)); // (as two launches -- T3D_FUSED=0 -- the update goes out here, its read-back behind it)
    T3D_TRY(flush(c)); // the frame goes out now (its counts start travelling back behind it); nothing waits for it
    // The previous pipelined frame's counts: its kernel ran while this frame was being prepared, so the copy has normally landed.
    if (c->pipelined_serial)
    {
        for (t3d_ctx::Feedback& fb : c->feedback)
        {
            if (fb.serial != c->pipelined_serial || fb.es.size() != n) continue;
            bool same = true;
            for (size_t i = 0; same && i < n; ++i) same = fb.es[i] == es[i];
            if (!same) break;
            if (fb.in_flight)
            {
                {
                    HostTimer ht(c, T3D_HOST_WAIT);
                    T3D_CUDA(cudaEventSynchronize(fb.done));
                }
                absorb_feedback(c, fb);
            }
            T3D_TRY(check_fault(c));
            if (previous_counts)
                for (size_t i = 0; i < n; ++i) previous_counts[i] = fb.host[2 * i];
            if (have_previous) *have_previous = 1;
            break;
        }
    }
    // (a frame whose update was suppressed -- rate cap, inactive emitters -- or whose read-back found every slot busy leaves none)
    c->pipelined_serial = c->feedback_serial > serial_before ? c->feedback_serial : 0; // the newest read-back: this frame's update
    return T3D_OK;
}

// ---- results ----
int t3d_emitter_vertices(t3d_emitter* e, const t3d_sprite_vertex** device_ptr, uint32_t* n_quads)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    t3d_ctx* c = e->ctx;
    uint32_t cnt, drawn;
    T3D_TRY(refresh_counts(c, &e, 1, &cnt, &drawn));
    const Slab& s = c->slabs[e->slab];
    const float* own = s.vtx ? s.vtx + (size_t)e->seg_offset * s.vtx_floats : nullptr;
    const float* where = e->last_draw_dst ? e->last_draw_dst : own; // (t3d_clip_vertex in T3D_VERTEX_CLIP mode)
    if (device_ptr) *device_ptr = reinterpret_cast<const t3d_sprite_vertex*>(where);
    if (n_quads) *n_quads = (where && !e->last_draw_empty) ? drawn : 0;
    return T3D_OK;
}

int t3d_emitter_depth_sorted_indices(t3d_emitter* e, const float view_dir[3], const uint32_t** device_ptr, uint64_t* n_indices)
{
    if (!e || !view_dir) return fail(T3D_ERR_INVALID, "null argument");
    t3d_ctx* c = e->ctx;
    uint32_t n = 0;
    T3D_TRY(refresh_counts(c, &e, 1, &n, nullptr));
    if (device_ptr) *device_ptr = nullptr;
    if (n_indices) *n_indices = 0;
    if (n == 0) return T3D_OK;
    if (n > (1u << 30)) return fail(T3D_ERR_INVALID, "too many particles for 32-bit vertex indices");
    const size_t temp_bytes = depth_sort_temp_bytes(n);
    if (c->sort_cap < n || c->sort_temp_cap < temp_bytes)
    {
        T3D_CUDA(cudaStreamSynchronize(c->stream));
        if (c->sort_scratch) cudaFree(c->sort_scratch);
        if (c->sort_temp) cudaFree(c->sort_temp);
        if (c->sort_indices) cudaFree(c->sort_indices);
        c->sort_scratch = c->sort_indices = nullptr;
        c->sort_temp = nullptr;
        c->sort_cap = c->sort_temp_cap = 0;
        const size_t cap = std::max<size_t>((size_t)n + n / 4, 4096);
        const size_t tcap = std::max(depth_sort_temp_bytes((uint32_t)cap), temp_bytes);
        T3D_CUDA(cudaMalloc((void**)&c->sort_scratch, cap * 4 * sizeof(uint32_t)));
        T3D_CUDA(cudaMalloc((void**)&c->sort_indices, cap * 6 * sizeof(uint32_t)));
        T3D_CUDA(cudaMalloc(&c->sort_temp, std::max<size_t>(tcap, 16)));
        c->sort_cap = cap;
        c->sort_temp_cap = tcap;
    }
    const int err = launch_depth_sort(c->stream, e->hc.st.rw, e->hc.st.stride, n, view_dir, c->sort_scratch, c->sort_temp, c->sort_temp_cap,
                                      c->sort_indices);
    if (err) return fail(T3D_ERR_CUDA, "depth sort failed: %s", cudaGetErrorString((cudaError_t)err));
    T3D_CUDA(cudaGetLastError());
    c->launches += 3; // keys, indices, and the library's sort passes counted as one
    if (device_ptr) *device_ptr = c->sort_indices;
    if (n_indices) *n_indices = 6ull * n;
    return T3D_OK;
}

int t3d_emitter_download_depth_sorted_indices(t3d_emitter* e, const float view_dir[3], uint32_t* out, uint64_t max_indices, uint64_t* n_indices)
{
    const uint32_t* dev = nullptr;
    uint64_t n = 0;
    T3D_TRY(t3d_emitter_depth_sorted_indices(e, view_dir, &dev, &n));
    if (n_indices) *n_indices = n;
    if (!out || !n) return T3D_OK;
    t3d_ctx* c = e->ctx;
    const uint64_t m = std::min(n, max_indices);
    T3D_CUDA(cudaMemcpyAsync(out, dev, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    T3D_CUDA(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += m * sizeof(uint32_t);
    return T3D_OK;
}

int t3d_ctx_strip_indices(t3d_ctx* c, uint32_t n_quads, const uint32_t** device_ptr, uint64_t* n_indices)
{
    if (!c) return fail(T3D_ERR_INVALID, "null context");
    T3D_TRY(use_device(c));
    const uint64_t n = n_quads ? 6ull * n_quads - 2ull : 0ull;
    if (n_quads > c->strip_quads)
    {
        if (c->strip_indices)
        {
            T3D_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->strip_indices);
            c->strip_indices = nullptr;
            c->strip_quads = 0;
        }
        const uint32_t cap = std::max<uint32_t>(n_quads, 1024);
        const uint64_t cap_idx = 6ull * cap - 2ull;
        T3D_CUDA(cudaMalloc((void**)&c->strip_indices, cap_idx * sizeof(uint32_t)));
        launch_strip_indices(c->stream, c->strip_indices, cap_idx); // the pattern of n quads is a prefix of any longer one
        T3D_CUDA(cudaGetLastError());
        c->launches++;
        c->strip_quads = cap;
    }
    if (device_ptr) *device_ptr = c->strip_indices;
    if (n_indices) *n_indices = n;
    return T3D_OK;
}

int t3d_ctx_triangle_indices(t3d_ctx* c, uint32_t n_quads, const uint32_t** device_ptr, uint64_t* n_indices)
{
    if (!c) return fail(T3D_ERR_INVALID, "null context");
    T3D_TRY(use_device(c));
    if (n_quads > c->tri_quads)
    {
        if (c->tri_indices)
        {
            T3D_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->tri_indices);
            c->tri_indices = nullptr;
            c->tri_quads = 0;
        }
        const uint32_t cap = std::max<uint32_t>(n_quads, 1024);
        T3D_CUDA(cudaMalloc((void**)&c->tri_indices, 6ull * cap * sizeof(uint32_t)));
        launch_triangle_indices(c->stream, c->tri_indices, 6ull * cap);
        T3D_CUDA(cudaGetLastError());
        c->launches++;
        c->tri_quads = cap;
    }
    if (device_ptr) *device_ptr = c->tri_indices;
    if (n_indices) *n_indices = 6ull * n_quads;
    return T3D_OK;
}

int t3d_ctx_download_triangle_indices(t3d_ctx* c, uint32_t n_quads, uint32_t* out, uint64_t max_indices, uint64_t* n_indices)
{
    const uint32_t* dev = nullptr;
    uint64_t n = 0;
    T3D_TRY(t3d_ctx_triangle_indices(c, n_quads, &dev, &n));
    if (n_indices) *n_indices = n;
    const uint64_t m = std::min<uint64_t>(n, max_indices);
    if (out && m)
    {
        T3D_CUDA(cudaMemcpyAsync(out, dev, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        T3D_CUDA(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += m * sizeof(uint32_t);
    }
    return T3D_OK;
}

int t3d_ctx_download_strip_indices(t3d_ctx* c, uint32_t n_quads, uint32_t* out, uint64_t max_indices, uint64_t* n_indices)
{
    const uint32_t* dev = nullptr;
    uint64_t n = 0;
    T3D_TRY(t3d_ctx_strip_indices(c, n_quads, &dev, &n));
    if (n_indices) *n_indices = n;
    const uint64_t m = std::min<uint64_t>(n, max_indices);
    if (out && m)
    {
        T3D_CUDA(cudaMemcpyAsync(out, dev, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        T3D_CUDA(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += m * sizeof(uint32_t);
    }
    return T3D_OK;
}

int t3d_emitter_download_vertices(t3d_emitter* e, float* out, size_t max_quads, uint32_t* n_quads)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    const t3d_sprite_vertex* dev;
    uint32_t n;
    T3D_TRY(t3d_emitter_vertices(e, &dev, &n));
    if (n_quads) *n_quads = n;
    const size_t m = std::min<size_t>(n, max_quads);
    if (out && m)
    {
        t3d_ctx* c = e->ctx;
        const size_t quad_bytes = (c->vertex_mode == T3D_VERTEX_CLIP ? T3D_FLOATS_PER_CLIP_QUAD : T3D_FLOATS_PER_QUAD) * sizeof(float);
        T3D_CUDA(cudaMemcpyAsync(out, dev, m * quad_bytes, cudaMemcpyDeviceToHost, c->stream));
        T3D_CUDA(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += m * quad_bytes;
    }
    return T3D_OK;
}

int t3d_emitter_download_state(t3d_emitter* e, uint32_t* out, size_t stride, uint32_t* count)
{
    if (!e) return fail(T3D_ERR_INVALID, "null emitter");
    t3d_ctx* c = e->ctx;
    uint32_t n;
    T3D_TRY(refresh_counts(c, &e, 1, &n, nullptr));
    if (count) *count = n;
    if (!out || n == 0) return T3D_OK;
    if (stride < n) return fail(T3D_ERR_INVALID, "download_state: stride %zu < live count %u", stride, n);
    {
        // the device keeps rw columns + records: a kernel rearranges them into the 34 exchange columns
        EmitterConst k;
        fill_const(e, &k);
        uint32_t* tmp = nullptr;
        T3D_CUDA(cudaMalloc((void**)&tmp, (size_t)n * T3D_NUM_COLUMNS * sizeof(uint32_t)));
        launch_export_state(c->stream, k.st, n, tmp, n);
        c->launches++;
        cudaError_t err = cudaMemcpy2DAsync(out, stride * sizeof(uint32_t), tmp, (size_t)n * sizeof(uint32_t), (size_t)n * sizeof(uint32_t),
                                            T3D_NUM_COLUMNS, cudaMemcpyDeviceToHost, c->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        if (err != cudaSuccess) return fail(T3D_ERR_CUDA, "download_state: %s", cudaGetErrorString(err));
    }
    c->d2h_bytes += (size_t)n * T3D_NUM_COLUMNS * sizeof(uint32_t);
    return T3D_OK;
}

int t3d_emitter_upload_state(t3d_emitter* e, const uint32_t* in, size_t stride, uint32_t count)
{
    if (!e || (!in && count)) return fail(T3D_ERR_INVALID, "null argument");
    if (count > e->p.particle_count_max) return fail(T3D_ERR_CAPACITY, "upload_state: %u particles exceed capacity %u", count, e->p.particle_count_max);
    if (stride < count) return fail(T3D_ERR_INVALID, "upload_state: stride %zu < count %u", stride, count);
    t3d_ctx* c = e->ctx;
    T3D_TRY(flush(c));
    T3D_TRY(use_device(c));
    uint32_t* tmp = nullptr;
    if (count)
    {
        EmitterConst k;
        fill_const(e, &k);
        T3D_CUDA(cudaMalloc((void**)&tmp, (size_t)count * T3D_NUM_COLUMNS * sizeof(uint32_t)));
        cudaError_t err = cudaMemcpy2DAsync(tmp, (size_t)count * sizeof(uint32_t), in, stride * sizeof(uint32_t),
                                            (size_t)count * sizeof(uint32_t), T3D_NUM_COLUMNS, cudaMemcpyHostToDevice, c->stream);
        if (err != cudaSuccess)
        {
            cudaFree(tmp);
            return fail(T3D_ERR_CUDA, "upload_state: %s", cudaGetErrorString(err));
        }
        launch_import_state(c->stream, k.st, count, tmp, count);
        c->launches++;
    }
    T3D_CUDA(cudaMemcpyAsync(&e->dev->cnt[e->parity].count, &count, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    T3D_CUDA(cudaStreamSynchronize(c->stream));
    if (tmp) cudaFree(tmp);
    c->h2d_bytes += (size_t)count * T3D_NUM_COLUMNS * sizeof(uint32_t) + sizeof(uint32_t);
    e->launch_version++;
    e->count_ub = count;
    e->count_exact = true;
    e->emitted_total += count; // keeps count feedback still in flight a valid upper bound
    e->alive_budget_ms = 0;
    // Uploaded particles may carry any rotation state.
    e->may_rotate = e->may_spin = true;
    e->const_dirty = true;
    return T3D_OK;
}

// --------------------------------------------------------------------------------------------------
// 2-D sprites and tile sets
// --------------------------------------------------------------------------------------------------
} // extern "C"

struct t3d_tileset
{
    t3d_ctx* ctx{ nullptr };
    uint32_t rows{ 0 }, cols{ 0 };
    float tile_w{ 0 }, tile_h{ 0 }, tex_w_ratio{ 0 }, tex_h_ratio{ 0 };
    std::vector<float> tiles;      // rows * cols * {x, y}; negative = blank (TileSet.cpp:41-42 fills the table with -1)
    float* tiles_dev{ nullptr };
    uint32_t* cells_dev{ nullptr }; // indices of the non-blank cells, row-major
    uint32_t n_quads{ 0 };
    bool dirty{ true };
};

struct t3d_font
{
    t3d_ctx* ctx{ nullptr };
    uint32_t size{ 0 };     // Font::_size
    float spacing{ 0.0f };  // Font::_spacing (Font.h:426)
    std::vector<t3d_glyph> glyphs;
    t3d_glyph* glyphs_dev{ nullptr };
};

// Where a list's quads go: the caller's buffer, or the context's own (grown on demand).
static int sprite_destination(t3d_ctx* c, size_t quads, void* device_dst, size_t dst_quads, float** out)
{
    if (device_dst)
    {
        if (reinterpret_cast<uintptr_t>(device_dst) & 15u) return fail(T3D_ERR_INVALID, "the destination must be 16-byte aligned");
        if (quads > dst_quads) return fail(T3D_ERR_CAPACITY, "the destination holds %zu quads, the list may write %zu", dst_quads, quads);
        *out = static_cast<float*>(device_dst);
        return T3D_OK;
    }
    if (c->sprite_vtx_quads < quads)
    {
        if (c->sprite_vtx)
        {
            T3D_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->sprite_vtx);
            c->sprite_vtx = nullptr;
        }
        const size_t cap = std::max<size_t>(quads + quads / 2, 4096);
        T3D_CUDA(cudaMalloc((void**)&c->sprite_vtx, cap * T3D_FLOATS_PER_QUAD * sizeof(float)));
        c->sprite_vtx_quads = cap;
    }
    *out = c->sprite_vtx;
    return T3D_OK;
}

extern "C" {

// The pinned / device buffer pair the next host list goes through, grown to `bytes` and no longer in flight.
static int sprite_stage_acquire(t3d_ctx* c, size_t bytes, t3d_ctx::SpriteStage** out)
{
    t3d_ctx::SpriteStage* stage = &c->sprite_stage[c->sprite_stage_next];
    if (stage->in_flight)
    {
        HostTimer ht(c, T3D_HOST_WAIT);
        T3D_CUDA(cudaEventSynchronize(stage->done));
        stage->in_flight = false;
    }
    if (stage->cap < bytes)
    {
        if (stage->host) cudaFreeHost(stage->host);
        if (stage->dev) cudaFree(stage->dev);
        stage->host = stage->dev = nullptr;
        stage->cap = 0;
        const size_t cap = std::max<size_t>(bytes + bytes / 2, 1 << 16);
        T3D_CUDA(cudaMallocHost((void**)&stage->host, cap));
        T3D_CUDA(cudaMalloc((void**)&stage->dev, cap));
        stage->cap = cap;
    }
    if (!stage->done) T3D_CUDA(cudaEventCreateWithFlags(&stage->done, cudaEventDisableTiming));
    *out = stage;
    return T3D_OK;
}

int t3d_ctx_map_sprites(t3d_ctx* c, size_t n, t3d_sprite2d** host_list)
{
    if (!c || !host_list) return fail(T3D_ERR_INVALID, "t3d_ctx_map_sprites: null argument");
    T3D_TRY(use_device(c));
    t3d_ctx::SpriteStage* stage;
    T3D_TRY(sprite_stage_acquire(c, std::max<size_t>(n, 1) * sizeof(t3d_sprite2d), &stage));
    *host_list = reinterpret_cast<t3d_sprite2d*>(stage->host);
    return T3D_OK;
}

int t3d_ctx_draw_sprites(t3d_ctx* c, const t3d_sprite2d* sprites, size_t n, int list_flags, void* device_dst, size_t dst_quads,
                         const t3d_sprite_vertex** device_ptr, uint32_t* n_quads)
{
    if (!c || (!sprites && n)) return fail(T3D_ERR_INVALID, "t3d_ctx_draw_sprites: null argument");
    if (n > 0xfffffff0u) return fail(T3D_ERR_INVALID, "t3d_ctx_draw_sprites: too many sprites");
    T3D_TRY(use_device(c));
    float* dst = nullptr;
    T3D_TRY(sprite_destination(c, n, device_dst, dst_quads, &dst));
    if (device_ptr) *device_ptr = reinterpret_cast<const t3d_sprite_vertex*>(dst);
    if (n_quads) *n_quads = 0;
    if (n == 0) return T3D_OK;
    const t3d_sprite2d* dev_list = sprites;
    bool may_cull = !(list_flags & T3D_SPRITES_NO_CLIP); // (a device-resident list is not inspected here)
    t3d_ctx::SpriteStage* stage = nullptr;
    if (!(list_flags & T3D_SPRITES_ON_DEVICE))
    {
        const size_t bytes = n * sizeof(t3d_sprite2d);
        // the list may already lie in the buffer t3d_ctx_map_sprites handed out: then there is nothing to copy
        t3d_ctx::SpriteStage* mapped = &c->sprite_stage[c->sprite_stage_next];
        const bool in_place = mapped->host && reinterpret_cast<const char*>(sprites) == mapped->host && bytes <= mapped->cap && !mapped->in_flight;
        if (in_place)
            stage = mapped;
        else
            T3D_TRY(sprite_stage_acquire(c, bytes, &stage));
        c->sprite_stage_next ^= 1;
        {
            HostTimer ht(c, T3D_HOST_BUILD);
            if (!in_place) memcpy(stage->host, sprites, bytes);
            if (may_cull)
            {
                may_cull = false;
                for (size_t i = 0; i < n && !may_cull; ++i) may_cull = (sprites[i].flags & T3D_SPRITE_CLIPPED) != 0;
            }
        }
        T3D_CUDA(cudaMemcpyAsync(stage->dev, stage->host, bytes, cudaMemcpyHostToDevice, c->stream));
        c->h2d_bytes += bytes;
        dev_list = reinterpret_cast<const t3d_sprite2d*>(stage->dev);
    }
    else if (reinterpret_cast<uintptr_t>(sprites) & 15u)
        return fail(T3D_ERR_INVALID, "t3d_ctx_draw_sprites: a device-resident list must be 16-byte aligned");
    uint32_t* scratch = nullptr;
    const uint32_t tiles = sprite_tiles((uint32_t)n);
    if (may_cull)
    {
        if (c->sprite_scratch_words < (size_t)tiles + 1)
        {
            if (c->sprite_scratch)
            {
                T3D_CUDA(cudaStreamSynchronize(c->stream));
                cudaFree(c->sprite_scratch);
            }
            const size_t cap = std::max<size_t>((size_t)tiles * 2 + 1, 1024);
            T3D_CUDA(cudaMalloc((void**)&c->sprite_scratch, cap * sizeof(uint32_t)));
            c->sprite_scratch_words = cap;
        }
        scratch = c->sprite_scratch;
    }
    launch_sprites(c->stream, dev_list, (uint32_t)n, scratch, dst);
    T3D_CUDA(cudaGetLastError());
    c->launches += scratch ? 3 : 1;
    if (stage)
    {
        T3D_CUDA(cudaEventRecord(stage->done, c->stream));
        stage->in_flight = true;
    }
    if (n_quads)
    {
        if (!scratch)
            *n_quads = (uint32_t)n;
        else
        {
            // culling is decided on the device: the quad count has to come back
            uint32_t total = 0;
            T3D_CUDA(cudaMemcpyAsync(&total, scratch + tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            {
                HostTimer ht(c, T3D_HOST_WAIT);
                T3D_CUDA(cudaStreamSynchronize(c->stream));
            }
            c->d2h_bytes += sizeof(uint32_t);
            *n_quads = total;
        }
    }
    return T3D_OK;
}

// ---- Font::drawText ----
int t3d_font_create(t3d_ctx* c, uint32_t size, const t3d_glyph* glyphs, uint32_t glyph_count, t3d_font** out)
{
    if (!c || !out || !glyphs) return fail(T3D_ERR_INVALID, "t3d_font_create: null argument (Font.cpp:111-112)");
    *out = nullptr;
    if (!size) return fail(T3D_ERR_INVALID, "a font has a size (Font.cpp:249)");
    if (!glyph_count) return fail(T3D_ERR_INVALID, "a font has glyphs: ' ' and '\\t' advance by the first one's (Font.cpp:344)");
    T3D_TRY(use_device(c));
    t3d_font* f = new t3d_font();
    f->ctx = c;
    f->size = size;
    f->glyphs.assign(glyphs, glyphs + glyph_count);
    if (cudaMalloc((void**)&f->glyphs_dev, glyph_count * sizeof(t3d_glyph)) != cudaSuccess ||
        cudaMemcpyAsync(f->glyphs_dev, f->glyphs.data(), glyph_count * sizeof(t3d_glyph), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
    {
        cudaFree(f->glyphs_dev);
        delete f;
        return fail(T3D_ERR_CUDA, "t3d_font_create: %s", cudaGetErrorString(cudaGetLastError()));
    }
    // (the copy reads f->glyphs, which lives as long as the font; pageable memory: the call returns once it is staged)
    c->h2d_bytes += glyph_count * sizeof(t3d_glyph);
    c->fonts.push_back(f);
    *out = f;
    return T3D_OK;
}

int t3d_font_destroy(t3d_font* f)
{
    if (!f) return T3D_OK;
    t3d_ctx* c = f->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(f->glyphs_dev);
    c->fonts.erase(std::remove(c->fonts.begin(), c->fonts.end(), f), c->fonts.end());
    delete f;
    return T3D_OK;
}

int t3d_font_set_character_spacing(t3d_font* f, float spacing)
{
    if (!f) return fail(T3D_ERR_INVALID, "null font");
    f->spacing = spacing;
    return T3D_OK;
}

int t3d_font_draw_text(t3d_font* f, const char* text, size_t length, uint32_t text_flags, int x, int y, const float color[4], uint32_t size,
                       int right_to_left, void* device_dst, size_t dst_quads, const t3d_sprite_vertex** device_ptr, uint32_t* n_quads)
{
    if (!f || (!text && length) || !color) return fail(T3D_ERR_INVALID, "t3d_font_draw_text: null argument");
    if (length > 0x7fffffffu) return fail(T3D_ERR_INVALID, "t3d_font_draw_text: text too long (2 GiB per call)");
    t3d_ctx* c = f->ctx;
    T3D_TRY(use_device(c));
    if (size == 0) size = f->size; // Font.cpp:253-256
    const bool on_device = (text_flags & T3D_TEXT_ON_DEVICE) != 0;
    // right to left the reference walks text.c_str(): the text ends at its first NUL (Font.cpp:275, :301-303)
    if (right_to_left && !on_device) length = strnlen(text, length);
    if (on_device && (reinterpret_cast<uintptr_t>(text) & 15u)) return fail(T3D_ERR_INVALID, "t3d_font_draw_text: device-resident text must be 16-byte aligned");
    // how many characters draw: known here for host text (Font.cpp:354-357: a signed char with 0 <= c - 32 < glyphCount that
    // is not ' '); for device text every byte might
    size_t quads = length;
    if (!on_device)
    {
        HostTimer ht(c, T3D_HOST_BUILD);
        quads = 0;
        const uint32_t gc = (uint32_t)f->glyphs.size();
        for (size_t i = 0; i < length; ++i)
        {
            const unsigned char b = (unsigned char)text[i];
            quads += (b > 32u && b < 128u && b - 32u < gc) ? 1u : 0u;
        }
    }
    float* dst = nullptr;
    T3D_TRY(sprite_destination(c, quads, device_dst, dst_quads, &dst));
    if (device_ptr) *device_ptr = reinterpret_cast<const t3d_sprite_vertex*>(dst);
    if (n_quads) *n_quads = 0;
    if (length == 0) return T3D_OK;
    const char* dev_text = text;
    t3d_ctx::SpriteStage* stage = nullptr;
    if (!on_device)
    {
        T3D_TRY(sprite_stage_acquire(c, length, &stage));
        c->sprite_stage_next ^= 1;
        {
            HostTimer ht(c, T3D_HOST_BUILD);
            memcpy(stage->host, text, length);
        }
        T3D_CUDA(cudaMemcpyAsync(stage->dev, stage->host, length, cudaMemcpyHostToDevice, c->stream));
        c->h2d_bytes += length;
        dev_text = stage->dev;
    }
    const size_t scratch_words = (text_scratch_bytes((uint32_t)length) + 3) / 4;
    if (c->sprite_scratch_words < scratch_words)
    {
        if (c->sprite_scratch)
        {
            T3D_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->sprite_scratch);
            c->sprite_scratch = nullptr;
            c->sprite_scratch_words = 0;
        }
        const size_t cap = std::max<size_t>(scratch_words * 2, 1024);
        T3D_CUDA(cudaMalloc((void**)&c->sprite_scratch, cap * sizeof(uint32_t)));
        c->sprite_scratch_words = cap;
    }
    TextParams P{};
    P.x = x;
    P.y = y;
    P.size = size;
    P.scale = (float)size / (float)f->size;   // Font.cpp:268
    P.spacing = (int)((float)size * f->spacing); // :269
    P.glyph_count = (uint32_t)f->glyphs.size();
    P.space_advance = f->glyphs[0].advance;
    memcpy(P.color, color, sizeof(P.color));
    TextTables T{};
    for (uint32_t ch = 0; ch < (uint32_t)kTextClasses; ++ch)
    {
        if (ch == ' ') { T.adv[ch] = (int)P.space_advance; T.kind[ch] = 3; }                  // Font.cpp:343-345
        else if (ch == '\t') { T.adv[ch] = (int)(P.space_advance * 4u); T.kind[ch] = 3; }     // :350-352
        else if (ch == '\r' || ch == '\n') T.kind[ch] = 1;                                    // :346-349
        else if (ch >= 32u && ch - 32u < P.glyph_count)                                       // :354-355 (index = c - 32)
        {
            T.kind[ch] = 2;
            T.adv[ch] = (int)floorf((float)f->glyphs[ch - 32u].advance * P.scale + (float)P.spacing); // :378
        }
    }
    launch_text(c->stream, dev_text, (uint32_t)length, f->glyphs_dev, T, P, right_to_left != 0, c->sprite_scratch, dst);
    T3D_CUDA(cudaGetLastError());
    c->launches += 3;
    if (stage)
    {
        T3D_CUDA(cudaEventRecord(stage->done, c->stream));
        stage->in_flight = true;
    }
    if (n_quads)
    {
        if (!on_device)
            *n_quads = (uint32_t)quads;
        else
        {
            uint32_t total[4] = { 0, 0, 0, 0 }; // {advance, line breaks, glyphs, .}: the first words of the whole text's element
            T3D_CUDA(cudaMemcpyAsync(total, reinterpret_cast<const char*>(c->sprite_scratch) + text_scratch_bytes((uint32_t)length) - 48, 16,
                                     cudaMemcpyDeviceToHost, c->stream));
            {
                HostTimer ht(c, T3D_HOST_WAIT);
                T3D_CUDA(cudaStreamSynchronize(c->stream));
            }
            c->d2h_bytes += 16;
            *n_quads = total[2];
        }
    }
    return T3D_OK;
}

int t3d_tileset_create(t3d_ctx* c, uint32_t texture_width, uint32_t texture_height, float tile_width, float tile_height, uint32_t row_count,
                       uint32_t column_count, t3d_tileset** out)
{
    if (!c || !out) return fail(T3D_ERR_INVALID, "t3d_tileset_create: null argument");
    *out = nullptr;
    if (!texture_width || !texture_height) return fail(T3D_ERR_INVALID, "the texture must have a size");
    if (!(tile_width > 0) || !(tile_height > 0)) return fail(T3D_ERR_INVALID, "tile width and height must be positive (TileSet.cpp:36)");
    if (!row_count || !column_count) return fail(T3D_ERR_INVALID, "row and column counts must be positive (TileSet.cpp:37)");
    if ((uint64_t)row_count * column_count > 0x7fffffffull) return fail(T3D_ERR_INVALID, "too many cells");
    T3D_TRY(use_device(c));
    t3d_tileset* t = new t3d_tileset();
    t->ctx = c;
    t->rows = row_count;
    t->cols = column_count;
    t->tile_w = tile_width;
    t->tile_h = tile_height;
    t->tex_w_ratio = 1.0f / (float)texture_width;   // SB.cpp:155-156
    t->tex_h_ratio = 1.0f / (float)texture_height;
    t->tiles.assign((size_t)row_count * column_count * 2, -1.0f);
    const size_t cells = (size_t)row_count * column_count;
    if (cudaMalloc((void**)&t->tiles_dev, cells * 2 * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void**)&t->cells_dev, cells * sizeof(uint32_t)) != cudaSuccess)
    {
        if (t->tiles_dev) cudaFree(t->tiles_dev);
        delete t;
        return fail(T3D_ERR_CUDA, "t3d_tileset_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    c->tilesets.push_back(t);
    *out = t;
    return T3D_OK;
}

int t3d_tileset_destroy(t3d_tileset* t)
{
    if (!t) return T3D_OK;
    t3d_ctx* c = t->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(t->tiles_dev);
    cudaFree(t->cells_dev);
    c->tilesets.erase(std::remove(c->tilesets.begin(), c->tilesets.end(), t), c->tilesets.end());
    delete t;
    return T3D_OK;
}

int t3d_tileset_set_tile_source(t3d_tileset* t, uint32_t column, uint32_t row, float x, float y)
{
    if (!t || column >= t->cols || row >= t->rows) return fail(T3D_ERR_INVALID, "setTileSource: cell out of range (TileSet.cpp:157-158)");
    float* cell = &t->tiles[2 * ((size_t)row * t->cols + column)];
    cell[0] = x;
    cell[1] = y;
    t->dirty = true;
    return T3D_OK;
}

int t3d_tileset_set_tiles(t3d_tileset* t, const float* sources)
{
    if (!t || !sources) return fail(T3D_ERR_INVALID, "null argument");
    memcpy(t->tiles.data(), sources, t->tiles.size() * sizeof(float));
    t->dirty = true;
    return T3D_OK;
}

int t3d_tileset_draw(t3d_tileset* t, const float position[3], const float color[4], void* device_dst, size_t dst_quads,
                     const t3d_sprite_vertex** device_ptr, uint32_t* n_quads)
{
    if (!t || !position || !color) return fail(T3D_ERR_INVALID, "t3d_tileset_draw: null argument");
    t3d_ctx* c = t->ctx;
    T3D_TRY(use_device(c));
    const size_t cells = (size_t)t->rows * t->cols;
    if (t->dirty)
    {
        // the table changed: upload it and the list of non-blank cells (TileSet.cpp:217) in draw order
        std::vector<uint32_t> live;
        live.reserve(cells);
        for (size_t k = 0; k < cells; ++k)
            if (t->tiles[2 * k] >= 0 && t->tiles[2 * k + 1] >= 0) live.push_back((uint32_t)k);
        T3D_CUDA(cudaStreamSynchronize(c->stream)); // (a draw in flight may still read the old table)
        T3D_CUDA(cudaMemcpy(t->tiles_dev, t->tiles.data(), cells * 2 * sizeof(float), cudaMemcpyHostToDevice));
        if (!live.empty()) T3D_CUDA(cudaMemcpy(t->cells_dev, live.data(), live.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        c->h2d_bytes += cells * 2 * sizeof(float) + live.size() * sizeof(uint32_t);
        t->n_quads = (uint32_t)live.size();
        t->dirty = false;
    }
    float* dst = nullptr;
    T3D_TRY(sprite_destination(c, t->n_quads, device_dst, dst_quads, &dst));
    if (device_ptr) *device_ptr = reinterpret_cast<const t3d_sprite_vertex*>(dst);
    if (n_quads) *n_quads = t->n_quads;
    if (!t->n_quads) return T3D_OK;
    // TileSet.cpp:207-235: the loop's running position -- sequential float additions, one per column and per row -- is the
    // only per-frame input: cols + rows floats
    StagingSlot* slot;
    const size_t bytes = ((size_t)t->cols + t->rows) * sizeof(float);
    T3D_TRY(staging_acquire(c, bytes, &slot));
    {
        HostTimer ht(c, T3D_HOST_BUILD);
        float* xs = reinterpret_cast<float*>(slot->host);
        float* ys = xs + t->cols;
        float x = position[0];
        for (uint32_t col = 0; col < t->cols; ++col)
        {
            xs[col] = x;
            x += t->tile_w; // :231
        }
        float y = position[1];
        y += t->tile_h * (float)(t->rows - 1); // :208
        for (uint32_t row = 0; row < t->rows; ++row)
        {
            ys[row] = y;
            y -= t->tile_h; // :234
        }
    }
    T3D_TRY(staging_submit(c, slot, bytes));
    const float* xs_dev = reinterpret_cast<const float*>(slot->dev);
    launch_tileset(c->stream, t->cells_dev, t->n_quads, t->tiles_dev, t->cols, xs_dev, xs_dev + t->cols, position[2], t->tile_w, t->tile_h,
                   t->tex_w_ratio, t->tex_h_ratio, color, dst);
    T3D_CUDA(cudaGetLastError());
    c->launches++;
    T3D_TRY(staging_release(c, slot));
    return T3D_OK;
}

int t3d_ctx_download_quads(t3d_ctx* c, const t3d_sprite_vertex* device_ptr, uint32_t n_quads, float* out)
{
    if (!c || (!device_ptr && n_quads) || (!out && n_quads)) return fail(T3D_ERR_INVALID, "null argument");
    T3D_TRY(flush(c));
    T3D_TRY(use_device(c));
    if (!n_quads) return T3D_OK;
    const size_t bytes = (size_t)n_quads * T3D_FLOATS_PER_QUAD * sizeof(float);
    T3D_CUDA(cudaMemcpyAsync(out, device_ptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    T3D_CUDA(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += bytes;
    return T3D_OK;
}

} // extern "C"
