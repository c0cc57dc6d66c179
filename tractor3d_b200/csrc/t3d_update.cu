// t3d_update.cu -- the update + dead-particle removal kernel (replaces the per-particle loop of
// ParticleEmitter::update, PE.cpp:750-831 of the reference) and its FUSED form, which also expands the frame's
// billboard quads (ParticleEmitter::draw + SpriteBatch::draw, PE.cpp:866-882, SB.cpp:292-363) in the same pass.
//
// One thread owns V consecutive particles of one emitter and reads/writes every column with one V x 32-bit
// access, so a warp touches 32 x V x 4 contiguous bytes per column per instruction.  The kernel is a pure
// HBM stream: 144 B per particle (160 B with sprite animation, +28 B with axis rotation), no reuse; the fused
// form adds the 144 B of vertices and saves the draw kernel's 40-byte re-read.
//
// removal.  The reference walks i = 0.. and, when particle i is dead, copies the LAST live particle into
// slot i, shrinks the count and moves on WITHOUT examining the particle it just moved (PE.cpp:821-830 and
// the loop increment at :750).  With N = count on entry, dead(k) evaluated on the ORIGINAL particle k and
// D(k) = number of dead originals with index < k:
//     original k is examined        <=>  k + D(k) < N
//     examined and alive            ->   updated in place
//     examined and dead             ->   slot k receives the untouched original N-1-D(k) (nothing if that is k)
//     new count                     =    N - D(K), K = number of examined originals
// Sources (>= new count, never examined, never written) and destinations (< K) are disjoint, so the fill is
// race-free in place.  D(k) is one exclusive scan: within a tile by warp shuffles, across the tiles of an
// emitter by decoupled look-back over per-tile status words.  Tile order = block order (a tile only ever waits
// for tiles with a smaller block index, which were dispatched before it); T3D_UPDATE_TICKET=1 hands the ids out
// with an atomic ticket instead.
//
// Numerics: --fmad=false, reference evaluation order (see t3d_kernels.cu).
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include "t3d_device.cuh"
#include "t3d_internal.h"

namespace t3d
{

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// status word: [63:34] launch epoch, [33:32] 1 = tile aggregate, 2 = inclusive prefix, [31:0] dead count
constexpr unsigned long long kStatusAggregate = 1ull << 32;
constexpr unsigned long long kStatusPrefix = 2ull << 32;

// V x 32-bit streaming accesses (evict-first: every byte is touched once per frame).
template <int V>
struct Vec;
template <>
struct Vec<4>
{
    static __device__ __forceinline__ void ld(const uint32_t* p, float* a)
    {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    }
    static __device__ __forceinline__ void ld(const uint32_t* p, int* a)
    {
        const int4 v = __ldcs(reinterpret_cast<const int4*>(p));
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    }
    static __device__ __forceinline__ void st(uint32_t* p, const float* a)
    {
        __stcs(reinterpret_cast<float4*>(p), make_float4(a[0], a[1], a[2], a[3]));
    }
    static __device__ __forceinline__ void st(uint32_t* p, const int* a)
    {
        __stcs(reinterpret_cast<int4*>(p), make_int4(a[0], a[1], a[2], a[3]));
    }
};
template <>
struct Vec<2>
{
    static __device__ __forceinline__ void ld(const uint32_t* p, float* a)
    {
        const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
        a[0] = v.x; a[1] = v.y;
    }
    static __device__ __forceinline__ void ld(const uint32_t* p, int* a)
    {
        const int2 v = __ldcs(reinterpret_cast<const int2*>(p));
        a[0] = v.x; a[1] = v.y;
    }
    static __device__ __forceinline__ void st(uint32_t* p, const float* a) { __stcs(reinterpret_cast<float2*>(p), make_float2(a[0], a[1])); }
    static __device__ __forceinline__ void st(uint32_t* p, const int* a) { __stcs(reinterpret_cast<int2*>(p), make_int2(a[0], a[1])); }
};
template <>
struct Vec<1>
{
    static __device__ __forceinline__ void ld(const uint32_t* p, float* a) { a[0] = __ldcs(reinterpret_cast<const float*>(p)); }
    static __device__ __forceinline__ void ld(const uint32_t* p, int* a) { a[0] = __ldcs(reinterpret_cast<const int*>(p)); }
    static __device__ __forceinline__ void st(uint32_t* p, const float* a) { __stcs(reinterpret_cast<float*>(p), a[0]); }
    static __device__ __forceinline__ void st(uint32_t* p, const int* a) { __stcs(reinterpret_cast<int*>(p), a[0]); }
};

// ---- register-free prefetch (ASYNC variants) -----------------------------------------------------------
// Each thread copies the 16 bytes it owns of every input column straight from global to shared memory with
// cp.async (LDGSTS), one commit group per particle group, two groups in flight per thread.  The bytes in flight
// per SM are then bounded by shared memory (2 CTAs x 2 stages x 48 KB), not by registers.  A thread only ever
// reads back its own 16-byte slots, so no block barrier is needed: cp.async.wait_group is the only sync.
enum : int
{
    RC_EST = 0, RC_POS = 1, RC_VEL = 4, RC_ACC = 7, RC_SPIN = 10, RC_ANGLE = 11, RC_CS = 12, RC_CE = 16, RC_SS = 20,
    RC_SE = 21, RC_TF = 22, RC_FRAME = 23, RC_COUNT = 24
};
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// slot(c): this thread's float4 of ring column c in the given stage
template <int THREADS>
struct Ring
{
    float4* base; // [2][RC_COUNT][THREADS]
    uint32_t tid;
    __device__ __forceinline__ float4* slot(int stage, int c) const { return base + ((size_t)(stage * RC_COUNT + c) * THREADS + tid); }
    // animated: the sprite-step inputs; need_frame: the frame index alone (the fused draw reads it for the uv lookup)
    __device__ __forceinline__ void issue(int stage, const uint32_t* g, size_t stride, bool animated, bool need_frame = false) const
    {
        auto G = [&](int col) { return g + (size_t)col * stride; };
        cp_async16(slot(stage, RC_EST), G(T3D_COL_ENERGY_START));
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            cp_async16(slot(stage, RC_POS + k), G(T3D_COL_POSITION + k));
            cp_async16(slot(stage, RC_VEL + k), G(T3D_COL_VELOCITY + k));
            cp_async16(slot(stage, RC_ACC + k), G(T3D_COL_ACCELERATION + k));
        }
        cp_async16(slot(stage, RC_SPIN), G(T3D_COL_ROT_PER_PARTICLE_SPEED));
        cp_async16(slot(stage, RC_ANGLE), G(T3D_COL_ANGLE));
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            cp_async16(slot(stage, RC_CS + k), G(T3D_COL_COLOR_START + k));
            cp_async16(slot(stage, RC_CE + k), G(T3D_COL_COLOR_END + k));
        }
        cp_async16(slot(stage, RC_SS), G(T3D_COL_SIZE_START));
        cp_async16(slot(stage, RC_SE), G(T3D_COL_SIZE_END));
        if (animated) cp_async16(slot(stage, RC_TF), G(T3D_COL_TIME_ON_FRAME));
        if (animated || need_frame) cp_async16(slot(stage, RC_FRAME), G(T3D_COL_FRAME));
    }
};
__device__ __forceinline__ void unpack4(const float4& v, float* a) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
__device__ __forceinline__ void unpack4(const float4& v, int* a)
{
    a[0] = __float_as_int(v.x); a[1] = __float_as_int(v.y); a[2] = __float_as_int(v.z); a[3] = __float_as_int(v.w);
}

// What one thread holds of one V-particle group.  The record is processed in two halves -- motion (position,
// velocity, acceleration, angle) and appearance (colour, size, sprite frame) -- each loaded, computed and stored
// before the other is touched, so that only half of the record occupies registers at a time and more CTAs fit
// on an SM; a compiler fence between the halves keeps the loads from being hoisted together again.
template <int V>
struct Motion
{
    float px[V], py[V], pz[V], vx[V], vy[V], vz[V], ax[V], ay[V], az[V];
    float spin[V], ang[V];
    float rx[V], ry[V], rz[V], rs[V];
    __device__ __forceinline__ void load(const uint32_t* base, size_t stride, bool may_rotate)
    {
        auto col = [&](int c) -> const uint32_t* { return base + (size_t)c * stride; };
        Vec<V>::ld(col(T3D_COL_POSITION + 0), px);
        Vec<V>::ld(col(T3D_COL_POSITION + 1), py);
        Vec<V>::ld(col(T3D_COL_POSITION + 2), pz);
        Vec<V>::ld(col(T3D_COL_VELOCITY + 0), vx);
        Vec<V>::ld(col(T3D_COL_VELOCITY + 1), vy);
        Vec<V>::ld(col(T3D_COL_VELOCITY + 2), vz);
        Vec<V>::ld(col(T3D_COL_ACCELERATION + 0), ax);
        Vec<V>::ld(col(T3D_COL_ACCELERATION + 1), ay);
        Vec<V>::ld(col(T3D_COL_ACCELERATION + 2), az);
        Vec<V>::ld(col(T3D_COL_ROT_PER_PARTICLE_SPEED), spin);
        Vec<V>::ld(col(T3D_COL_ANGLE), ang);
        if (may_rotate) load_rotation(base, stride);
    }
    __device__ __forceinline__ void load_rotation(const uint32_t* base, size_t stride)
    {
        auto col = [&](int c) -> const uint32_t* { return base + (size_t)c * stride; };
        Vec<V>::ld(col(T3D_COL_ROTATION_AXIS + 0), rx);
        Vec<V>::ld(col(T3D_COL_ROTATION_AXIS + 1), ry);
        Vec<V>::ld(col(T3D_COL_ROTATION_AXIS + 2), rz);
        Vec<V>::ld(col(T3D_COL_ROTATION_SPEED), rs);
    }
    template <int THREADS>
    __device__ __forceinline__ void load_ring(const Ring<THREADS>& r, int stage) // V == 4
    {
        unpack4(*r.slot(stage, RC_POS + 0), px); unpack4(*r.slot(stage, RC_POS + 1), py); unpack4(*r.slot(stage, RC_POS + 2), pz);
        unpack4(*r.slot(stage, RC_VEL + 0), vx); unpack4(*r.slot(stage, RC_VEL + 1), vy); unpack4(*r.slot(stage, RC_VEL + 2), vz);
        unpack4(*r.slot(stage, RC_ACC + 0), ax); unpack4(*r.slot(stage, RC_ACC + 1), ay); unpack4(*r.slot(stage, RC_ACC + 2), az);
        unpack4(*r.slot(stage, RC_SPIN), spin);
        unpack4(*r.slot(stage, RC_ANGLE), ang);
    }
};
template <int V>
struct Appearance
{
    int est[V];
    float ss[V], se[V];
    float cs[4][V], ce[4][V];
    float tf[V];
    int fr[V];
    __device__ __forceinline__ void load(const uint32_t* base, size_t stride, bool animated, bool need_frame = false)
    {
        auto col = [&](int c) -> const uint32_t* { return base + (size_t)c * stride; };
        Vec<V>::ld(col(T3D_COL_ENERGY_START), est);
#pragma unroll
        for (int c = 0; c < 4; ++c)
        {
            Vec<V>::ld(col(T3D_COL_COLOR_START + c), cs[c]);
            Vec<V>::ld(col(T3D_COL_COLOR_END + c), ce[c]);
        }
        Vec<V>::ld(col(T3D_COL_SIZE_START), ss);
        Vec<V>::ld(col(T3D_COL_SIZE_END), se);
        if (animated) Vec<V>::ld(col(T3D_COL_TIME_ON_FRAME), tf);
        if (animated || need_frame) Vec<V>::ld(col(T3D_COL_FRAME), fr);
    }
    template <int THREADS>
    __device__ __forceinline__ void load_ring(const Ring<THREADS>& r, int stage, bool animated, bool need_frame = false) // V == 4
    {
        unpack4(*r.slot(stage, RC_EST), est);
#pragma unroll
        for (int c = 0; c < 4; ++c)
        {
            unpack4(*r.slot(stage, RC_CS + c), cs[c]);
            unpack4(*r.slot(stage, RC_CE + c), ce[c]);
        }
        unpack4(*r.slot(stage, RC_SS), ss);
        unpack4(*r.slot(stage, RC_SE), se);
        if (animated) unpack4(*r.slot(stage, RC_TF), tf);
        if (animated || need_frame) unpack4(*r.slot(stage, RC_FRAME), fr);
    }
};
__device__ __forceinline__ void compiler_fence() { asm volatile("" ::: "memory"); }

// V = particles per thread per group, THREADS per CTA, SUB = groups per thread: a tile is V*THREADS*SUB
// consecutive particles of one emitter and costs one look-back, so the scan frontier only has to advance
// one tile per V*THREADS*SUB particles streamed.
//
// FUSED (with ASYNC, V = 4): the same pass also emits the frame's billboard quads (what k_draw would produce from
// the state this kernel leaves), so position/colour/size/angle/frame never make the round trip through HBM
// between update and draw: 304 B per particle per frame instead of 344.  Each thread expands the 4 particles of
// its group into the warp's piece of shared memory (128 consecutive quads, 18 KB), and one lane hands that piece to
// the TMA engine with a cp.async.bulk -- warp-level sync only on the vertex path.  Slots that receive a moved
// particle get their quad in the move pass; slots that are not examined lie beyond the new live count and are not
// part of the frame's vertex stream.
template <int V, int THREADS, int SUB, int MIN_BLOCKS, bool ASYNC, bool TICKET, bool FUSED = false>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
    k_update(const UpdateItem* __restrict__ items, uint32_t n_items, uint32_t total_tiles,
             unsigned long long* __restrict__ tile_status, uint32_t* __restrict__ ticket, uint32_t ticket_base,
             uint32_t epoch, uint32_t* __restrict__ fault, const DrawItem* __restrict__ ditems = nullptr, int rot_mode = 0,
             const FrozenRotation* __restrict__ frozen = nullptr)
{
    static_assert(!FUSED || !ASYNC || V == 4, "the cp.async ring moves 4 particles per thread");
    static_assert(V * THREADS * SUB < 65535, "tile-local offsets and ranks are kept as 16-bit values");
    constexpr int kQuadBytes = T3D_FLOATS_PER_QUAD * 4;
    constexpr int kStageWarp = 32 * V * kQuadBytes; // bytes of vertex staging per warp: its 32 * V consecutive quads
    __shared__ __align__(16) float s_uv[FUSED ? kMaxSmemFrames * 4 : 4];
    constexpr int GROUP = V * THREADS;  // particles per sub-tile
    constexpr int TILE = GROUP * SUB;
    constexpr int WARPS = THREADS / 32;
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[SUB][WARPS];
    __shared__ uint32_t s_excl;
    // Slots of this tile that receive a moved particle, indexed by the dead particle's rank inside the tile
    // (offset from the tile's first slot; kNoMove: nothing to copy).
    __shared__ unsigned short s_move[TILE];
    constexpr unsigned short kNoMove = 0xFFFFu;
    // Per-thread results of phase A (decremented energies, dead particles preceding each group), parked in shared
    // memory so that phase B can be a rolled loop: unrolling it SUB times overflowed the instruction cache.
    __shared__ __align__(16) int s_en[SUB * THREADS * V];
    __shared__ unsigned short s_base[SUB * THREADS]; // < TILE

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    // Tile order = block order.  The look-back below only ever waits for tiles with a smaller id; like CUB's
    // single-pass scan this relies on CTAs being dispatched in increasing blockIdx order.  TICKET hands the ids out
    // with an atomic instead (an extra round trip and a barrier before the first load can be issued).
    uint32_t tile = blockIdx.x;
    if (TICKET)
    {
        if (tid == 0) s_tile = atomicAdd(ticket, 1u) - ticket_base;
        __syncthreads();
        tile = s_tile;
    }
    if (tile >= total_tiles) return;

    // One dependent fetch (the launch descriptor) and the loads can start; the live count is fetched in parallel.
    const uint32_t item_index = find_item(items, n_items, tile, total_tiles);
    const UpdateItem& it = items[item_index];
    EmitterDev* const d = it.dev;
    const uint32_t parity = it.parity;
    uint32_t* const cols = it.cols;
    const size_t stride = it.stride;
    const uint32_t bshift = it.block_shift;
    const uint32_t cap = it.seg_capacity;
    const uint32_t jt = tile - it.tile_begin; // tile index inside the emitter
    const uint32_t first = jt * TILE;
    const uint32_t flags = it.flags;
    const bool animated = (flags & kFlagAnimated) != 0;
    const bool looped = (flags & kFlagLooped) != 0;
    const bool may_rotate = (flags & kFlagMayRotate) != 0;
    const float elapsed_ms = it.elapsed_ms;
    const float secs = elapsed_ms * 0.001f; // PE.cpp:727
    const uint32_t N = d->cnt[parity].count;

    // ---- phase A: the energy column of the whole tile -> dead flags (PE.cpp:753-755: `long -= float`, `> 0`) ----
    // Issued before the live count has arrived: every slot below the padded segment size is safe to read.
    int en[SUB][V];
#pragma unroll
    for (int s = 0; s < SUB; ++s)
    {
        const uint32_t k0 = first + s * GROUP + tid * V;
        if (k0 < cap) Vec<V>::ld(particle_base(cols, bshift, k0) + (size_t)T3D_COL_ENERGY * stride, en[s]);
    }
    // The first group(s) are requested now so that they are in flight while the scan and look-back run.
    Motion<V> m;
    // The ring starts on a 128-byte boundary so that a warp's 512-byte cp.async destination covers whole lines.
    // (The offset is added to the array itself so that the compiler keeps addressing it as shared memory.)
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    unsigned char* const ring_bytes = dyn_smem + ((128u - ((uint32_t)__cvta_generic_to_shared(dyn_smem) & 127u)) & 127u);
    Ring<THREADS> ring{ reinterpret_cast<float4*>(ring_bytes), tid };
    if (ASYNC)
    {
#pragma unroll
        for (int s = 0; s < (SUB < 2 ? SUB : 2); ++s)
        {
            const uint32_t k0 = first + s * GROUP + tid * V;
            if (k0 < cap) ring.issue(s, particle_base(cols, bshift, k0), stride, animated, FUSED);
            cp_async_commit();
        }
    }
    // With one or two particles per thread the whole record fits in registers twice: both halves are requested at
    // once and one group ahead (one exposed round trip to HBM per tile instead of two per group).
    constexpr bool kWholeRecord = FUSED && !ASYNC && V <= 2;
    Appearance<V> a_cur;
    if (!ASYNC)
    {
        const uint32_t k0 = first + tid * V;
        if (k0 < cap)
        {
            m.load(particle_base(cols, bshift, k0), stride, may_rotate);
            if (kWholeRecord) a_cur.load(particle_base(cols, bshift, k0), stride, animated, FUSED);
        }
    }
    // The host sized this emitter's share of the grid from an upper bound of its live count; if the bound was wrong,
    // say so instead of silently leaving particles out of the frame.
    if (jt == 0 && tid == 0 && (unsigned long long)N > (unsigned long long)it.n_tiles * TILE) atomicOr(fault, kFaultUpdateGrid);
    if (first >= N)
    {
        if (ASYNC) cp_async_wait<0>(); // nothing of this thread may still be landing in shared memory when the CTA retires
        if (jt == 0 && tid == 0)
        {
            d->cnt[parity ^ 1].count = 0;
            d->cnt[parity ^ 1].serial = d->cnt[parity].serial;
            if (FUSED) d->cnt[0].drawn = 0;
        }
        return;
    }
    // the draw side of the launch descriptor (FUSED)
    float right[3] = { 0, 0, 0 }, up[3] = { 0, 0, 0 };
    float* vtx = nullptr;
    unsigned char* stage_warp = ring_bytes + (ASYNC ? 2 * (size_t)RC_COUNT * THREADS * sizeof(float4) : 0) + (size_t)warp * kStageWarp;
    Rot3 frozen_rot = {};
    if (FUSED)
    {
        frozen_rot = load_frozen(frozen);
        const DrawItem& di = ditems[item_index];
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            right[k] = di.right[k];
            up[k] = di.up[k];
        }
        vtx = di.vtx;
        // the host only fuses emitters whose uv table fits (frame_count <= kMaxSmemFrames)
        const float* uv_g = di.uv;
        const uint32_t uv_words = (it.frame_count < (uint32_t)kMaxSmemFrames ? it.frame_count : (uint32_t)kMaxSmemFrames) * 4u;
        for (uint32_t q = tid; q < uv_words; q += THREADS) s_uv[q] = uv_g[q]; // visible after the scan's barrier
    }
    uint32_t cnt[SUB];
#pragma unroll
    for (int s = 0; s < SUB; ++s)
    {
        const uint32_t k0 = first + s * GROUP + tid * V;
        cnt[s] = 0;
#pragma unroll
        for (int l = 0; l < V; ++l)
        {
            if (k0 + l < N)
            {
                en[s][l] = __float2int_rz((float)en[s][l] - elapsed_ms);
                if (!(en[s][l] > 0)) cnt[s]++;
            }
        }
    }

    // ---- exclusive scan of the dead counts inside the tile (particle order: sub-tile, thread, lane) ----
    uint32_t incl[SUB];
#pragma unroll
    for (int s = 0; s < SUB; ++s)
    {
        uint32_t v = cnt[s];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= (uint32_t)o) v += t;
        }
        incl[s] = v;
        if (lane == 31) s_warp[s][warp] = v;
    }
    __syncthreads();
    uint32_t base[SUB]; // dead particles of this tile that precede this thread's group s
    uint32_t tile_dead = 0;
#pragma unroll
    for (int s = 0; s < SUB; ++s)
    {
        uint32_t before = tile_dead;
#pragma unroll
        for (int wi = 0; wi < WARPS; ++wi)
        {
            const uint32_t v = s_warp[s][wi];
            if ((uint32_t)wi < warp) before += v;
            tile_dead += v;
        }
        base[s] = before + incl[s] - cnt[s];
        s_base[s * THREADS + tid] = (unsigned short)base[s];
#pragma unroll
        for (int l = 0; l < V; ++l) s_en[(s * THREADS + tid) * V + l] = en[s][l];
    }

    // A tile in the first half of the emitter needs no prefix to know that every slot of it is examined:
    // D(k) <= k, so k + D(k) <= 2k < N.  Its warps start streaming at once; the prefix (resolved by warp 0 below,
    // which then joins them) is only needed for the sources of the moves, after the barrier at the end.
    const bool certain = 2ull * ((unsigned long long)first + TILE) <= (unsigned long long)N;
    // Ranks whose dead particle turns out not to be examined (or to be the last one) keep kNoMove.  Only an uncertain
    // tile can have such ranks, and only there a barrier separates this initialisation from the writes of phase B:
    // in a certain tile every rank is written in phase B, and initialising it here would race with that write.
    if (!certain)
        for (uint32_t r = tid; r < tile_dead; r += THREADS) s_move[r] = kNoMove;

    // ---- decoupled look-back over the preceding tiles of this emitter ----
    const unsigned long long tag = (unsigned long long)epoch << 34;
    unsigned long long* my_status = tile_status + tile;
    if (jt == 0)
    {
        if (tid == 0)
        {
            st_status(my_status, tag | kStatusPrefix | tile_dead);
            s_excl = 0;
        }
    }
    else if (warp == 0)
    {
        if (lane == 0) st_status(my_status, tag | kStatusAggregate | tile_dead);
        uint32_t excl = 0;
        int look = (int)jt - 1; // nearest predecessor, as a tile index inside the emitter
        while (true)
        {
            const int idx = look - (int)lane;
            unsigned long long s = tag | kStatusPrefix; // before the emitter's first tile: prefix 0
            if (idx >= 0)
            {
                const unsigned long long* sp = tile_status + (it.tile_begin + (uint32_t)idx);
                do
                {
                    s = ld_status(sp);
                } while ((s >> 34) != (unsigned long long)epoch || ((s >> 32) & 3ull) == 0ull);
            }
            const bool is_prefix = ((s >> 32) & 3ull) == 2ull;
            const unsigned pmask = __ballot_sync(0xffffffffu, is_prefix);
            const int first_p = __ffs((int)pmask) - 1;
            uint32_t contrib = (pmask == 0u || (int)lane <= first_p) ? (uint32_t)s : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            excl += contrib;
            if (pmask != 0u) break;
            look -= 32;
        }
        if (lane == 0)
        {
            st_status(my_status, tag | kStatusPrefix | (excl + tile_dead));
            s_excl = excl;
        }
    }
    if (!certain) __syncthreads();
    const uint32_t tile_excl = certain ? 0u : s_excl; // certain: ranks below are tile-local

    // ---- phase B: stream the groups ----
    float ppf = it.percent_per_frame;
    float dur = it.frame_duration_secs;
    uint32_t frame_count = it.frame_count;
    if (kWholeRecord)
    {
        // The group loop keeps a whole record of loads in flight.  A value that is still "a pending load" when the loop
        // is entered would make its consumers inside the loop wait on the scoreboard those loads share (ptxas puts
        // nearly all of them on one), i.e. for the record that was just requested.  Passing the per-tile constants
        // through one ALU instruction here settles them before the loop (the operand is 0 in this launch, unknown to
        // the compiler).
        auto settle_f = [&](float& v) { v = __uint_as_float(__float_as_uint(v) ^ ticket_base); };
        auto settle_u = [&](uint32_t& v) { v ^= ticket_base; };
        settle_f(ppf);
        settle_f(dur);
        settle_u(frame_count);
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            settle_f(right[k]);
            settle_f(up[k]);
        }
        settle_f(frozen_rot.m0); settle_f(frozen_rot.m1); settle_f(frozen_rot.m2);
        settle_f(frozen_rot.m4); settle_f(frozen_rot.m5); settle_f(frozen_rot.m6);
        settle_f(frozen_rot.m8); settle_f(frozen_rot.m9); settle_f(frozen_rot.m10);
    }
    // With one or two particles per thread the whole record fits in registers twice: both halves are requested at
    // once and one group ahead (one exposed round trip to HBM per tile instead of two per group).
#pragma unroll 1
    for (int s = 0; s < SUB; ++s)
    {
        const uint32_t k0 = first + s * GROUP + tid * V;
        // (FUSED: warps leave together -- the vertex staging below is warp-collective)
        if ((FUSED ? first + s * GROUP + (tid & ~31u) * V : k0) >= N) break;
        int e4[V];
#pragma unroll
        for (int l = 0; l < V; ++l) e4[l] = s_en[(s * THREADS + tid) * V + l];
        const uint32_t base_s = s_base[s * THREADS + tid];
        uint32_t* const gbase = particle_base(cols, bshift, k0);
        auto col = [&](int c) -> uint32_t* { return gbase + (size_t)c * stride; };

        // which originals the reference's loop examines, and what lands in each slot
        // all_store: every slot of the group is examined, so the whole group is stored with vector stores --
        // the dead ones with don't-care values, overwritten after the barrier below by the particle moved in.
        bool valid[V], alive[V], examined[V];
        uint32_t D[V];
        bool all_store = true;
        {
            uint32_t run = tile_excl + base_s;
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                valid[l] = k0 + l < N;
                alive[l] = valid[l] && e4[l] > 0;
                D[l] = run;
                examined[l] = valid[l] && (certain || k0 + l + run < N);
                if (examined[l] && !alive[l])
                {
                    // PE.cpp:825-828: the last live particle, untouched this frame, takes the dead one's place
                    // (unless the dead one IS the last: only possible in an uncertain tile).
                    const uint32_t src = N - 1u - run;
                    if (certain || src != k0 + l) s_move[run - tile_excl] = (unsigned short)(k0 + l - first);
                }
                run += (valid[l] && !alive[l]) ? 1u : 0u;
                all_store = all_store && examined[l];
            }
        }

        // kWholeRecord: this group's record is already in registers; the next group's is requested now, so that it
        // travels while this one is computed and expanded.
        Appearance<V> a;
        Motion<V> m_nxt;
        Appearance<V> a_nxt;
        if (kWholeRecord)
        {
            a = a_cur;
            const uint32_t kn = k0 + GROUP;
            if (s + 1 < SUB && kn < cap)
            {
                const uint32_t* nbase = particle_base(cols, bshift, kn);
                m_nxt.load(nbase, stride, may_rotate);
                a_nxt.load(nbase, stride, animated, FUSED);
            }
        }

        // -- motion half: PE.cpp:757-774 --
        if (ASYNC)
        {
            if (s + 1 < SUB)
                cp_async_wait<1>(); // group s has landed; group s+1 may still be in flight
            else
                cp_async_wait<0>();
            m.load_ring(ring, s & 1);
            if (may_rotate) m.load_rotation(gbase, stride);
        }
        else if (!kWholeRecord && s > 0)
            m.load(gbase, stride, may_rotate);
#pragma unroll
        for (int l = 0; l < V; ++l)
        {
            if (!alive[l]) continue;
            if (may_rotate && m.rs[l] != 0.0f && !(m.rx[l] == 0.0f && m.ry[l] == 0.0f && m.rz[l] == 0.0f))
            {
                const Rot3 r = create_rotation(m.rx[l], m.ry[l], m.rz[l], m.rs[l] * secs); // :759
                rot_apply(r, 1.0f, m.vx[l], m.vy[l], m.vz[l]);                              // :761
                rot_apply(r, 1.0f, m.ax[l], m.ay[l], m.az[l]);                              // :762
            }
            m.vx[l] += m.ax[l] * secs; // :766-768
            m.vy[l] += m.ay[l] * secs;
            m.vz[l] += m.az[l] * secs;
            m.px[l] += m.vx[l] * secs; // :770-772
            m.py[l] += m.vy[l] * secs;
            m.pz[l] += m.vz[l] * secs;
            m.ang[l] += m.spin[l] * secs; // :774
        }
        if (all_store)
        {
            Vec<V>::st(col(T3D_COL_VELOCITY + 0), m.vx);
            Vec<V>::st(col(T3D_COL_VELOCITY + 1), m.vy);
            Vec<V>::st(col(T3D_COL_VELOCITY + 2), m.vz);
            Vec<V>::st(col(T3D_COL_POSITION + 0), m.px);
            Vec<V>::st(col(T3D_COL_POSITION + 1), m.py);
            Vec<V>::st(col(T3D_COL_POSITION + 2), m.pz);
            Vec<V>::st(col(T3D_COL_ANGLE), m.ang);
            if (may_rotate)
            {
                Vec<V>::st(col(T3D_COL_ACCELERATION + 0), m.ax);
                Vec<V>::st(col(T3D_COL_ACCELERATION + 1), m.ay);
                Vec<V>::st(col(T3D_COL_ACCELERATION + 2), m.az);
            }
        }
        else
        {
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                if (!(examined[l] && alive[l])) continue;
                auto S = [&](int c, float v) { col(c)[l] = __float_as_uint(v); };
                S(T3D_COL_VELOCITY + 0, m.vx[l]);
                S(T3D_COL_VELOCITY + 1, m.vy[l]);
                S(T3D_COL_VELOCITY + 2, m.vz[l]);
                S(T3D_COL_POSITION + 0, m.px[l]);
                S(T3D_COL_POSITION + 1, m.py[l]);
                S(T3D_COL_POSITION + 2, m.pz[l]);
                S(T3D_COL_ANGLE, m.ang[l]);
                if (may_rotate)
                {
                    S(T3D_COL_ACCELERATION + 0, m.ax[l]);
                    S(T3D_COL_ACCELERATION + 1, m.ay[l]);
                    S(T3D_COL_ACCELERATION + 2, m.az[l]);
                }
            }
        }
        if (!kWholeRecord) compiler_fence();

        // -- appearance half: PE.cpp:777-819 --
        if (ASYNC)
        {
            a.load_ring(ring, s & 1, animated, FUSED);
            // This stage's slots are consumed: refill them with the group after next.
            if (s + 2 < SUB)
            {
                const uint32_t kn = first + (s + 2) * GROUP + tid * V;
                if (kn < N) ring.issue(s & 1, particle_base(cols, bshift, kn), stride, animated, FUSED);
                cp_async_commit();
            }
        }
        else if (!kWholeRecord)
            a.load(gbase, stride, animated, FUSED);
        float col_out[4][V], size_out[V];
#pragma unroll
        for (int l = 0; l < V; ++l)
        {
            if (!alive[l]) continue;
            const float percent = 1.0f - ((float)e4[l] / (float)a.est[l]); // :777
#pragma unroll
            for (int c = 0; c < 4; ++c) col_out[c][l] = a.cs[c][l] + (a.ce[c][l] - a.cs[c][l]) * percent; // :779-782
            size_out[l] = a.ss[l] + (a.se[l] - a.ss[l]) * percent;                                          // :784
            if (animated)
            {
                if (!looped)
                {
                    float spent = 0.0f; // :792-796, repeated addition
                    for (int q = 0; q < a.fr[l]; ++q) spent += ppf;
                    a.tf[l] = percent - spent;
                    if ((uint32_t)a.fr[l] < frame_count - 1u && a.tf[l] >= ppf) ++a.fr[l];
                }
                else
                {
                    a.tf[l] += secs; // :808-817
                    if (a.tf[l] >= dur)
                    {
                        a.tf[l] -= dur;
                        ++a.fr[l];
                        if ((uint32_t)a.fr[l] == frame_count) a.fr[l] = 0;
                    }
                }
            }
        }
        if (all_store)
        {
            Vec<V>::st(col(T3D_COL_ENERGY), e4);
#pragma unroll
            for (int c = 0; c < 4; ++c) Vec<V>::st(col(T3D_COL_COLOR + c), col_out[c]);
            Vec<V>::st(col(T3D_COL_SIZE), size_out);
            if (animated)
            {
                Vec<V>::st(col(T3D_COL_TIME_ON_FRAME), a.tf);
                Vec<V>::st(col(T3D_COL_FRAME), a.fr);
            }
        }
        else
        {
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                if (!(examined[l] && alive[l])) continue;
                auto S = [&](int c, float v) { col(c)[l] = __float_as_uint(v); };
                col(T3D_COL_ENERGY)[l] = (uint32_t)e4[l];
#pragma unroll
                for (int c = 0; c < 4; ++c) S(T3D_COL_COLOR + c, col_out[c][l]);
                S(T3D_COL_SIZE, size_out[l]);
                if (animated)
                {
                    S(T3D_COL_TIME_ON_FRAME, a.tf[l]);
                    col(T3D_COL_FRAME)[l] = (uint32_t)a.fr[l];
                }
            }
        }

        if (FUSED)
        {
            // -- the quads of this group: PE.cpp:866-882 + SB.cpp:292-363 on the values just computed --
            // Examined slots form a prefix of the group (k + D(k) is increasing); the rest lie beyond the new count.
            uint32_t n_quads = 0;
#pragma unroll
            for (int l = 0; l < V; ++l) n_quads += examined[l] ? 1u : 0u;
            const uint32_t warp_quads = __reduce_add_sync(0xffffffffu, n_quads);
            if (warp_quads)
            {
                // the warp's piece of shared memory is reused: its previous bulk store must have read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int l = 0; l < V; ++l)
                {
                    if (!(examined[l] && alive[l])) continue; // dead: the move pass writes this quad
                    const uint32_t frame = (uint32_t)a.fr[l] < frame_count ? (uint32_t)a.fr[l] : frame_count - 1u;
                    // (always from shared memory: a global load here would share a scoreboard with the record in flight)
                    const float4 uv = *reinterpret_cast<const float4*>(s_uv + frame * 4u);
                    const float pos[3] = { m.px[l], m.py[l], m.pz[l] };
                    const float rgba[4] = { col_out[0][l], col_out[1][l], col_out[2][l], col_out[3][l] };
                    expand_quad(pos, rgba, size_out[l], m.ang[l], uv.x, uv.y, uv.z, uv.w, right, up, rot_mode, [&] { return frozen_rot; },
                                reinterpret_cast<float4*>(stage_warp + (lane * V + l) * kQuadBytes));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                {
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(stage_warp);
                    const uint32_t bytes = warp_quads * (uint32_t)kQuadBytes;
                    float* dst = vtx + (size_t)(first + s * GROUP + warp * 32u * V) * T3D_FLOATS_PER_QUAD;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }

        // the last examined original publishes the new count (PE.cpp:829); it never lies in a certain tile
#pragma unroll
        for (int l = 0; l < V; ++l)
        {
            if (certain || !examined[l]) continue;
            const uint32_t dead_l = alive[l] ? 0u : 1u;
            const bool next_examined = (k0 + l + 1u) + D[l] + dead_l < N;
            if (!next_examined)
            {
                d->cnt[parity ^ 1].count = N - D[l] - dead_l;
                d->cnt[parity ^ 1].serial = d->cnt[parity].serial;
                if (FUSED) d->cnt[0].drawn = N - D[l] - dead_l;
            }
        }
        if (kWholeRecord)
        {
            m = m_nxt;
            a_cur = a_nxt;
        }
        compiler_fence();
    }
    // (threads that left the loop early still have their first ring groups in flight)
    if (ASYNC) cp_async_wait<0>();
    // shared memory must outlive the bulk stores, and their writes must have landed before the move pass overwrites
    // the quads of the slots it fills
    if (FUSED && lane == 0)
    {
        if (tile_dead)
        {
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        else // nothing is overwritten afterwards: it is enough that the staging memory has been read
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }

    // ---- the moves: rank r of this tile's dead particles pulls original N-1-(tile_excl+r) into its slot.  One thread
    // per move (consecutive lanes read consecutive sources, contiguous at the end of the live range, so each column
    // is a coalesced read).  The moves are latency-bound, not bandwidth-bound: a thread requests the 34 words of two
    // moves (68 independent loads) before it stores any, so a tile with up to 2 x THREADS dead particles pays one round
    // trip.  Sources lie beyond every examined slot: nothing read here is written by anyone.
    if (tile_dead)
    {
        __syncthreads(); // s_move complete; this CTA's stores to the dead slots are ordered before the moves
        const uint32_t src_last = N - 1u - s_excl; // the emitter-wide prefix, visible to all after the barrier
        if (!FUSED)
        {
            constexpr int PER = 2;
            for (uint32_t r0 = tid; r0 < tile_dead; r0 += PER * THREADS)
            {
                uint32_t val[PER][T3D_NUM_COLUMNS];
                uint32_t* dst[PER];
#pragma unroll
                for (int p = 0; p < PER; ++p)
                {
                    const uint32_t r = r0 + p * THREADS;
                    const unsigned short off = r < tile_dead ? s_move[r] : kNoMove;
                    dst[p] = nullptr;
                    if (off == kNoMove) continue;
                    dst[p] = particle_base(cols, bshift, first + off);
                    const uint32_t* const src = particle_base(cols, bshift, src_last - r);
#pragma unroll
                    for (int j = 0; j < T3D_NUM_COLUMNS; ++j) val[p][j] = src[(size_t)j * stride];
                }
#pragma unroll
                for (int p = 0; p < PER; ++p)
                {
                    if (!dst[p]) continue;
#pragma unroll
                    for (int j = 0; j < T3D_NUM_COLUMNS; ++j) dst[p][(size_t)j * stride] = val[p][j];
                }
            }
        }
        // FUSED: the same two moves per thread; the moved particle is also drawn, as it stands (PE.cpp:866-882)
        for (uint32_t r0 = tid; FUSED && r0 < tile_dead; r0 += 2 * THREADS)
        {
            uint32_t val[2][T3D_NUM_COLUMNS];
            uint32_t slot[2];
#pragma unroll
            for (int p = 0; p < 2; ++p)
            {
                const uint32_t r = r0 + p * THREADS;
                const unsigned short off = r < tile_dead ? s_move[r] : kNoMove;
                slot[p] = 0xffffffffu;
                if (off == kNoMove) continue;
                slot[p] = first + off;
                const uint32_t* const src = particle_base(cols, bshift, src_last - r);
#pragma unroll
                for (int j = 0; j < T3D_NUM_COLUMNS; ++j) val[p][j] = src[(size_t)j * stride];
            }
#pragma unroll
            for (int p = 0; p < 2; ++p)
            {
                if (slot[p] == 0xffffffffu) continue;
                uint32_t* const dst = particle_base(cols, bshift, slot[p]);
#pragma unroll
                for (int j = 0; j < T3D_NUM_COLUMNS; ++j) dst[(size_t)j * stride] = val[p][j];
                auto F = [&](int c) { return __uint_as_float(val[p][c]); };
                const uint32_t frame = val[p][T3D_COL_FRAME] < frame_count ? val[p][T3D_COL_FRAME] : frame_count - 1u;
                const float4 uv = *reinterpret_cast<const float4*>(s_uv + frame * 4u);
                const float pos[3] = { F(T3D_COL_POSITION), F(T3D_COL_POSITION + 1), F(T3D_COL_POSITION + 2) };
                const float rgba[4] = { F(T3D_COL_COLOR), F(T3D_COL_COLOR + 1), F(T3D_COL_COLOR + 2), F(T3D_COL_COLOR + 3) };
                float4 q[9];
                expand_quad(pos, rgba, F(T3D_COL_SIZE), F(T3D_COL_ANGLE), uv.x, uv.y, uv.z, uv.w, right, up, rot_mode, [&] { return frozen_rot; }, q);
                float4* const o = reinterpret_cast<float4*>(vtx + (size_t)slot[p] * T3D_FLOATS_PER_QUAD);
#pragma unroll
                for (int j = 0; j < 9; ++j) __stcs(o + j, q[j]);
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------
// variants: {particles per thread per group, threads per CTA, groups per thread, min CTAs per SM}.  The
// default is chosen from measurements on B200 (profiles/); T3D_UPDATE_VARIANT selects another for experiments.
// --------------------------------------------------------------------------------------------------
struct UpdateVariant
{
    const char* name;
    int tile;
    void (*launch)(cudaStream_t, uint32_t, const UpdateItem*, uint32_t, unsigned long long*, uint32_t*, uint32_t, uint32_t, uint32_t*);
};

constexpr int kMaxDevices = 64;
static int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

// T3D_UPDATE_TICKET=1: hand tile ids out with an atomic ticket instead of relying on block dispatch order.
static bool use_ticket()
{
    static int v = -1;
    if (v < 0)
    {
        const char* env = getenv("T3D_UPDATE_TICKET");
        v = (env && env[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

template <int V, int THREADS, int SUB, int MIN_BLOCKS>
static void launch_variant(cudaStream_t s, uint32_t tiles, const UpdateItem* items, uint32_t n_items,
                           unsigned long long* status, uint32_t* ticket, uint32_t ticket_base, uint32_t epoch, uint32_t* fault)
{
    if (use_ticket())
        k_update<V, THREADS, SUB, MIN_BLOCKS, false, true><<<tiles, THREADS, 0, s>>>(items, n_items, tiles, status, ticket, ticket_base, epoch, fault);
    else
        k_update<V, THREADS, SUB, MIN_BLOCKS, false, false><<<tiles, THREADS, 0, s>>>(items, n_items, tiles, status, ticket, ticket_base, epoch, fault);
}

template <int THREADS, int SUB, int MIN_BLOCKS>
static void launch_async(cudaStream_t s, uint32_t tiles, const UpdateItem* items, uint32_t n_items,
                         unsigned long long* status, uint32_t* ticket, uint32_t ticket_base, uint32_t epoch, uint32_t* fault)
{
    constexpr size_t smem = 2 * (size_t)RC_COUNT * THREADS * sizeof(float4) + 128;
    // the opt-in to more than 48 KB of dynamic shared memory is per device: once for each device this process launches on
    static bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (!configured[dev])
    {
        cudaFuncSetAttribute(k_update<4, THREADS, SUB, MIN_BLOCKS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev] = true;
    }
    k_update<4, THREADS, SUB, MIN_BLOCKS, true, false><<<tiles, THREADS, smem, s>>>(items, n_items, tiles, status, ticket, ticket_base, epoch, fault);
}

#define T3D_VARIANT(V, T, S, B) { "v" #V "t" #T "s" #S "b" #B, V * T * S, launch_variant<V, T, S, B> }
#define T3D_ASYNC(T, S, B) { "a4t" #T "s" #S "b" #B, 4 * T * S, launch_async<T, S, B> }
static const UpdateVariant kVariants[] = {
    // default: 96 threads x 4 particles x 11 groups = 4 224-particle tiles, cp.async ring, 2 CTAs/SM (shared-memory
    // bound; 12 groups no longer fit twice).  Measured on B200 (profiles/r1_update_variant_sweep*.txt).
    T3D_ASYNC(96, 11, 2),
    // alternatives kept for experiments and for the parity tests of the other code paths (tests/test_gpu_variants.py)
    T3D_ASYNC(96, 8, 2), T3D_ASYNC(96, 10, 2), T3D_ASYNC(64, 10, 3), T3D_ASYNC(64, 8, 3),
    T3D_VARIANT(4, 128, 4, 3), T3D_VARIANT(4, 128, 1, 4), T3D_VARIANT(4, 256, 4, 2), T3D_VARIANT(2, 256, 4, 4),
};
static int g_variant = -1;

static const UpdateVariant& variant()
{
    if (g_variant < 0)
    {
        g_variant = 0;
        if (const char* env = getenv("T3D_UPDATE_VARIANT"))
            for (int i = 0; i < (int)(sizeof(kVariants) / sizeof(kVariants[0])); ++i)
                if (!strcmp(env, kVariants[i].name)) g_variant = i;
    }
    return kVariants[g_variant];
}

// ---- fused update + draw ----------------------------------------------------------------------------------
struct FusedVariant
{
    const char* name;
    int tile;
    void (*launch)(cudaStream_t, uint32_t, const UpdateItem*, const DrawItem*, uint32_t, unsigned long long*, uint32_t, int,
                   const FrozenRotation*, uint32_t*);
};
template <int THREADS, int SUB, int MIN_BLOCKS>
static void launch_fused(cudaStream_t s, uint32_t tiles, const UpdateItem* items, const DrawItem* ditems, uint32_t n_items,
                         unsigned long long* status, uint32_t epoch, int rot_mode, const FrozenRotation* frozen, uint32_t* fault)
{
    // ring + vertex staging (4 quads per thread, contiguous per warp)
    constexpr size_t smem = 2 * (size_t)RC_COUNT * THREADS * sizeof(float4) + (size_t)THREADS * (4 * T3D_FLOATS_PER_QUAD * 4) + 128;
    // the opt-in to more than 48 KB of dynamic shared memory is per device: once for each device this process launches on
    static bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (!configured[dev])
    {
        cudaFuncSetAttribute(k_update<4, THREADS, SUB, MIN_BLOCKS, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev] = true;
    }
    k_update<4, THREADS, SUB, MIN_BLOCKS, true, false, true><<<tiles, THREADS, smem, s>>>(items, n_items, tiles, status, nullptr, 0u, epoch,
                                                                                          fault, ditems, rot_mode, frozen);
}
// the same without the cp.async ring: loads through registers, more warps per SM
template <int V, int THREADS, int SUB, int MIN_BLOCKS>
static void launch_fused_lsu(cudaStream_t s, uint32_t tiles, const UpdateItem* items, const DrawItem* ditems, uint32_t n_items,
                             unsigned long long* status, uint32_t epoch, int rot_mode, const FrozenRotation* frozen, uint32_t* fault)
{
    constexpr size_t smem = (size_t)THREADS * (V * T3D_FLOATS_PER_QUAD * 4) + 128;
    // the opt-in to more than 48 KB of dynamic shared memory is per device: once for each device this process launches on
    static bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (!configured[dev])
    {
        cudaFuncSetAttribute(k_update<V, THREADS, SUB, MIN_BLOCKS, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev] = true;
    }
    k_update<V, THREADS, SUB, MIN_BLOCKS, false, false, true><<<tiles, THREADS, smem, s>>>(items, n_items, tiles, status, nullptr, 0u, epoch,
                                                                                           fault, ditems, rot_mode, frozen);
}
#define T3D_FUSED(T, S, B) { "f4t" #T "s" #S "b" #B, 4 * T * S, launch_fused<T, S, B> }
#define T3D_FUSED_LSU(V, T, S, B) { "g" #V "t" #T "s" #S "b" #B, V * T * S, launch_fused_lsu<V, T, S, B> }
static const FusedVariant kFused[] = {
    // default: 2 particles per thread, whole record in registers one group ahead, 192 threads x 12 groups = 4 608-particle
    // tiles, 2 CTAs/SM.  Measured on B200 (profiles/r1_fused_sweep.txt): 3.30 ms for 64 Mi particles (0.95 of the HBM
    // peak on 304 B/particle) against 1.69 + 1.83 ms for the two separate launches.  The 4-particle variants (f4: with the
    // cp.async ring, g4: without) are capped at 4-8 warps per SM by their vertex staging and lose.
    T3D_FUSED_LSU(2, 192, 12, 2),
    // alternatives kept for experiments and for the parity tests of the other code paths (tests/test_gpu_variants.py)
    T3D_FUSED_LSU(2, 192, 8, 2), T3D_FUSED_LSU(2, 384, 6, 1), T3D_FUSED_LSU(4, 128, 8, 2), T3D_FUSED(64, 10, 2),
};
static const FusedVariant& fused_variant()
{
    static int v = -1;
    if (v < 0)
    {
        v = 0;
        if (const char* env = getenv("T3D_FUSED_VARIANT"))
            for (int i = 0; i < (int)(sizeof(kFused) / sizeof(kFused[0])); ++i)
                if (!strcmp(env, kFused[i].name)) v = i;
    }
    return kFused[v];
}
int fused_tile_size() { return fused_variant().tile; }
// An update of a set of emitters directly followed by the draw of the same set can go out as one launch.
// T3D_FUSED=1 always, T3D_FUSED=0 never, unset: when the runtime expects it to pay (see hold_for_fusion()).
int fused_mode()
{
    static int v = -2;
    if (v == -2)
    {
        const char* env = getenv("T3D_FUSED");
        v = !env ? -1 : (env[0] == '1' ? 1 : 0);
    }
    return v;
}
void launch_update_draw(void* stream, const UpdateItem* items, const DrawItem* ditems, uint32_t n_items, uint32_t total_tiles,
                        unsigned long long* tile_status, uint32_t epoch, int rot_mode, const FrozenRotation* frozen, uint32_t* fault)
{
    fused_variant().launch((cudaStream_t)stream, total_tiles, items, ditems, n_items, tile_status, epoch, rot_mode, frozen, fault);
}

int update_tile_size() { return variant().tile; }
const char* update_variant_name() { return variant().name; }

void launch_update(void* stream, const UpdateItem* items, uint32_t n_items, uint32_t total_tiles,
                   unsigned long long* tile_status, uint32_t* ticket, uint32_t ticket_base, uint32_t epoch, uint32_t* fault)
{
    variant().launch((cudaStream_t)stream, total_tiles, items, n_items, tile_status, ticket, ticket_base, epoch, fault);
}

} // namespace t3d
