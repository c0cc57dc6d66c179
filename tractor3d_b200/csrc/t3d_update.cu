// t3d_update.cu -- the update + dead-particle removal kernel (replaces the per-particle loop of
// ParticleEmitter::update, PE.cpp:750-831 of the reference) and its FUSED form, which also expands the frame's
// billboard quads (ParticleEmitter::draw + SpriteBatch::draw, PE.cpp:866-882, SB.cpp:292-363) in the same pass.
//
// Storage (t3d_internal.h): 15 read-write columns (struct-of-arrays) + one 64-byte life record per particle + one
// 16-byte rotation record.  One thread owns V consecutive particles of one emitter: it reads and writes each rw column
// with one V x 32-bit access (a warp touches 32 x V x 4 contiguous bytes per column per instruction); the warp fetches
// the 32 x V life records as ONE contiguous run with cp.async (512 bytes per instruction) into shared memory and every
// lane reads its own back.  The kernel is a pure HBM stream: 164 B per particle with sprite animation (40 B of rw
// columns + 64 B record read, 60 B written), the fused form adds the 144 B of vertices and saves the draw kernel's
// 40-byte re-read.
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
// emitter by decoupled look-back over per-tile status words.
//
// The move is folded into the stream: the thread that owns a dead slot fetches the mover's rw words INSTEAD of the
// slot's own (one group ahead, like every other load), so the mover's values leave with the group's vector stores and
// its quad with the warp's bulk store -- no scattered word stores, no second pass over the tile.  Only the mover's two
// records travel apart: 64 + 16 contiguous bytes, whole sectors.
//
// Tile order = block order (a tile only ever waits for tiles with a smaller block index, which were dispatched before
// it); T3D_UPDATE_TICKET=1 hands the ids out with an atomic ticket instead (every variant honours it).
//
// Numerics: --fmad=false, reference evaluation order (see t3d_kernels.cu).
#include <cuda_runtime.h>

#include <cstdio>
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

// The rw words a live particle's update reads (energy comes from phase A; colour and size are only written).
template <int V>
struct RwIn
{
    float px[V], py[V], pz[V], vx[V], vy[V], vz[V], ang[V], tf[V];
    int fr[V];
};

__device__ __forceinline__ void compiler_fence() { asm volatile("" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// V = particles per thread per group, THREADS per CTA, SUB = groups per thread: a tile is V*THREADS*SUB
// consecutive particles of one emitter and costs one look-back, so the scan frontier only has to advance
// one tile per V*THREADS*SUB particles streamed.
//
// FUSED: the same pass also emits the frame's billboard quads (what k_draw would produce from the state this kernel
// leaves), so position/colour/size/angle/frame never make the round trip through HBM between update and draw.
// Each thread expands its V particles into the warp's piece of shared memory (32 * V consecutive quads) and one lane
// hands that piece to the TMA engine with a cp.async.bulk -- warp-level sync only on the vertex path.  Slots that are
// not examined lie beyond the new live count and are not part of the frame's vertex stream.
//
// BOUNDS: also reduces the bounding box of pos +- size over the particles that are live after the update.
template <int V, int THREADS, int SUB, int MIN_BLOCKS, bool TICKET, bool FUSED, bool BOUNDS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
    k_update(const UpdateItem* __restrict__ items, uint32_t n_items, uint32_t total_tiles,
             unsigned long long* __restrict__ tile_status, uint32_t* __restrict__ ticket, uint32_t ticket_base,
             uint32_t epoch, uint32_t* __restrict__ fault, uint32_t zero, const DrawItem* __restrict__ ditems, int rot_mode,
             const FrozenRotation* __restrict__ frozen)
{
    static_assert(V == 2 || V == 4, "a thread owns 2 or 4 consecutive particles");
    static_assert(V * THREADS * SUB < 65535, "tile-local dead counts are kept as 16-bit values");
    constexpr int kQuadBytes = T3D_FLOATS_PER_QUAD * 4;
    constexpr int WG = 32 * V;                  // particles per warp per group
    constexpr int kRecWarp = WG * REC_WORDS * 4; // bytes of life records per warp per group
    constexpr int kStageWarp = WG * kQuadBytes;  // bytes of vertex staging per warp: its 32 * V consecutive quads
    constexpr int kSideLane = 48;                // bytes per lane of the mover side stage: rotation record + 6 rw words
    constexpr int kSideWarp = 32 * kSideLane;
    constexpr int GROUP = V * THREADS;  // particles per group
    constexpr int TILE = GROUP * SUB;
    constexpr int WARPS = THREADS / 32;
    __shared__ __align__(16) float s_uv[FUSED ? kMaxFusedFrames * 4 : 4];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[SUB][WARPS];
    __shared__ uint32_t s_excl;
    // Per-thread results of phase A (decremented energies, dead particles preceding each group), parked in shared
    // memory so that phase B can be a rolled loop: unrolling it SUB times overflowed the instruction cache.
    __shared__ __align__(16) int s_en[SUB * THREADS * V];
    __shared__ unsigned short s_base[SUB * THREADS]; // < TILE
    __shared__ uint32_t s_box[6];

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

    // One dependent fetch (the launch descriptor) and the loads can start; the live count is fetched in parallel.  The whole
    // entry of the emitter this tile is expected to belong to (exact when the emitters have equal tile counts, and for a single
    // emitter) is requested at once and checked; only a wrong guess pays for the search (dependent loads, one after another).
    auto load_item = [&](uint32_t i)
    {
        UpdateItem r;
        const uint4* src = reinterpret_cast<const uint4*>(items + i);
        uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
        for (int q = 0; q < (int)(sizeof(UpdateItem) / 16); ++q) dst[q] = __ldg(src + q);
        return r;
    };
    uint32_t item_index = guess_item(n_items, tile, total_tiles);
    UpdateItem it = load_item(item_index);
    if (tile - it.tile_begin >= it.n_tiles)
    {
        item_index = find_item(items, n_items, tile, total_tiles);
        it = load_item(item_index);
    }
    EmitterDev* const d = it.dev;
    const uint32_t parity = it.parity;
    uint32_t* const rw = it.st.rw;
    uint32_t* const rec = it.st.rec;
    uint32_t* const rot = it.st.rot;
    const size_t stride = it.st.stride;
    const uint32_t cap = it.st.seg_capacity;
    const uint32_t jt = tile - it.tile_begin; // tile index inside the emitter
    const uint32_t first = jt * TILE;
    const uint32_t flags = it.flags;
    const bool animated = (flags & kFlagAnimated) != 0;
    const bool looped = (flags & kFlagLooped) != 0;
    const bool may_rotate = (flags & kFlagMayRotate) != 0;
    const bool need_frame = animated || FUSED; // the fused draw reads the frame index for the uv lookup
    const float elapsed_ms = it.elapsed_ms;
    const float secs = elapsed_ms * 0.001f; // PE.cpp:727
    const uint32_t N = d->cnt[parity].count;
    auto col = [&](int c) -> uint32_t* { return rw + (size_t)c * stride; };

    extern __shared__ __align__(128) unsigned char dyn_smem[];
    // (the offset is added to the array itself so that the compiler keeps addressing it as shared memory)
    unsigned char* const dyn = dyn_smem + ((128u - ((uint32_t)__cvta_generic_to_shared(dyn_smem) & 127u)) & 127u);
    unsigned char* const rec_warp = dyn + (size_t)warp * kRecWarp;
    unsigned char* const stage_warp = dyn + (size_t)WARPS * kRecWarp + (size_t)warp * kStageWarp;
    // the words of a mover that the slot's thread only passes on (rotation record, energy, colour, size) land here,
    // requested one group ahead like everything else: [0,16) rotation record, [16,40) energy, colour x 4, size
    // 32 slots per warp, handed out per group by rank among the warp's movers (a thread's second mover takes a slot that another
    // lane does not need: a group holds ~5 movers per warp at 7 % of the particles replaced per frame)
    unsigned char* const side_warp = dyn + (size_t)WARPS * (kRecWarp + (FUSED ? kStageWarp : 0)) + (size_t)warp * kSideWarp;
    const uint32_t lanes_below = (1u << lane) - 1u;

    // ---- phase A: the energy column of the whole tile -> dead flags (PE.cpp:753-755: `long -= float`, `> 0`) ----
    // Issued before the live count has arrived: every slot below the padded segment size is safe to read.
    int en[SUB][V];
#pragma unroll
    for (int s = 0; s < SUB; ++s)
    {
        const uint32_t k0 = first + s * GROUP + tid * V;
        if (k0 < cap) Vec<V>::ld(col(RW_ENERGY) + k0, en[s]);
    }
    // Group 0 is requested now so that it is in flight while the scan and the look-back run: the warp's life records into
    // shared memory, the thread's own rw words into registers.
    {
        const uint32_t w0 = first + warp * WG;
        if (w0 < cap) rec_issue<V>(rec_warp, rec + (size_t)w0 * REC_WORDS, cap - w0 < (uint32_t)WG ? cap - w0 : (uint32_t)WG, lane);
        cp_async_commit();
    }
    auto load_own = [&](uint32_t k0, RwIn<V>& r)
    {
        Vec<V>::ld(col(RW_POS + 0) + k0, r.px);
        Vec<V>::ld(col(RW_POS + 1) + k0, r.py);
        Vec<V>::ld(col(RW_POS + 2) + k0, r.pz);
        Vec<V>::ld(col(RW_VEL + 0) + k0, r.vx);
        Vec<V>::ld(col(RW_VEL + 1) + k0, r.vy);
        Vec<V>::ld(col(RW_VEL + 2) + k0, r.vz);
        Vec<V>::ld(col(RW_ANGLE) + k0, r.ang);
        if (animated) Vec<V>::ld(col(RW_TIME_ON_FRAME) + k0, r.tf);
        if (need_frame) Vec<V>::ld(col(RW_FRAME) + k0, r.fr);
    };
    // the same words of ONE particle (a mover: the untouched original that takes a dead particle's slot)
    auto load_one = [&](uint32_t i, RwIn<V>& r, int l)
    {
        r.px[l] = ld_f(col(RW_POS + 0) + i);
        r.py[l] = ld_f(col(RW_POS + 1) + i);
        r.pz[l] = ld_f(col(RW_POS + 2) + i);
        r.vx[l] = ld_f(col(RW_VEL + 0) + i);
        r.vy[l] = ld_f(col(RW_VEL + 1) + i);
        r.vz[l] = ld_f(col(RW_VEL + 2) + i);
        r.ang[l] = ld_f(col(RW_ANGLE) + i);
        r.tf[l] = ld_f(col(RW_TIME_ON_FRAME) + i); // (a mover brings these two along even when the emitter does not animate)
        r.fr[l] = ld_i(col(RW_FRAME) + i);
    };
    auto prefetch_mover = [&](uint32_t i)
    {
        prefetch_l2(rec + (size_t)i * REC_WORDS);
        prefetch_l2(rot + (size_t)i * ROT_WORDS);
        prefetch_l2(col(RW_ENERGY) + i);
#pragma unroll
        for (int c = 0; c < 4; ++c) prefetch_l2(col(RW_COLOR + c) + i);
        prefetch_l2(col(RW_SIZE) + i);
    };
    RwIn<V> cur;
    {
        const uint32_t k0 = first + tid * V;
        if (k0 < cap) load_own(k0, cur);
    }
    // The host sized this emitter's share of the grid from an upper bound of its live count; if the bound was wrong,
    // say so instead of silently leaving particles out of the frame.
    if (jt == 0 && tid == 0 && (unsigned long long)N > (unsigned long long)it.n_tiles * TILE) atomicOr(fault, kFaultUpdateGrid);
    if (first >= N)
    {
        cp_async_wait<0>(); // nothing of this thread may still be landing in shared memory when the CTA retires
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
        // the host only fuses emitters whose uv table fits (frame_count <= kMaxFusedFrames)
        const float* uv_g = di.uv;
        const uint32_t uv_words = (it.frame_count < (uint32_t)kMaxFusedFrames ? it.frame_count : (uint32_t)kMaxFusedFrames) * 4u;
        for (uint32_t q = tid; q < uv_words; q += THREADS) s_uv[q] = uv_g[q]; // visible after the scan's barrier
    }
    if (BOUNDS && tid < 6) s_box[tid] = tid < 3 ? 0xffffffffu : 0u; // min of nothing, max of nothing (ordered encoding)
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

    // ---- exclusive scan of the dead counts inside the tile (particle order: group, thread, lane) ----
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
        s_base[s * THREADS + tid] = (unsigned short)(before + incl[s] - cnt[s]); // dead particles of the tile before this thread's group s
#pragma unroll
        for (int l = 0; l < V; ++l) s_en[(s * THREADS + tid) * V + l] = en[s][l];
    }

    // A tile in the first half of the emitter needs no prefix to know that every slot of it is examined:
    // D(k) <= k, so k + D(k) <= 2k < N.  (Strictly below N/2: the last examined original, which publishes the new
    // count, then never lies in such a tile -- K >= N/2 examined originals, and 2 * (first + TILE) < N puts every k of
    // the tile below N/2 - 1.)  If it has no dead particle either, nothing in it depends on the emitter-wide prefix and
    // its warps start streaming at once; warp 0 still resolves and publishes the prefix for the tiles behind it.
    const bool certain = 2ull * ((unsigned long long)first + TILE) < (unsigned long long)N;
    const bool need_prefix = !certain || tile_dead != 0;

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
        // Every round inspects the 32 * kLook nearest predecessors not yet accounted for: each lane requests kLook status
        // words at once (the windows further back travel while the nearest one is examined), so a tile that starts in the
        // middle of a wave of ~300 concurrent tiles resolves its prefix in two or three L2 round trips instead of ten.
#if defined(T3D_EXP_LOOK)
        constexpr int kLook = T3D_EXP_LOOK;
#else
        constexpr int kLook = 4;
#endif
        uint32_t excl = 0;
        int look = (int)jt - 1; // nearest predecessor, as a tile index inside the emitter
        bool resolved = false;
        while (!resolved)
        {
            unsigned long long st[kLook];
            const unsigned long long* sp[kLook];
#pragma unroll
            for (int w = 0; w < kLook; ++w)
            {
                const int idx = look - (int)lane - 32 * w;
                sp[w] = idx >= 0 ? tile_status + (it.tile_begin + (uint32_t)idx) : nullptr;
                st[w] = sp[w] ? ld_status(sp[w]) : (tag | kStatusPrefix); // before the emitter's first tile: prefix 0
            }
#pragma unroll
            for (int w = 0; w < kLook; ++w)
            {
                if (resolved) break;
                while (true) // until this window's 32 tiles have all published something in this launch
                {
                    const bool ready = (st[w] >> 34) == (unsigned long long)epoch && ((st[w] >> 32) & 3ull) != 0ull;
                    if (__all_sync(0xffffffffu, ready)) break;
                    if (!ready) st[w] = ld_status(sp[w]);
                }
                const bool is_prefix = ((st[w] >> 32) & 3ull) == 2ull;
                const unsigned pmask = __ballot_sync(0xffffffffu, is_prefix);
                const int first_p = __ffs((int)pmask) - 1;
                uint32_t contrib = (pmask == 0u || (int)lane <= first_p) ? (uint32_t)st[w] : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                excl += contrib;
                resolved = pmask != 0u;
            }
            look -= 32 * kLook;
        }
        if (lane == 0)
        {
            st_status(my_status, tag | kStatusPrefix | (excl + tile_dead));
            s_excl = excl;
        }
    }
    if (need_prefix) __syncthreads();
    const uint32_t tile_excl = need_prefix ? s_excl : 0u; // (not needed otherwise: no particle of the tile is dead)

    // What the reference's loop does with the V slots of one thread's group s (k0 = index of the first).
    struct Plan
    {
        int e[V];            // energy after this frame's decrement (PE.cpp:753)
        bool alive[V];       // examined and alive: updated in place
        bool mover[V];       // examined and dead, and not the last particle: original src[l] takes the slot, as it stands
        bool stored[V];      // alive or mover: the slot belongs to the new live range
        uint32_t src[V];
        bool all_examined;   // every slot of the group is examined (whole-group vector stores are safe)
        bool any_mover;
        bool last_examined;  // this group holds the last examined original: publishes the new count
        uint32_t new_count;
    };
    auto make_plan = [&](int s, uint32_t k0, Plan& p)
    {
        const int* e_s = s_en + (s * THREADS + tid) * V;
        uint32_t run = tile_excl + s_base[s * THREADS + tid];
        p.all_examined = true;
        p.any_mover = false;
        p.last_examined = false;
        p.new_count = 0;
#pragma unroll
        for (int l = 0; l < V; ++l)
        {
            const uint32_t k = k0 + l;
            p.e[l] = e_s[l];
            const bool valid = k < N;
            const bool live = valid && p.e[l] > 0;
            // without the prefix (certain tile, nobody dead) `run` is not the emitter-wide count, and is not needed
            const bool examined = valid && (!need_prefix || k + run < N);
            p.alive[l] = examined && live;
            p.src[l] = N - 1u - run; // PE.cpp:825-828: the last live particle, untouched this frame, takes the dead one's place
            p.mover[l] = examined && !live && p.src[l] != k; // (the dead one IS the last: nothing moves)
            p.stored[l] = p.alive[l] || p.mover[l];
            p.any_mover = p.any_mover || p.mover[l];
            p.all_examined = p.all_examined && examined;
            const uint32_t dead_l = (valid && !live) ? 1u : 0u;
            if (need_prefix && examined && !((k + 1u) + run + dead_l < N))
            {
                p.last_examined = true; // PE.cpp:829: the count the loop leaves behind
                p.new_count = N - run - dead_l;
            }
            run += dead_l;
        }
    };

    // ---- phase B: stream the groups ----
    float ppf = it.percent_per_frame;
    float dur = it.frame_duration_secs;
    uint32_t frame_count = it.frame_count;
    {
        // The group loop keeps the next group's loads in flight.  A value that is still "a pending load" when the loop
        // is entered would make its consumers inside the loop wait on the scoreboard those loads share (ptxas puts
        // nearly all of them on one), i.e. for the group that was just requested.  Passing the per-tile constants
        // through one ALU instruction here settles them before the loop (`zero` is 0, unknown to the compiler).
        auto settle_f = [&](float& v) { v = __uint_as_float(__float_as_uint(v) ^ zero); };
        auto settle_u = [&](uint32_t& v) { v ^= zero; };
        settle_f(ppf);
        settle_f(dur);
        settle_u(frame_count);
        if (FUSED)
        {
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
    }
    // Group 0 was requested before the prefix was known: the slots that turn out to receive a mover fetch it now.
    if (need_prefix)
    {
        const uint32_t k0 = first + tid * V;
        if (k0 < N)
        {
            Plan p0;
            make_plan(0, k0, p0);
            if (p0.any_mover)
            {
#pragma unroll
                for (int l = 0; l < V; ++l)
                    if (p0.mover[l])
                    {
                        load_one(p0.src[l], cur, l);
                        prefetch_mover(p0.src[l]); // the rest is fetched when the slot is written (group 0 only)
                    }
            }
        }
    }
    float box_lo[3] = { 3.402823466e38f, 3.402823466e38f, 3.402823466e38f };
    float box_hi[3] = { -3.402823466e38f, -3.402823466e38f, -3.402823466e38f };

#pragma unroll 1
    for (int s = 0; s < SUB; ++s)
    {
        const uint32_t w0 = first + s * GROUP + warp * WG; // first particle of the warp's group
        const uint32_t k0 = w0 + lane * V;
        if (w0 >= N) break; // warps leave together: the record and vertex paths are warp-collective
        Plan p;
        make_plan(s, k0, p);

        // -- the warp's life records of this group have landed: every lane takes its own.  A slot that receives a mover
        //    (PE.cpp:825-828) finds the MOVER's record there: it was requested in the dead particle's place, one group
        //    ahead (below); the words of the mover that this thread only passes on came through the side stage.  Group 0
        //    was requested before the prefix was known: its movers are fetched directly, further down. --
        const bool staged = s > 0;
        bool fastm[V];        // movers whose extras came through the side stage: the first 32 of the warp's group
        uint32_t side_slot[V]; // ... and where (the same ranks as when they were requested: slot l of every lane, then slot l + 1)
        {
            uint32_t base = 0;
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                const unsigned m = __ballot_sync(0xffffffffu, staged && p.mover[l]);
                side_slot[l] = base + (uint32_t)__popc(m & lanes_below);
                fastm[l] = staged && p.mover[l] && side_slot[l] < 32u;
                base += (uint32_t)__popc(m);
            }
        }
        uint32_t rc[V][REC_WORDS];
        float col_out[4][V], size_out[V];
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int l = 0; l < V; ++l) rec_read<V>(rec_warp, lane, l, rc[l]);
        if (staged && p.any_mover)
        {
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                if (!p.mover[l]) continue;
                uint4* rdst = reinterpret_cast<uint4*>(rec + (size_t)(k0 + l) * REC_WORDS);
#pragma unroll
                for (int c = 0; c < 4; ++c) rdst[c] = make_uint4(rc[l][c * 4], rc[l][c * 4 + 1], rc[l][c * 4 + 2], rc[l][c * 4 + 3]);
                if (!fastm[l]) continue;
                const unsigned char* const side = side_warp + side_slot[l] * kSideLane;
                const uint4 ro = *reinterpret_cast<const uint4*>(side);
                const uint4 w0s = *reinterpret_cast<const uint4*>(side + 16);
                const uint2 w1s = *reinterpret_cast<const uint2*>(side + 32);
                *reinterpret_cast<uint4*>(rot + (size_t)(k0 + l) * ROT_WORDS) = ro;
                p.e[l] = (int)w0s.x;
                col_out[0][l] = __uint_as_float(w0s.y);
                col_out[1][l] = __uint_as_float(w0s.z);
                col_out[2][l] = __uint_as_float(w0s.w);
                col_out[3][l] = __uint_as_float(w1s.x);
                size_out[l] = __uint_as_float(w1s.y);
            }
        }
        __syncwarp();

        // -- the next group: its plan, its records (movers' in place of the dead), its rw words (the slot's own, or those
        //    of the original that takes its place) --
        const bool more = s + 1 < SUB;
        const uint32_t wn = w0 + GROUP, kn = k0 + GROUP;
        Plan pn;
        bool mv_next = false;
        if (more && need_prefix && kn < N)
        {
            make_plan(s + 1, kn, pn);
            mv_next = pn.any_mover;
        }
        const bool warp_mv = need_prefix && __any_sync(0xffffffffu, mv_next);
        if (more && wn < N)
        {
            const uint32_t nrec = cap - wn < (uint32_t)WG ? cap - wn : (uint32_t)WG;
            if (!warp_mv)
                rec_issue<V>(rec_warp, rec + (size_t)wn * REC_WORDS, nrec, lane);
            else
            {
                uint32_t idx[V];
#pragma unroll
                for (int l = 0; l < V; ++l) idx[l] = (mv_next && pn.mover[l]) ? pn.src[l] : kn + l;
                rec_issue_indexed<V>(rec_warp, rec, nrec, lane, idx);
            }
        }
        if (warp_mv)
        {
            uint32_t base = 0;
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                const bool mine = mv_next && pn.mover[l];
                const unsigned m = __ballot_sync(0xffffffffu, mine);
                const uint32_t slot = base + (uint32_t)__popc(m & lanes_below);
                base += (uint32_t)__popc(m);
                if (!mine) continue;
                const uint32_t i = pn.src[l];
                if (slot < 32u)
                {
                    unsigned char* const side = side_warp + slot * kSideLane;
                    cp_async16(side, rot + (size_t)i * ROT_WORDS);
                    cp_async4(side + 16, col(RW_ENERGY) + i);
#pragma unroll
                    for (int c = 0; c < 4; ++c) cp_async4(side + 20 + 4 * c, col(RW_COLOR + c) + i);
                    cp_async4(side + 36, col(RW_SIZE) + i);
                }
                else
                    prefetch_mover(i); // more than 32 movers in the warp's group: fetched when the slot is written, ask L2 for it now
            }
        }
        cp_async_commit();
        RwIn<V> nxt;
        if (more && kn < cap)
        {
            if (!mv_next)
                load_own(kn, nxt);
            else
            {
#pragma unroll
                for (int l = 0; l < V; ++l) load_one(pn.mover[l] ? pn.src[l] : kn + l, nxt, l);
            }
        }

        // -- PE.cpp:757-819 on the particles that are alive --
        float ax[V], ay[V], az[V];
#pragma unroll
        for (int l = 0; l < V; ++l)
        {
            ax[l] = __uint_as_float(rc[l][REC_ACC + 0]);
            ay[l] = __uint_as_float(rc[l][REC_ACC + 1]);
            az[l] = __uint_as_float(rc[l][REC_ACC + 2]);
        }
        if (may_rotate)
        {
            // rare (no stock .particle file rotates): the 16-byte rotation records are fetched here, unprefetched
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                if (!p.alive[l]) continue;
                const uint4 rr = *reinterpret_cast<const uint4*>(rot + (size_t)(k0 + l) * ROT_WORDS);
                const float rx = __uint_as_float(rr.x), ry = __uint_as_float(rr.y), rz = __uint_as_float(rr.z), rs = __uint_as_float(rr.w);
                if (rs != 0.0f && !(rx == 0.0f && ry == 0.0f && rz == 0.0f))
                {
                    const Rot3 r = create_rotation(rx, ry, rz, rs * secs); // :759
                    rot_apply(r, 1.0f, cur.vx[l], cur.vy[l], cur.vz[l]);   // :761
                    rot_apply(r, 1.0f, ax[l], ay[l], az[l]);               // :762
                    // the rotated acceleration goes back into the life record (its chunk also holds the spin)
                    *reinterpret_cast<uint4*>(rec + (size_t)(k0 + l) * REC_WORDS + REC_ACC) =
                        make_uint4(__float_as_uint(ax[l]), __float_as_uint(ay[l]), __float_as_uint(az[l]), rc[l][REC_SPIN]);
                }
            }
        }
#pragma unroll
        for (int l = 0; l < V; ++l)
        {
            if (!p.alive[l]) continue;
            cur.vx[l] += ax[l] * secs; // :766-768
            cur.vy[l] += ay[l] * secs;
            cur.vz[l] += az[l] * secs;
            cur.px[l] += cur.vx[l] * secs; // :770-772
            cur.py[l] += cur.vy[l] * secs;
            cur.pz[l] += cur.vz[l] * secs;
            cur.ang[l] += __uint_as_float(rc[l][REC_SPIN]) * secs; // :774
            const float percent = 1.0f - ((float)p.e[l] / (float)(int)rc[l][REC_ENERGY_START]); // :777
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
                const float cs = __uint_as_float(rc[l][REC_COLOR_START + c]), ce = __uint_as_float(rc[l][REC_COLOR_END + c]);
                col_out[c][l] = cs + (ce - cs) * percent; // :779-782
            }
            {
                const float ss = __uint_as_float(rc[l][REC_SIZE_START]), se = __uint_as_float(rc[l][REC_SIZE_END]);
                size_out[l] = ss + (se - ss) * percent; // :784
            }
            if (animated)
            {
                if (!looped)
                {
                    float spent = 0.0f; // :792-796, repeated addition
                    for (int q = 0; q < cur.fr[l]; ++q) spent += ppf;
                    cur.tf[l] = percent - spent;
                    if ((uint32_t)cur.fr[l] < frame_count - 1u && cur.tf[l] >= ppf) ++cur.fr[l];
                }
                else
                {
                    cur.tf[l] += secs; // :808-817
                    if (cur.tf[l] >= dur)
                    {
                        cur.tf[l] -= dur;
                        ++cur.fr[l];
                        if ((uint32_t)cur.fr[l] == frame_count) cur.fr[l] = 0;
                    }
                }
            }
        }

        // -- movers that did not come through the stages: those of group 0 (requested before the prefix was known) and a
        //    warp's movers beyond the 32nd of a group.  Their motion words came with the prefetch; the rest is fetched here, from
        //    L2 where the prefetch got there in time. --
        if (p.any_mover)
        {
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                if (!p.mover[l] || fastm[l]) continue;
                const uint32_t i = p.src[l];
                if (!staged)
                {
                    const uint4* rsrc = reinterpret_cast<const uint4*>(rec + (size_t)i * REC_WORDS);
                    const uint4 r0 = rsrc[0], r1 = rsrc[1], r2 = rsrc[2], r3 = rsrc[3];
                    uint4* rdst = reinterpret_cast<uint4*>(rec + (size_t)(k0 + l) * REC_WORDS);
                    rdst[0] = r0; rdst[1] = r1; rdst[2] = r2; rdst[3] = r3;
                }
                const uint4 ro = *reinterpret_cast<const uint4*>(rot + (size_t)i * ROT_WORDS);
                p.e[l] = ld_i(col(RW_ENERGY) + i);
#pragma unroll
                for (int c = 0; c < 4; ++c) col_out[c][l] = ld_f(col(RW_COLOR + c) + i);
                size_out[l] = ld_f(col(RW_SIZE) + i);
                *reinterpret_cast<uint4*>(rot + (size_t)(k0 + l) * ROT_WORDS) = ro;
            }
        }

        if (FUSED)
        {
            // -- the quads of this group: PE.cpp:866-882 + SB.cpp:292-363 on the values the slots now hold --
            // Slots of the new live range form a prefix of the group (k + D(k) is increasing); the rest lie beyond the
            // new count.
            uint32_t n_quads = 0;
#pragma unroll
            for (int l = 0; l < V; ++l) n_quads += p.stored[l] ? 1u : 0u;
            const uint32_t warp_quads = __reduce_add_sync(0xffffffffu, n_quads);
            if (warp_quads)
            {
                // the warp's piece of shared memory is reused: its previous bulk store must have read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int l = 0; l < V; ++l)
                {
                    if (!p.stored[l]) continue;
                    const uint32_t frame = (uint32_t)cur.fr[l] < frame_count ? (uint32_t)cur.fr[l] : frame_count - 1u;
                    // (always from shared memory: a global load here would share a scoreboard with the loads in flight)
                    const float4 uv = *reinterpret_cast<const float4*>(s_uv + frame * 4u);
                    const float pos[3] = { cur.px[l], cur.py[l], cur.pz[l] };
                    const float rgba[4] = { col_out[0][l], col_out[1][l], col_out[2][l], col_out[3][l] };
                    expand_quad(pos, rgba, size_out[l], cur.ang[l], uv.x, uv.y, uv.z, uv.w, right, up, rot_mode, [&] { return frozen_rot; },
                                reinterpret_cast<float4*>(stage_warp + (lane * V + l) * kQuadBytes));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                {
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(stage_warp);
                    const uint32_t bytes = warp_quads * (uint32_t)kQuadBytes;
                    float* dst = vtx + (size_t)w0 * T3D_FLOATS_PER_QUAD;
#if defined(T3D_EXP_BULK_POLICY) // e.g. -DT3D_EXP_BULK_POLICY="evict_first": the vertex stream with an L2 eviction priority
                    unsigned long long pol;
                    asm volatile("createpolicy.fractional.L2::" T3D_EXP_BULK_POLICY ".b64 %0, 1.0;" : "=l"(pol));
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(pol) : "memory");
#else
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
#endif
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (BOUNDS)
        {
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                if (!p.stored[l]) continue;
                const float c3[3] = { cur.px[l], cur.py[l], cur.pz[l] };
#pragma unroll
                for (int k = 0; k < 3; ++k)
                {
                    box_lo[k] = fminf(box_lo[k], c3[k] - size_out[l]);
                    box_hi[k] = fmaxf(box_hi[k], c3[k] + size_out[l]);
                }
            }
        }

        // -- the group's rw words.  Every slot examined: whole-group vector stores (a dead last particle that nobody
        //    replaces gets don't-care values; it lies beyond the new count).  Otherwise the group straddles the end of the
        //    examined range: the slots behind it are the movers' sources and must not be touched. --
        if (p.all_examined)
        {
            Vec<V>::st(col(RW_ENERGY) + k0, p.e);
            Vec<V>::st(col(RW_POS + 0) + k0, cur.px);
            Vec<V>::st(col(RW_POS + 1) + k0, cur.py);
            Vec<V>::st(col(RW_POS + 2) + k0, cur.pz);
            Vec<V>::st(col(RW_VEL + 0) + k0, cur.vx);
            Vec<V>::st(col(RW_VEL + 1) + k0, cur.vy);
            Vec<V>::st(col(RW_VEL + 2) + k0, cur.vz);
            Vec<V>::st(col(RW_ANGLE) + k0, cur.ang);
#pragma unroll
            for (int c = 0; c < 4; ++c) Vec<V>::st(col(RW_COLOR + c) + k0, col_out[c]);
            Vec<V>::st(col(RW_SIZE) + k0, size_out);
            if (animated)
            {
                Vec<V>::st(col(RW_TIME_ON_FRAME) + k0, cur.tf);
                Vec<V>::st(col(RW_FRAME) + k0, cur.fr);
            }
            else if (p.any_mover)
            {
                // an emitter that does not animate never rewrites these two columns; a mover still brings its own
#pragma unroll
                for (int l = 0; l < V; ++l)
                {
                    if (!p.mover[l]) continue;
                    col(RW_TIME_ON_FRAME)[k0 + l] = __float_as_uint(cur.tf[l]);
                    col(RW_FRAME)[k0 + l] = (uint32_t)cur.fr[l];
                }
            }
        }
        else
        {
#pragma unroll
            for (int l = 0; l < V; ++l)
            {
                if (!p.stored[l]) continue;
                const uint32_t k = k0 + l;
                auto S = [&](int c, float v) { col(c)[k] = __float_as_uint(v); };
                col(RW_ENERGY)[k] = (uint32_t)p.e[l];
                S(RW_POS + 0, cur.px[l]);
                S(RW_POS + 1, cur.py[l]);
                S(RW_POS + 2, cur.pz[l]);
                S(RW_VEL + 0, cur.vx[l]);
                S(RW_VEL + 1, cur.vy[l]);
                S(RW_VEL + 2, cur.vz[l]);
                S(RW_ANGLE, cur.ang[l]);
#pragma unroll
                for (int c = 0; c < 4; ++c) S(RW_COLOR + c, col_out[c][l]);
                S(RW_SIZE, size_out[l]);
                if (animated)
                {
                    S(RW_TIME_ON_FRAME, cur.tf[l]);
                    col(RW_FRAME)[k] = (uint32_t)cur.fr[l];
                }
                else if (p.mover[l])
                {
                    S(RW_TIME_ON_FRAME, cur.tf[l]);
                    col(RW_FRAME)[k] = (uint32_t)cur.fr[l];
                }
            }
        }

        // the last examined original publishes the new count (PE.cpp:829)
        if (p.last_examined)
        {
            d->cnt[parity ^ 1].count = p.new_count;
            d->cnt[parity ^ 1].serial = d->cnt[parity].serial;
            if (FUSED) d->cnt[0].drawn = p.new_count;
        }
        cur = nxt;
        compiler_fence();
    }
    // (threads that left the loop early may still have a record fetch in flight)
    cp_async_wait<0>();
    // shared memory must outlive the bulk stores' reads
    if (FUSED && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");

    if (BOUNDS)
    {
        // warp -> CTA -> emitter: 64-bit atomicMax on (epoch << 32 | ordered bits); lo keeps the complement (max = min)
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const uint32_t lo = __reduce_min_sync(0xffffffffu, float_order(box_lo[k]));
            const uint32_t hi = __reduce_max_sync(0xffffffffu, float_order(box_hi[k]));
            if (lane == 0)
            {
                atomicMin(&s_box[k], lo);
                atomicMax(&s_box[3 + k], hi);
            }
        }
        __syncthreads();
        if (tid < 6 && (flags & kFlagBounds))
        {
            const uint32_t v = s_box[tid];
            const bool empty = tid < 3 ? v == float_order(3.402823466e38f) : v == float_order(-3.402823466e38f);
            if (!empty)
            {
                unsigned long long* dst = tid < 3 ? &d->bounds.lo[tid] : &d->bounds.hi[tid - 3];
                atomicMax(dst, ((unsigned long long)epoch << 32) | (tid < 3 ? ~v : v));
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------
// variants: {particles per thread per group, threads per CTA, groups per thread, min CTAs per SM}.  The
// defaults are chosen from measurements on B200 (profiles/); T3D_UPDATE_VARIANT / T3D_FUSED_VARIANT select
// another for experiments and for the parity tests of the other shapes.
// --------------------------------------------------------------------------------------------------
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

template <int V, int THREADS, int SUB, int MIN_BLOCKS, bool TICKET, bool FUSED, bool BOUNDS>
static void launch_one(cudaStream_t s, const UpdateLaunch& L)
{
    // life records of one group (+ the group's vertex staging) + alignment slack
    constexpr size_t smem = (size_t)THREADS * V * (REC_WORDS * 4 + (FUSED ? T3D_FLOATS_PER_QUAD * 4 : 0)) + (size_t)THREADS * 48 + 128;
    auto kernel = k_update<V, THREADS, SUB, MIN_BLOCKS, TICKET, FUSED, BOUNDS>;
    // the opt-in to more than 48 KB of dynamic shared memory is per device: once for each device this process launches on
    static bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (!configured[dev])
    {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured[dev] = true;
        // the shapes are tuned for MIN_BLOCKS resident CTAs per SM; shared memory is within 4 KB of that limit, so say so
        // loudly if a change ever pushes a variant over it (the kernel would still be correct, at half the speed)
        int resident = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, THREADS, smem) == cudaSuccess && resident < MIN_BLOCKS)
            fprintf(stderr, "t3d: k_update<%d,%d,%d,%d,fused=%d>: %d resident CTA(s) per SM instead of %d (shared memory %zu B dynamic)\n", V,
                    THREADS, SUB, MIN_BLOCKS, (int)FUSED, resident, MIN_BLOCKS, smem);
    }
    kernel<<<L.total_tiles, THREADS, smem, s>>>(L.items, L.n_items, L.total_tiles, L.tile_status, L.ticket, L.ticket_base, L.epoch, L.fault, 0u,
                                                L.ditems, L.rot_mode, L.frozen);
}
template <int V, int THREADS, int SUB, int MIN_BLOCKS, bool FUSED>
static void launch_shape(cudaStream_t s, const UpdateLaunch& L)
{
    const bool t = use_ticket();
    if (L.bounds)
    {
        if (t) launch_one<V, THREADS, SUB, MIN_BLOCKS, true, FUSED, true>(s, L);
        else launch_one<V, THREADS, SUB, MIN_BLOCKS, false, FUSED, true>(s, L);
    }
    else
    {
        if (t) launch_one<V, THREADS, SUB, MIN_BLOCKS, true, FUSED, false>(s, L);
        else launch_one<V, THREADS, SUB, MIN_BLOCKS, false, FUSED, false>(s, L);
    }
}

struct Variant
{
    const char* name;
    int tile;
    void (*launch)(cudaStream_t, const UpdateLaunch&);
};
#define T3D_UPDATE_SHAPE(V, T, S, B) { "u" #V "t" #T "s" #S "b" #B, V * T * S, launch_shape<V, T, S, B, false> }
#define T3D_FUSED_SHAPE(V, T, S, B) { "f" #V "t" #T "s" #S "b" #B, V * T * S, launch_shape<V, T, S, B, true> }
static const Variant kUpdate[] = {
    // default: 2 particles per thread x 128 threads x 12 groups = 3 072-particle tiles, 4 CTAs per SM.  Without the vertex
    // stage the kernel needs little shared memory, and four small CTAs overlap each other's tile prologues better than two
    // large ones: 0.90 of the HBM peak on C3 against 0.88 for 192 threads x 2 CTAs (profiles/r2_update_alone_variants.txt)
    T3D_UPDATE_SHAPE(2, 128, 12, 4),
    T3D_UPDATE_SHAPE(2, 192, 12, 2), T3D_UPDATE_SHAPE(2, 256, 8, 2), T3D_UPDATE_SHAPE(4, 128, 8, 2),
};
static const Variant kFused[] = {
    // default: 2 particles per thread, 192 threads x 12 groups = 4 608-particle tiles, 2 CTAs/SM
    T3D_FUSED_SHAPE(2, 192, 12, 2),
    T3D_FUSED_SHAPE(2, 192, 6, 2), T3D_FUSED_SHAPE(2, 128, 12, 3), T3D_FUSED_SHAPE(2, 128, 8, 3), T3D_FUSED_SHAPE(2, 128, 6, 3),
};
template <size_t N>
static const Variant& pick(const Variant (&table)[N], const char* env_name, int& cache)
{
    if (cache < 0)
    {
        cache = 0;
        if (const char* env = getenv(env_name))
            for (int i = 0; i < (int)N; ++i)
                if (!strcmp(env, table[i].name)) cache = i;
    }
    return table[cache];
}
static int g_update_variant = -1, g_fused_variant = -1;
static const Variant& variant() { return pick(kUpdate, "T3D_UPDATE_VARIANT", g_update_variant); }
static const Variant& fused_variant() { return pick(kFused, "T3D_FUSED_VARIANT", g_fused_variant); }
// A launch that is only a few waves of CTAs wide loses the idle part of its last wave; with half-size tiles the tail is
// half as long (at ~3 % more per-tile overhead).  shape 0 = the variant above, shape 1 = its half-size sibling; only the
// default variant has one (an explicit T3D_*_VARIANT is taken as is).
static const Variant kUpdateSmall = T3D_UPDATE_SHAPE(2, 128, 6, 4);
static const Variant kFusedSmall = T3D_FUSED_SHAPE(2, 192, 6, 2);
static const Variant& shaped(const Variant& v, const Variant* table, const Variant& small_one, int shape)
{
    return (shape == 1 && &v == table) ? small_one : v;
}

int fused_tile_size(int shape) { return shaped(fused_variant(), kFused, kFusedSmall, shape).tile; }
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
void launch_update_draw(void* stream, const UpdateLaunch& L, int shape)
{
    shaped(fused_variant(), kFused, kFusedSmall, shape).launch((cudaStream_t)stream, L);
}

int update_tile_size(int shape) { return shaped(variant(), kUpdate, kUpdateSmall, shape).tile; }
const char* update_variant_name() { return variant().name; }
void launch_update(void* stream, const UpdateLaunch& L, int shape)
{
    shaped(variant(), kUpdate, kUpdateSmall, shape).launch((cudaStream_t)stream, L);
}

// 1 when the half-size tiles are expected to finish the launch sooner: tiles_big / tiles_small = tiles of the launch
// with either shape, slots = CTAs the device runs at once.
int pick_tile_shape(uint64_t tiles_big, uint64_t tiles_small, uint32_t slots)
{
    if (!slots || tiles_big >= 16ull * slots) return 0; // many waves: the tail does not matter
    auto cost = [&](uint64_t tiles, double tile_cost) { return (double)((tiles + slots - 1) / slots) * tile_cost; };
    // a half-size tile costs a little more than half a big one (per-tile prologue, look-back)
    return cost(tiles_small, 0.53) < cost(tiles_big, 1.0) ? 1 : 0;
}

} // namespace t3d
