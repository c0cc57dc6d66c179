// t3d_kernels.cu -- the sm_100a kernels of the particle path.
//
//   k_spawn   replaces the loop of ParticleEmitter::emitOnce          (PE.cpp:308-368 + generators :611-679)
//   k_update  replaces the per-particle loop of ParticleEmitter::update (PE.cpp:750-831), including the
//             swap-with-last removal, reproduced exactly by a prefix scan (see "removal" below)
//   k_draw    replaces ParticleEmitter::draw's loop + SpriteBatch::draw   (PE.cpp:866-882, SB.cpp:292-363)
//   k_latch   finds the first rotated quad for the reference's frozen billboard rotation (SB.cpp:339)
//
// PE.cpp = Tractor3Dlib/src/graphics/ParticleEmitter.cpp, SB.cpp = src/graphics/SpriteBatch.cpp,
// Matrix.cpp = src/math/Matrix.cpp, MathUtil.h = include/math/MathUtil.h of the reference.
//
// Numerics: compiled with --fmad=false (and the default -prec-div=true -prec-sqrt=true), and every
// expression keeps the reference's evaluation order, so all results are bit-identical to the reference's
// SSE scalar code except where it calls cosf/sinf (Matrix.cpp:375-376), which differ by a few ulp.
//
// All three hot kernels are HBM-bound streaming kernels (no reuse, ~0.3 flop/B): no tensor cores, the
// design rules are coalescing, 128-bit accesses and enough bytes in flight per SM.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include "t3d_device.cuh"
#include "t3d_internal.h"

namespace t3d
{

// --------------------------------------------------------------------------------------------------
// spawn
// --------------------------------------------------------------------------------------------------
// STREAM: some emitter of the launch reads recorded draws (T3D_RNG_STREAM / T3D_RNG_LIBC); otherwise every draw is the
// counter-based hash and the kernel carries no second path (it is bound by instruction issue, not by HBM: ~30 hash
// evaluations per particle).
template <bool STREAM>
struct DrawSource
{
    const int32_t* p; // recorded stream positioned at this particle's first draw, or nullptr
    uint32_t key;     // device generator: the per-particle part of the hash
    uint32_t j;
    __device__ __forceinline__ int32_t next()
    {
        const int32_t r = (STREAM && p) ? p[j] : device_rng_from_key(key, j);
        ++j;
        return r;
    }
};

template <bool STREAM>
__global__ void __launch_bounds__(kSpawnThreads) k_spawn(const SpawnItem* __restrict__ items, uint32_t n_items)
{
    // A CTA generates kSpawnSub consecutive groups of kSpawnThreads particles, so the descriptor fetch and the count are paid
    // once per kSpawnTile particles.  A group's life records and rotation records are assembled in shared memory and leave as
    // two contiguous bulk stores (the appended range [count0, count0 + n) is contiguous in both arrays) while the next group
    // is generated into the other half; the rw words go out column by column, one coalesced 128-byte line per warp per column.
    constexpr int kBufs = kSpawnSub > 1 ? 2 : 1;
    __shared__ __align__(128) uint32_t s_rec[kBufs][kSpawnThreads * REC_WORDS];
    __shared__ __align__(128) uint32_t s_rot[kBufs][kSpawnThreads * ROT_WORDS];
    const uint32_t tile = blockIdx.x;
    // the head of the entry this tile is expected to belong to, requested at once and checked (see k_update)
    struct Head
    {
        EmitterDev* dev;
        const int32_t* draws;
        const uint32_t* offsets;
        uint32_t draw_stride, request, parity, tile_begin, n_tiles, pad;
    };
    static_assert(sizeof(Head) == 48, "three 128-bit words");
    auto load_head = [&](uint32_t i)
    {
        Head h;
        const uint4* src = reinterpret_cast<const uint4*>(items + i);
        uint4* dst = reinterpret_cast<uint4*>(&h);
#pragma unroll
        for (int q = 0; q < 3; ++q) dst[q] = __ldg(src + q);
        return h;
    };
    uint32_t item_index = guess_item(n_items, tile, gridDim.x);
    Head it = load_head(item_index);
    if (tile - it.tile_begin >= it.n_tiles)
    {
        item_index = find_item(items, n_items, tile, gridDim.x);
        it = load_head(item_index);
    }
    const float* const world = items[item_index].world;
    EmitterDev* d = it.dev;
    const uint32_t parity = it.parity;
    const uint32_t count0 = d->cnt[parity].count;
    const uint64_t serial0 = d->cnt[parity].serial;
    const uint32_t cap = d->c.capacity;
    // PE.cpp:293-296: never go over particleCountMax.
    const uint32_t room = cap > count0 ? cap - count0 : 0u; // (the cap may have been lowered below the live count: nothing spawns)
    const uint32_t n = it.request < room ? it.request : room;
    const uint32_t tile_first = (tile - it.tile_begin) * kSpawnTile;
    if (tile_first == 0 && threadIdx.x == 0)
    {
        d->cnt[parity ^ 1].count = count0 + n;
        d->cnt[parity ^ 1].serial = serial0 + n;
    }
    if (tile_first >= n) return;
    const Storage st = d->c.st;
    const SpawnParams& P = d->c.sp;

#pragma unroll 1
    for (int sub = 0; sub < kSpawnSub; ++sub)
    {
        const uint32_t sub_first = tile_first + (uint32_t)sub * kSpawnThreads;
        if (sub_first >= n) break;
        const uint32_t i = sub_first + threadIdx.x;
        const int buf = sub & (kBufs - 1);
        if (sub >= 2)
        {
            // this half of shared memory was handed to the bulk stores two groups ago: they must have read it
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
        }
        if (i < n)
        {
            DrawSource<STREAM> src;
            src.p = (STREAM && it.draws) ? it.draws + (it.offsets ? it.offsets[i] : i * it.draw_stride) : nullptr;
            src.key = device_rng_key(P.rng_seed, serial0 + i);
            src.j = 0;
            const SpawnedParticle sp = spawn_particle(P, world, src); // PE.cpp:312-364
            const float* cs = sp.cs; const float* ce = sp.ce; const float* pos = sp.pos; const float* vel = sp.vel; const float* acc = sp.acc;
            const float* axis = sp.axis;
            const int energy = sp.energy;
            const float size_start = sp.size_start, size_end = sp.size_end, spin = sp.spin, angle = sp.angle, rot_speed = sp.rot_speed;
            const uint32_t frame = sp.frame;

            // what update() rewrites every frame: 15 rw columns
            uint32_t* const rw = st.rw + (count0 + i);
            const size_t stride = st.stride;
            auto F = [&](int c, float v) { rw[(size_t)c * stride] = __float_as_uint(v); };
            rw[(size_t)RW_ENERGY * stride] = (uint32_t)energy;
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                F(RW_POS + k, pos[k]);
                F(RW_VEL + k, vel[k]);
            }
            F(RW_ANGLE, angle);
#pragma unroll
            for (int k = 0; k < 4; ++k) F(RW_COLOR + k, cs[k]); // :314
            F(RW_SIZE, size_start);
            F(RW_TIME_ON_FRAME, 0.0f); // :365
            rw[(size_t)RW_FRAME * stride] = frame;
            // what the particle keeps for life: the 64-byte record and the rotation record, through shared memory
            uint4* const r4 = reinterpret_cast<uint4*>(s_rec[buf] + threadIdx.x * REC_WORDS);
            auto U = [](float v) { return __float_as_uint(v); };
            r4[0] = make_uint4(U(cs[0]), U(cs[1]), U(cs[2]), U(cs[3]));
            r4[1] = make_uint4(U(ce[0]), U(ce[1]), U(ce[2]), U(ce[3]));
            r4[2] = make_uint4(U(acc[0]), U(acc[1]), U(acc[2]), U(spin));
            r4[3] = make_uint4((uint32_t)energy, U(size_start), U(size_end), 0u);
            *reinterpret_cast<uint4*>(s_rot[buf] + threadIdx.x * ROT_WORDS) = make_uint4(U(axis[0]), U(axis[1]), U(axis[2]), U(rot_speed));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const uint32_t nt = (n - sub_first) < (uint32_t)kSpawnThreads ? (n - sub_first) : (uint32_t)kSpawnThreads;
            const size_t p0 = (size_t)count0 + sub_first;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(st.rec + p0 * REC_WORDS),
                         "r"((uint32_t)__cvta_generic_to_shared(s_rec[buf])), "r"(nt * (uint32_t)(REC_WORDS * 4)) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(st.rot + p0 * ROT_WORDS),
                         "r"((uint32_t)__cvta_generic_to_shared(s_rot[buf])), "r"(nt * (uint32_t)(ROT_WORDS * 4)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the stores' reads
}

// --------------------------------------------------------------------------------------------------
// draw
// --------------------------------------------------------------------------------------------------
// OUT = 0: the tile's vertices are copied from shared memory with coalesced 128-bit streaming stores.
// OUT = 1: one thread hands the tile to the TMA engine (cp.async.bulk shared -> global, SASS UBLKCP), which
//          writes the nq * 144 contiguous bytes while the other warps are already gone.
// CLIP:    vertices carry the projected position (sprite.vert:19 applied; 40 floats per quad) instead of SpriteVertex.
template <int OUT, int DSUB, bool CLIP>
__global__ void __launch_bounds__(kDrawThreads)
    k_draw(const DrawItem* __restrict__ items, uint32_t n_items, int rot_mode, const FrozenRotation* __restrict__ frozen,
           uint32_t* __restrict__ fault)
{
    constexpr int QF = CLIP ? T3D_FLOATS_PER_CLIP_QUAD : T3D_FLOATS_PER_QUAD;
    __shared__ __align__(128) float s_vtx[kDrawThreads * QF];
    __shared__ float s_uv[kMaxSmemFrames * 4];
    __shared__ float s_vp[16];

    // One CTA draws DSUB consecutive sub-tiles of kDrawThreads quads: the descriptor fetch, the uv table and the
    // live count are paid once per DSUB * 256 quads, and the next sub-tile's particles are requested before the
    // current one is expanded.
    const uint32_t tile = blockIdx.x;
    const uint32_t tid = threadIdx.x;
    // the head of the entry this tile is expected to belong to (everything but the matrix), requested at once and checked
    // (see k_update): one dependent fetch instead of three before the particle loads can start
    struct Head
    {
        EmitterDev* dev;
        const uint32_t* rw;
        float* vtx;
        const float* uv;
        uint32_t parity, tile_begin, stride, seg_capacity, frame_count, n_tiles;
        float right[3], up[3];
    };
    static_assert(sizeof(Head) == 80, "five 128-bit words");
    auto load_head = [&](uint32_t i)
    {
        Head h;
        const uint4* src = reinterpret_cast<const uint4*>(items + i);
        uint4* dsth = reinterpret_cast<uint4*>(&h);
#pragma unroll
        for (int q = 0; q < 5; ++q) dsth[q] = __ldg(src + q);
        return h;
    };
    uint32_t item_index = guess_item(n_items, tile, gridDim.x);
    Head it = load_head(item_index);
    if (tile - it.tile_begin >= it.n_tiles)
    {
        item_index = find_item(items, n_items, tile, gridDim.x);
        it = load_head(item_index);
    }
    const EmitterDev* d = it.dev;
    const uint32_t first_tile = (tile - it.tile_begin) * (kDrawThreads * DSUB);
    const uint32_t N = d->cnt[it.parity].count;
    const uint32_t seg_capacity = it.seg_capacity;
    const uint32_t* const rw = it.rw;
    const size_t stride = it.stride;

    struct Quad
    {
        float pos[3], color[4], size, angle;
        uint32_t frame;
    };
    // Requested before the live count has arrived: any slot below the padded segment size is inside the slab.
    auto load = [&](uint32_t i, Quad& q)
    {
        if (i < seg_capacity)
        {
            const uint32_t* cols = rw + i;
            auto LF = [&](int c) { return __ldcs(reinterpret_cast<const float*>(cols + (size_t)c * stride)); };
            q.pos[0] = LF(RW_POS); q.pos[1] = LF(RW_POS + 1); q.pos[2] = LF(RW_POS + 2);
            q.size = LF(RW_SIZE);
            q.color[0] = LF(RW_COLOR); q.color[1] = LF(RW_COLOR + 1); q.color[2] = LF(RW_COLOR + 2); q.color[3] = LF(RW_COLOR + 3);
            q.angle = LF(RW_ANGLE);
            q.frame = __ldcs(cols + (size_t)RW_FRAME * stride);
        }
    };
    Quad cur = {};
    load(first_tile + tid, cur);

    const uint32_t frame_count = it.frame_count;
    const float* __restrict__ uv_g = it.uv;
    const bool uv_in_smem = frame_count <= (uint32_t)kMaxSmemFrames;
    if (uv_in_smem)
        for (uint32_t q = tid; q < frame_count * 4u; q += kDrawThreads) s_uv[q] = uv_g[q];
    if (CLIP && tid < 16) s_vp[tid] = items[item_index].vp[tid];
    if (first_tile == 0 && tid == 0)
    {
        const_cast<EmitterDev*>(d)->cnt[0].drawn = N;
        // (see k_update: the grid was sized from the host's upper bound of the live count)
        if ((unsigned long long)N > (unsigned long long)it.n_tiles * (kDrawThreads * DSUB)) atomicOr(fault, kFaultDrawGrid);
    }
    if (first_tile >= N) return;
    __syncthreads();

#pragma unroll 1
    for (int sub = 0; sub < DSUB; ++sub)
    {
    const uint32_t first = first_tile + sub * kDrawThreads;
    if (first >= N) break;
    const uint32_t i = first + tid;
    Quad nxt = {};
    if (sub + 1 < DSUB) load(i + kDrawThreads, nxt);
    if (sub > 0)
    {
        // s_vtx is reused: the previous sub-tile's bulk store must have finished reading it
        if (OUT == 1 && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
    }
    const float* pos = cur.pos;
    const float* color = cur.color;
    const float size = cur.size, angle = cur.angle;
    // A frame index past the table (frameRandomOffset > frameCount set by hand: the reference then reads past its
    // array) is clamped instead of faulting.
    const uint32_t frame = cur.frame < frame_count ? cur.frame : frame_count - 1u;
    if (i < N)
    {
        const float* uv = uv_in_smem ? (s_uv + frame * 4u) : (uv_g + (size_t)frame * 4u);
        const float u1 = uv[0], v1 = uv[1], u2 = uv[2], v2 = uv[3];

        const float right[3] = { it.right[0], it.right[1], it.right[2] };
        const float up[3] = { it.up[0], it.up[1], it.up[2] };
        if (CLIP)
            expand_quad_clip(pos, color, size, angle, u1, v1, u2, v2, right, up, rot_mode, [&] { return load_frozen(frozen); }, s_vp,
                             reinterpret_cast<float4*>(s_vtx + tid * QF));
        else
            expand_quad(pos, color, size, angle, u1, v1, u2, v2, right, up, rot_mode, [&] { return load_frozen(frozen); },
                        reinterpret_cast<float4*>(s_vtx + tid * QF));
    }
    // The sub-tile's vertices: nq quads, contiguous in the emitter's vertex stream (16-byte aligned).
    const uint32_t nq = (N - first) < (uint32_t)kDrawThreads ? (N - first) : (uint32_t)kDrawThreads;
    float4* __restrict__ dst = reinterpret_cast<float4*>(it.vtx + (size_t)first * QF);
    if (OUT == 1)
    {
        // Make this thread's shared-memory writes visible to the async proxy, then let one thread issue the bulk store.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0)
        {
            const uint32_t src = (uint32_t)__cvta_generic_to_shared(s_vtx);
            const uint32_t bytes = nq * (uint32_t)(QF * sizeof(float));
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    else
    {
        __syncthreads();
        const float4* src4 = reinterpret_cast<const float4*>(s_vtx);
        for (uint32_t q = tid; q < nq * (uint32_t)(QF / 4); q += kDrawThreads) __stcs(dst + q, src4[q]);
    }
    cur = nxt;
    }
    // shared memory must outlive the last bulk store's read
    if (OUT == 1 && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// --------------------------------------------------------------------------------------------------
// frozen-rotation latch: first particle, in draw order, whose angle is non-zero (SB.cpp:327, 337-339).
// One CTA.  Writes the axis u = right x up and the angle; the host turns them into the matrix with the
// same libm sincosf the reference uses and uploads it (once per process).
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_latch(const DrawItem* __restrict__ items, uint32_t n_items, FrozenRotation* frozen)
{
    __shared__ uint32_t s_min;
    if (frozen->latched) return;
    for (uint32_t e = 0; e < n_items; ++e)
    {
        const DrawItem& it = items[e];
        const EmitterDev* d = it.dev;
        if (!(d->c.flags & kFlagMaySpin)) continue;
        const uint32_t N = d->cnt[it.parity].count;
        const uint32_t* ang0 = it.rw + (size_t)RW_ANGLE * it.stride;
        auto ANG = [&](uint32_t i) { return __uint_as_float(ang0[i]); };
        for (uint32_t base = 0; base < N; base += 256)
        {
            if (threadIdx.x == 0) s_min = 0xffffffffu;
            __syncthreads();
            const uint32_t i = base + threadIdx.x;
            if (i < N && ANG(i) != 0.0f) atomicMin(&s_min, i);
            __syncthreads();
            const uint32_t m = s_min;
            __syncthreads();
            if (m != 0xffffffffu)
            {
                if (threadIdx.x == 0)
                {
                    const float* r = it.right;
                    const float* u = it.up;
                    frozen->m[0] = (r[1] * u[2]) - (r[2] * u[1]);
                    frozen->m[1] = (r[2] * u[0]) - (r[0] * u[2]);
                    frozen->m[2] = (r[0] * u[1]) - (r[1] * u[0]);
                    frozen->m[3] = ANG(m);
                    frozen->latched = 2; // 2 = axis/angle found, matrix pending on the host
                }
                return;
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------
// launch wrappers
// --------------------------------------------------------------------------------------------------
void launch_spawn(void* stream, const SpawnItem* items, uint32_t n_items, uint32_t total_tiles, bool host_draws)
{
    if (host_draws) k_spawn<true><<<total_tiles, kSpawnThreads, 0, (cudaStream_t)stream>>>(items, n_items);
    else k_spawn<false><<<total_tiles, kSpawnThreads, 0, (cudaStream_t)stream>>>(items, n_items);
}

template <int OUT, bool CLIP>
static void launch_draw_sub(cudaStream_t st, int sub, uint32_t total_tiles, const DrawItem* items, uint32_t n_items, int rot_mode,
                            const FrozenRotation* frozen, uint32_t* fault)
{
    if (sub == 1) k_draw<OUT, 1, CLIP><<<total_tiles, kDrawThreads, 0, st>>>(items, n_items, rot_mode, frozen, fault);
    else if (sub == 2) k_draw<OUT, 2, CLIP><<<total_tiles, kDrawThreads, 0, st>>>(items, n_items, rot_mode, frozen, fault);
    else if (sub == 4) k_draw<OUT, 4, CLIP><<<total_tiles, kDrawThreads, 0, st>>>(items, n_items, rot_mode, frozen, fault);
    else k_draw<OUT, 8, CLIP><<<total_tiles, kDrawThreads, 0, st>>>(items, n_items, rot_mode, frozen, fault);
}

void launch_draw(void* stream, const DrawItem* items, uint32_t n_items, uint32_t total_tiles, int rot_mode,
                 const FrozenRotation* frozen, uint32_t* fault, int vertex_mode)
{
    const int sub = draw_sub(n_items);
    static int mode = -1;
    if (mode < 0)
    {
        const char* env = getenv("T3D_DRAW_VARIANT");
        mode = (env && !strcmp(env, "lsu")) ? 0 : 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (vertex_mode == T3D_VERTEX_CLIP)
        launch_draw_sub<1, true>(st, sub, total_tiles, items, n_items, rot_mode, frozen, fault);
    else if (mode == 1)
        launch_draw_sub<1, false>(st, sub, total_tiles, items, n_items, rot_mode, frozen, fault);
    else
        launch_draw_sub<0, false>(st, sub, total_tiles, items, n_items, rot_mode, frozen, fault);
}

// Sub-tiles of 256 quads per draw CTA.  Measured on B200 (profiles/r1_draw_subtile_sweep.txt): 4 is best when the
// launch covers many emitters (the per-CTA descriptor/uv/count prologue is amortised: 8 192 x 32 Ki draws 9 % faster),
// 2 when it covers a few large ones.  T3D_DRAW_SUB = 1, 2, 4 or 8 overrides for experiments.
int draw_sub(uint32_t n_items)
{
    static int forced = -1;
    if (forced < 0)
    {
        forced = 0;
        if (const char* env = getenv("T3D_DRAW_SUB"))
        {
            const int v = atoi(env);
            if (v == 1 || v == 2 || v == 4 || v == 8) forced = v;
        }
    }
    if (forced) return forced;
    return n_items > 8 ? 4 : 2;
}
int draw_tile_size(uint32_t n_items) { return kDrawThreads * draw_sub(n_items); }

void launch_latch(void* stream, const DrawItem* items, uint32_t n_items, FrozenRotation* frozen)
{
    k_latch<<<1, 256, 0, (cudaStream_t)stream>>>(items, n_items, frozen);
}

__global__ void k_gather_counts(const CountQuery* __restrict__ q, uint32_t n, uint32_t* __restrict__ out,
                                const uint32_t* __restrict__ fault, unsigned long long* __restrict__ boxes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) out[2 * n] = *fault;
    if (i >= n) return;
    out[2 * i] = q[i].dev->cnt[q[i].parity].count;
    out[2 * i + 1] = q[i].dev->cnt[0].drawn;
    if (boxes)
    {
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            boxes[6 * i + k] = q[i].dev->bounds.lo[k];
            boxes[6 * i + 3 + k] = q[i].dev->bounds.hi[k];
        }
    }
}

void launch_gather_counts(void* stream, const CountQuery* q, uint32_t n, uint32_t* out, const uint32_t* fault, unsigned long long* boxes)
{
    k_gather_counts<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(q, n, out, fault, boxes);
}

// MB.cpp:117-141: the first quad's four indices, then per quad two degenerate indices (previous last, next first)
// stitching the strips, then its four.
__global__ void k_strip_indices(uint32_t* __restrict__ out, uint64_t n_indices)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_indices) return;
    uint32_t v;
    if (p < 4)
        v = (uint32_t)p;
    else
    {
        const uint64_t q = (p - 4) / 6 + 1; // quad
        const uint32_t r = (uint32_t)((p - 4) % 6);
        v = r == 0 ? (uint32_t)(4 * (q - 1) + 3) : r == 1 ? (uint32_t)(4 * q) : (uint32_t)(4 * q + (r - 2));
    }
    out[p] = v;
}

__global__ void k_triangle_indices(uint32_t* __restrict__ out, uint64_t n_indices)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_indices) return;
    const uint32_t corner[6] = { 0, 1, 2, 2, 1, 3 }; // strip (a,b,c,d) = triangles (a,b,c), (c,b,d)
    out[p] = (uint32_t)(4 * (p / 6)) + corner[p % 6];
}

void launch_triangle_indices(void* stream, uint32_t* out, uint64_t n_indices)
{
    if (n_indices) k_triangle_indices<<<(unsigned)((n_indices + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n_indices);
}

void launch_strip_indices(void* stream, uint32_t* out, uint64_t n_indices)
{
    if (n_indices) k_strip_indices<<<(unsigned)((n_indices + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n_indices);
}

// Exchange with the ABI's layout (download_state / upload_state): one thread per particle, the 34 T3D_COL_* columns
// on one side, the rw columns and the two records on the other.
__device__ __forceinline__ uint32_t* storage_word(const Storage& st, uint32_t i, int c)
{
    // where word T3D_COL_* c of particle i lives
    auto RW = [&](int col) { return st.rw + (size_t)col * st.stride + i; };
    auto REC = [&](int w) { return st.rec + (size_t)i * REC_WORDS + w; };
    auto ROT = [&](int w) { return st.rot + (size_t)i * ROT_WORDS + w; };
    if (c < T3D_COL_COLOR_END) return REC(REC_COLOR_START + (c - T3D_COL_COLOR_START));
    if (c < T3D_COL_COLOR) return REC(REC_COLOR_END + (c - T3D_COL_COLOR_END));
    if (c < T3D_COL_POSITION) return RW(RW_COLOR + (c - T3D_COL_COLOR));
    if (c < T3D_COL_VELOCITY) return RW(RW_POS + (c - T3D_COL_POSITION));
    if (c < T3D_COL_ACCELERATION) return RW(RW_VEL + (c - T3D_COL_VELOCITY));
    if (c < T3D_COL_ROTATION_AXIS) return REC(REC_ACC + (c - T3D_COL_ACCELERATION));
    if (c < T3D_COL_ENERGY_START) return ROT(ROT_AXIS + (c - T3D_COL_ROTATION_AXIS));
    switch (c)
    {
    case T3D_COL_ENERGY_START: return REC(REC_ENERGY_START);
    case T3D_COL_ENERGY: return RW(RW_ENERGY);
    case T3D_COL_SIZE_START: return REC(REC_SIZE_START);
    case T3D_COL_SIZE_END: return REC(REC_SIZE_END);
    case T3D_COL_SIZE: return RW(RW_SIZE);
    case T3D_COL_ROT_PER_PARTICLE_SPEED: return REC(REC_SPIN);
    case T3D_COL_ROTATION_SPEED: return ROT(ROT_SPEED);
    case T3D_COL_ANGLE: return RW(RW_ANGLE);
    case T3D_COL_TIME_ON_FRAME: return RW(RW_TIME_ON_FRAME);
    default: return RW(RW_FRAME);
    }
}
template <bool EXPORT>
__global__ void k_exchange_state(Storage st, uint32_t n, uint32_t* host_layout, uint32_t host_stride)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < T3D_NUM_COLUMNS; ++c)
    {
        uint32_t* w = storage_word(st, i, c);
        if (EXPORT)
            host_layout[(size_t)c * host_stride + i] = *w;
        else
            *w = host_layout[(size_t)c * host_stride + i];
    }
    if (!EXPORT) st.rec[(size_t)i * REC_WORDS + REC_PAD] = 0u;
}

void launch_export_state(void* stream, const Storage& st, uint32_t n, uint32_t* out, uint32_t host_stride)
{
    if (n) k_exchange_state<true><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(st, n, out, host_stride);
}

void launch_import_state(void* stream, const Storage& st, uint32_t n, const uint32_t* in, uint32_t host_stride)
{
    if (n) k_exchange_state<false><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(st, n, const_cast<uint32_t*>(in), host_stride);
}

} // namespace t3d
