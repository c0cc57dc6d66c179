// t3d_sprites.cu -- SURVEY 8(f3): the other producers of SpriteBatch quads, expanded on the device.
//
//   k_sprite_expand   one SpriteBatch::draw 2-D call per t3d_sprite2d (reference SpriteBatch.cpp:235-289 rotated / centred,
//                     :366-413 clipped, :475-506 plain; clipSprite :523-597; Vector2::rotate, src/math/Vector2.cpp:181-200)
//   k_sprite_count /  ordered compaction for lists with clipped sprites (a sprite entirely outside its clip rectangle
//   k_sprite_scan     leaves no quad, SB.cpp:409-412): kept sprites per tile, exclusive scan over the tiles
//   k_tileset_draw    TileSet::draw (src/graphics/TileSet.cpp:207-236) from a device-resident tile table
//
// Same output as the particle path: 4 x SpriteVertex {x, y, z, u, v, r, g, b, a} per quad, in call order.  All three are
// streaming kernels: 84 B read + 144 B written per sprite (8 + 4 B read per tile); numerics as in t3d_kernels.cu
// (--fmad=false, reference evaluation order) -- bit-identical to the reference except sprites rotated by a non-zero angle,
// whose sine and cosine come from the device's sincosf instead of glibc's (stated tolerance 1e-5 relative).
#include <cuda_runtime.h>

#include "t3d_device.cuh"
#include "t3d_internal.h"

namespace t3d
{

constexpr int kSpriteThreads = 256;
constexpr int kQuadFloats = T3D_FLOATS_PER_QUAD;

// SB.cpp:523-597
__device__ __forceinline__ bool clip_sprite(const float* clip, float& x, float& y, float& w, float& h, float& u1, float& v1, float& u2,
                                            float& v2)
{
    if (x + w < clip[0] || x > clip[0] + clip[2] || y + h < clip[1] || y > clip[1] + clip[3]) return false;
    float uvw = u2 - u1, uvh = v2 - v1;
    if (x < clip[0])
    {
        const float percent = (clip[0] - x) / w, dx = clip[0] - x, du = uvw * percent;
        x = clip[0];
        w -= dx;
        u1 += du;
        uvw -= du;
    }
    if (y < clip[1])
    {
        const float percent = (clip[1] - y) / h, dy = clip[1] - y, dv = uvh * percent;
        y = clip[1];
        h -= dy;
        v1 += dv;
        uvh -= dv;
    }
    const float cx2 = clip[0] + clip[2];
    const float x2 = x + w;
    if (x2 > cx2)
    {
        const float percent = (x2 - cx2) / w;
        w = cx2 - x;
        u2 -= uvw * percent;
    }
    const float cy2 = clip[1] + clip[3];
    const float y2 = y + h;
    if (y2 > cy2)
    {
        const float percent = (y2 - cy2) / h;
        h = cy2 - y;
        v2 -= uvh * percent;
    }
    return true;
}

// Vector2::rotate (Vector2.cpp:181-200): float sine / cosine widened to double, products in double, one rounding to float.
__device__ __forceinline__ void rotate2(float& x, float& y, float px, float py, double s, double c)
{
    if (px == 0.0f && py == 0.0f)
    {
        const float tx = (float)((double)x * c - (double)y * s);
        y = (float)((double)y * c + (double)x * s);
        x = tx;
    }
    else
    {
        const float tx = x - px, ty = y - py;
        x = (float)((double)tx * c - (double)ty * s + (double)px);
        y = (float)((double)ty * c + (double)tx * s + (double)py);
    }
}

__device__ __forceinline__ void write_quad(float4* o, const float* px, const float* py, float z, float u1, float v1, float u2, float v2,
                                           const float* c)
{
    // vertex k = (px[k], py[k], z, u, v, colour); u = u1 for k < 2, v = v1 for even k (SB.cpp:279-283, :497-501)
    o[0] = make_float4(px[0], py[0], z, u1);
    o[1] = make_float4(v1, c[0], c[1], c[2]);
    o[2] = make_float4(c[3], px[1], py[1], z);
    o[3] = make_float4(u1, v2, c[0], c[1]);
    o[4] = make_float4(c[2], c[3], px[2], py[2]);
    o[5] = make_float4(z, u2, v1, c[0]);
    o[6] = make_float4(c[1], c[2], c[3], px[3]);
    o[7] = make_float4(py[3], z, u2, v2);
    o[8] = make_float4(c[0], c[1], c[2], c[3]);
}

// The cull test alone (the first statement of clipSprite): does this sprite leave a quad?
__device__ __forceinline__ bool sprite_kept(const t3d_sprite2d& sp)
{
    if (!(sp.flags & T3D_SPRITE_CLIPPED)) return true;
    return !(sp.x + sp.width < sp.clip[0] || sp.x > sp.clip[0] + sp.clip[2] || sp.y + sp.height < sp.clip[1] ||
             sp.y > sp.clip[1] + sp.clip[3]);
}

__global__ void __launch_bounds__(kSpriteThreads) k_sprite_count(const t3d_sprite2d* __restrict__ sprites, uint32_t n, uint32_t* __restrict__ tile_counts)
{
    const uint32_t i = blockIdx.x * kSpriteThreads + threadIdx.x;
    const bool kept = i < n && sprite_kept(sprites[i]);
    const uint32_t c = __syncthreads_count(kept);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = c;
}

// counts[0 .. n_tiles) -> exclusive prefix in place; counts[n_tiles] = total.  One CTA.
__global__ void __launch_bounds__(1024) k_sprite_scan(uint32_t* __restrict__ counts, uint32_t n_tiles)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? counts[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t before = s_carry;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) before += s_warp[w];
        if (i < n_tiles) counts[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[n_tiles] = s_carry;
}

// COMPACT: quads of the tile go to tile_base[tile] + rank among the tile's kept sprites; otherwise sprite i -> quad i.
template <bool COMPACT>
__global__ void __launch_bounds__(kSpriteThreads)
    k_sprite_expand(const t3d_sprite2d* __restrict__ sprites, uint32_t n, const uint32_t* __restrict__ tile_base, float* __restrict__ dst)
{
    // The tile's 256 records (21 504 contiguous bytes) come in with coalesced 128-bit loads; the same shared memory then
    // stages the tile's quads, which leave as one bulk store.
    __shared__ __align__(128) float4 s_buf[kSpriteThreads * 9];
    __shared__ uint32_t s_warp[kSpriteThreads / 32];
    constexpr uint32_t kWords = sizeof(t3d_sprite2d) / 4; // 21
    static_assert(sizeof(t3d_sprite2d) == 84 && (kSpriteThreads * sizeof(t3d_sprite2d)) % 16 == 0, "tiles of records are whole 16-byte chunks");
    const uint32_t tid = threadIdx.x;
    const uint32_t first = blockIdx.x * kSpriteThreads;
    const uint32_t count = n - first < (uint32_t)kSpriteThreads ? n - first : (uint32_t)kSpriteThreads;
    {
        const uint4* src = reinterpret_cast<const uint4*>(sprites + first); // 84 * 256 * tile: 16-byte aligned when `sprites` is
        uint4* dstw = reinterpret_cast<uint4*>(s_buf);
        const uint32_t chunks = (count * kWords * 4u + 15u) / 16u;
        const uint32_t whole = (count * kWords * 4u) / 16u;
        for (uint32_t q = tid; q < whole; q += kSpriteThreads) dstw[q] = __ldcs(src + q);
        if (chunks != whole && tid == 0)
        {
            // a ragged last tile ends in half a chunk: word by word, never past the end of the list
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + whole);
            uint32_t* d32 = reinterpret_cast<uint32_t*>(dstw + whole);
            for (uint32_t w = 0; w < (count * kWords) % 4u; ++w) d32[w] = s32[w];
        }
    }
    __syncthreads();
    t3d_sprite2d sp;
    {
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(s_buf) + tid * kWords;
        uint32_t* w = reinterpret_cast<uint32_t*>(&sp);
#pragma unroll
        for (uint32_t k = 0; k < kWords; ++k) w[k] = tid < count ? mine[k] : 0u;
    }
    __syncthreads(); // every record is in registers: the buffer now stages vertices

    float x = sp.x, y = sp.y, w = sp.width, h = sp.height, u1 = sp.u1, v1 = sp.v1, u2 = sp.u2, v2 = sp.v2;
    bool kept = tid < count;
    if (kept && (sp.flags & T3D_SPRITE_CLIPPED))
        kept = clip_sprite(sp.clip, x, y, w, h, u1, v1, u2, v2); // SB.cpp:409-412
    else if (sp.flags & T3D_SPRITE_CENTERED)
    {
        x -= 0.5f * w; // SB.cpp:248-252, :487-491
        y -= 0.5f * h;
    }
    uint32_t rank = tid;
    uint32_t tile_kept = count;
    if (COMPACT)
    {
        const unsigned ballot = __ballot_sync(0xffffffffu, kept);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(ballot);
        __syncthreads();
        rank = __popc(ballot & ((1u << (tid & 31)) - 1u));
        tile_kept = 0;
#pragma unroll
        for (int k = 0; k < kSpriteThreads / 32; ++k)
        {
            if ((uint32_t)k < (tid >> 5)) rank += s_warp[k];
            tile_kept += s_warp[k];
        }
    }
    if (kept)
    {
        const float x2 = x + w, y2 = y + h;
        float px[4], py[4];
        if ((sp.flags & T3D_SPRITE_ROTATED) && !(sp.flags & T3D_SPRITE_CLIPPED))
        {
            // SB.cpp:254-283: downLeft, upLeft, downRight, upRight
            px[0] = x; py[0] = y2; px[1] = x; py[1] = y; px[2] = x2; py[2] = y2; px[3] = x2; py[3] = y;
            if (sp.rotation_angle != 0)
            {
                float pvx = sp.rotation_point[0], pvy = sp.rotation_point[1];
                pvx *= w; pvy *= h; pvx += x; pvy += y;
                float s, c;
                sincosf(sp.rotation_angle, &s, &c);
                // the reference turns upLeft, upRight, downLeft, downRight in that order; the four are independent
#pragma unroll
                for (int k = 0; k < 4; ++k) rotate2(px[k], py[k], pvx, pvy, (double)s, (double)c);
            }
        }
        else
        {
            // SB.cpp:493-501: (x, y), (x, y2), (x2, y), (x2, y2)
            px[0] = x; py[0] = y; px[1] = x; py[1] = y2; px[2] = x2; py[2] = y; px[3] = x2; py[3] = y2;
        }
        write_quad(s_buf + rank * 9, px, py, sp.z, u1, v1, u2, v2, sp.color);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0 && tile_kept)
    {
        const size_t q0 = COMPACT ? (size_t)tile_base[blockIdx.x] : (size_t)first;
        float* out = dst + q0 * kQuadFloats;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out),
                     "r"((uint32_t)__cvta_generic_to_shared(s_buf)), "r"(tile_kept * (uint32_t)(kQuadFloats * 4)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the store's reads
    }
}

// TileSet::draw (TileSet.cpp:207-236).  cells[q] = index of the q-th non-blank cell (row-major); xs / ys = the running
// sums the reference's loop keeps for the columns and rows of this frame (computed on the host: they are sequential float
// additions).  Each quad is the rotationPoint overload with angle 0 (TileSet.cpp:223-228 -> SB.cpp:235-289).
__global__ void __launch_bounds__(kSpriteThreads)
    k_tileset_draw(const uint32_t* __restrict__ cells, uint32_t n_quads, const float2* __restrict__ tiles, uint32_t cols,
                   const float* __restrict__ xs, const float* __restrict__ ys, float z, float tile_w, float tile_h, float wr, float hr,
                   float4 color, float* __restrict__ dst)
{
    __shared__ __align__(128) float4 s_buf[kSpriteThreads * 9];
    const uint32_t tid = threadIdx.x;
    const uint32_t first = blockIdx.x * kSpriteThreads;
    const uint32_t q = first + tid;
    if (q < n_quads)
    {
        const uint32_t cell = cells[q];
        const float2 src = tiles[cell];
        const float x = xs[cell % cols], y = ys[cell / cols];
        const float u1 = wr * src.x;          // SB.cpp:208-211
        const float v1 = 1.0f - hr * src.y;
        const float u2 = u1 + wr * tile_w;
        const float v2 = v1 - hr * tile_h;
        const float x2 = x + tile_w, y2 = y + tile_h;
        const float px[4] = { x, x, x2, x2 }, py[4] = { y2, y, y2, y };
        const float c[4] = { color.x, color.y, color.z, color.w };
        write_quad(s_buf + tid * 9, px, py, z, u1, v1, u2, v2, c);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0)
    {
        const uint32_t count = n_quads - first < (uint32_t)kSpriteThreads ? n_quads - first : (uint32_t)kSpriteThreads;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)first * kQuadFloats),
                     "r"((uint32_t)__cvta_generic_to_shared(s_buf)), "r"(count * (uint32_t)(kQuadFloats * 4)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// --------------------------------------------------------------------------------------------------
// launch wrappers
// --------------------------------------------------------------------------------------------------
uint32_t sprite_tiles(uint32_t n) { return (n + kSpriteThreads - 1) / kSpriteThreads; }

void launch_sprites(void* stream, const t3d_sprite2d* sprites, uint32_t n, uint32_t* tile_scratch, float* dst)
{
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t tiles = sprite_tiles(n);
    if (!tiles) return;
    if (tile_scratch)
    {
        k_sprite_count<<<tiles, kSpriteThreads, 0, st>>>(sprites, n, tile_scratch);
        k_sprite_scan<<<1, 1024, 0, st>>>(tile_scratch, tiles);
        k_sprite_expand<true><<<tiles, kSpriteThreads, 0, st>>>(sprites, n, tile_scratch, dst);
    }
    else
        k_sprite_expand<false><<<tiles, kSpriteThreads, 0, st>>>(sprites, n, nullptr, dst);
}

void launch_tileset(void* stream, const uint32_t* cells, uint32_t n_quads, const float* tiles_xy, uint32_t cols, const float* xs,
                    const float* ys, float z, float tile_w, float tile_h, float wr, float hr, const float color[4], float* dst)
{
    if (!n_quads) return;
    k_tileset_draw<<<sprite_tiles(n_quads), kSpriteThreads, 0, (cudaStream_t)stream>>>(
        cells, n_quads, reinterpret_cast<const float2*>(tiles_xy), cols, xs, ys, z, tile_w, tile_h, wr, hr,
        make_float4(color[0], color[1], color[2], color[3]), dst);
}

} // namespace t3d
