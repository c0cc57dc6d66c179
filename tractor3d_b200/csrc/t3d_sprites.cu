// t3d_sprites.cu -- SURVEY 8(f3): the other producers of SpriteBatch quads, expanded on the device.
//
//   k_sprite_expand   one SpriteBatch::draw 2-D call per t3d_sprite2d (reference SpriteBatch.cpp:235-289 rotated / centred,
//                     :366-413 clipped, :475-506 plain; clipSprite :523-597; Vector2::rotate, src/math/Vector2.cpp:181-200)
//   k_sprite_count /  ordered compaction for lists with clipped sprites (a sprite entirely outside its clip rectangle
//   k_sprite_scan     leaves no quad, SB.cpp:409-412): kept sprites per tile, exclusive scan over the tiles
//   k_tileset_draw    TileSet::draw (src/graphics/TileSet.cpp:207-236) from a device-resident tile table
//   k_text_summary /  Font::drawText (src/renderer/Font.cpp:242-390): the pen walk as a segmented scan, one quad per glyph
//   k_text_scan /     (see the section further down)
//   k_text_expand
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
// Font::drawText(text, x, y, color, size, rightToLeft) (src/renderer/Font.cpp:242-390): the reference walks the string with a
// running pen -- ' ' and '\t' advance it by the first glyph's advance (x 4), '\r' / '\n' start a new line, a character with a
// glyph draws one quad (SpriteBatch::draw(x, y, w, h, u1, v1, u2, v2, color), SB.cpp:366-378 -> :475-506) and advances by
// floor(advance * scale + spacing); everything else (also every byte >= 128: `char` is signed, Font.cpp:359) is skipped.
// The pen is an int and the advances are integers (the reference adds them in double, exactly), so the walk is a
// segmented scan: element = (advance since the last line break, line breaks, glyphs), combined left to right.
//   k_text_summary   every thread folds 16 characters (one 128-bit load), 8 lanes a tile of 128, a CTA its block of 32 tiles
//   k_text_scan      exclusive scan of the block elements (one CTA; a tile adds its block's tiles in front of it)
//   k_text_expand    a tile's characters: in-tile scan, one quad per glyph at its rank among the text's glyphs, bulk store
// Right to left (Font.cpp:283-327) the reference skips the blanks at the start of a line going forwards, then walks the rest
// of the line BACKWARDS from its last character: the pen before character i is the line's leading blanks plus the advances of
// the characters behind i, and its quad comes after those of the glyphs behind it.  That needs the same scan from the right
// as well (advance / glyphs until the next line break) and the blanks a line starts with; quads then leave one by one (a
// tile's quads are no longer one contiguous run).
// --------------------------------------------------------------------------------------------------
struct TextElem
{
    // from the left
    int adv;              // pen advance since the last line break (or since the start, when there is none)
    uint32_t breaks;      // line breaks
    uint32_t glyphs;      // characters that draw
    uint32_t line_glyphs; // ... since the last line break
    int lead;             // advance of the blanks the current line starts with
    uint32_t open;        // the current line has been blanks only so far
    // from the right
    int radv;             // pen advance until the next line break
    uint32_t rglyphs;     // glyphs until the next line break
    uint32_t rbreak;      // there is a line break further right
    uint32_t pad[3];
};
static_assert(sizeof(TextElem) == 48, "tile elements are 48 bytes");
__device__ __forceinline__ TextElem text_identity()
{
    TextElem e = {};
    e.open = 1u;
    return e;
}
template <bool RTL>
__device__ __forceinline__ TextElem text_combine(const TextElem& a, const TextElem& b) // a before b
{
    TextElem r = {};
    r.adv = b.breaks ? b.adv : a.adv + b.adv;
    r.breaks = a.breaks + b.breaks;
    r.glyphs = a.glyphs + b.glyphs;
    if (RTL)
    {
        r.line_glyphs = b.breaks ? b.line_glyphs : a.line_glyphs + b.line_glyphs;
        // the line b ends in started inside b, or is a's current line continued
        r.lead = b.breaks ? b.lead : (a.open ? a.lead + b.lead : a.lead);
        r.open = b.breaks ? b.open : (a.open && b.open);
        r.radv = a.rbreak ? a.radv : a.radv + b.radv;
        r.rglyphs = a.rbreak ? a.rglyphs : a.rglyphs + b.rglyphs;
        r.rbreak = a.rbreak | b.rbreak;
    }
    return r;
}
template <bool RTL>
__device__ __forceinline__ TextElem text_shfl(const TextElem& v, int delta, int width, bool up)
{
    TextElem t = {};
    auto S = [&](auto x) { return up ? __shfl_up_sync(0xffffffffu, x, delta, width) : __shfl_down_sync(0xffffffffu, x, delta, width); };
    t.adv = S(v.adv);
    t.breaks = S(v.breaks);
    t.glyphs = S(v.glyphs);
    if (RTL)
    {
        t.line_glyphs = S(v.line_glyphs);
        t.lead = S(v.lead);
        t.open = S(v.open);
        t.radv = S(v.radv);
        t.rglyphs = S(v.rglyphs);
        t.rbreak = S(v.rbreak);
    }
    return t;
}
// Tile elements in memory: left to right only the first four words are used (one 128-bit access).
template <bool RTL>
__device__ __forceinline__ TextElem text_load(const TextElem* p)
{
    if (RTL) return *p;
    TextElem e = {};
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    e.adv = (int)v.x; e.breaks = v.y; e.glyphs = v.z; e.line_glyphs = v.w;
    return e;
}
template <bool RTL>
__device__ __forceinline__ void text_store(TextElem* p, const TextElem& e)
{
    if (RTL) *p = e;
    else *reinterpret_cast<uint4*>(p) = make_uint4((uint32_t)e.adv, e.breaks, e.glyphs, e.line_glyphs);
}
#if defined(T3D_EXP_TEXT_TILE)
constexpr int kTextTile = T3D_EXP_TEXT_TILE;
#else
// characters per tile (one expansion CTA): 128 measured 4 % faster than 256 at 8 Mi characters and more (eleven resident CTAs per
// SM instead of five cover each other's load -> scan -> store chains better), the same at 1 Mi
constexpr int kTextTile = 128;
#endif
constexpr int kTextLanes = kTextTile / 16;     // lanes of the summary kernel that fold one tile (16 characters each)
constexpr int kTextBlock = 4096 / kTextTile;   // tiles per block: what one summary CTA (256 threads x 16 characters) covers
static_assert(kTextTile == 128 || kTextTile == 256, "a block's tiles are folded by the lanes of one warp");
// per-character-code tables (TextTables, t3d_internal.h: built once per text on the host, handed over as a kernel parameter):
// every CTA copies them into shared memory, where 256 lanes can look up 256 different codes at once
__device__ __forceinline__ void text_copy_tables(TextTables& dst, const TextTables& src, uint32_t tid, uint32_t threads)
{
    static_assert(sizeof(TextTables) % 4 == 0, "copied word by word");
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(&src);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(&dst);
    for (uint32_t w = tid; w < sizeof(TextTables) / 4; w += threads) d32[w] = s32[w];
}
__device__ __forceinline__ TextElem text_elem(const TextTables& t, unsigned char byte)
{
    TextElem e = {};
    if (byte < (unsigned char)kTextClasses)
    {
        const unsigned char k = t.kind[byte];
        e.adv = e.radv = t.adv[byte];
        e.breaks = e.rbreak = k == 1;
        e.glyphs = e.line_glyphs = e.rglyphs = k == 2;
        e.open = k == 1 || k == 3; // a line break opens a new (so far blank) line; a blank keeps one open
        e.lead = k == 3 ? e.adv : 0;
    }
    return e;
}

template <bool RTL>
__global__ void __launch_bounds__(256) k_text_summary(const unsigned char* __restrict__ text, uint32_t n, const __grid_constant__ TextTables T,
                                                      TextElem* __restrict__ tile_elems, TextElem* __restrict__ block_elems)
{
    __shared__ TextTables s_t;
    text_copy_tables(s_t, T, threadIdx.x, 256);
    __syncthreads();
    const uint32_t first = (blockIdx.x * 256u + threadIdx.x) * 16u; // kTextLanes threads = one tile
    unsigned char b[16];
    if (first + 16u <= n)
    {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + first));
        memcpy(b, &v, 16);
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 16; ++k) b[k] = first + k < n ? text[first + k] : (unsigned char)0xff; // past the end: a byte that does nothing
    }
    TextElem e = text_elem(s_t, b[0]);
#pragma unroll
    for (int k = 1; k < 16; ++k) e = text_combine<RTL>(e, text_elem(s_t, b[k]));
    // fold the 16 lanes of a tile (an inclusive scan; its last lane holds the tile's element)
#pragma unroll
    for (int o = 1; o < kTextLanes; o <<= 1)
    {
        const TextElem t = text_shfl<RTL>(e, o, kTextLanes, true);
        if ((int)(threadIdx.x & (kTextLanes - 1)) >= o) e = text_combine<RTL>(t, e);
    }
    const uint32_t tile = first / (uint32_t)kTextTile;
    if ((threadIdx.x & (kTextLanes - 1)) == kTextLanes - 1 && (unsigned long long)tile * kTextTile < n) text_store<RTL>(tile_elems + tile, e);
    // ... and the CTA's tiles one block of 4 096 characters: the scan kernel only walks the blocks
    __shared__ TextElem s_tiles[kTextBlock];
    if ((threadIdx.x & (kTextLanes - 1)) == kTextLanes - 1) s_tiles[threadIdx.x / kTextLanes] = e;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        TextElem b = s_tiles[0];
#pragma unroll
        for (int k = 1; k < kTextBlock; ++k) b = text_combine<RTL>(b, s_tiles[k]);
        text_store<RTL>(block_elems + blockIdx.x, b);
    }
}

// elems[0 .. n_tiles) -> in place: the fields "from the left" become the exclusive prefix (everything before the tile), the
// fields "from the right" the exclusive suffix (everything behind it); elems[n_tiles] = the whole text.  One CTA.
template <bool RTL>
__global__ void __launch_bounds__(1024) k_text_scan(TextElem* __restrict__ elems, uint32_t n_tiles)
{
    __shared__ TextElem s_warp[32];
    __shared__ TextElem s_carry;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (RTL)
    {
        // from the right first: the suffix fields are parked in the pad words until the forward pass has read the elements
        if (threadIdx.x == 0) s_carry = text_identity();
        __syncthreads();
        const uint32_t chunks = (n_tiles + 1023u) / 1024u;
        for (uint32_t ch = chunks; ch-- > 0;)
        {
            const uint32_t i = ch * 1024u + threadIdx.x;
            const TextElem own = i < n_tiles ? elems[i] : text_identity();
            TextElem incl = own; // fold of [i, end of the warp]
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const TextElem t = text_shfl<true>(incl, o, 32, false);
                if ((int)lane + o < 32) incl = text_combine<true>(incl, t);
            }
            if (lane == 0) s_warp[warp] = incl;
            __syncthreads();
            TextElem behind = text_identity(); // warps behind this one in the chunk, then the chunks behind
            for (uint32_t w = warp + 1; w < 32; ++w) behind = text_combine<true>(behind, s_warp[w]);
            behind = text_combine<true>(behind, s_carry);
            const TextElem next = text_shfl<true>(incl, 1, 32, false);
            const TextElem excl = lane < 31 ? text_combine<true>(next, behind) : behind;
            if (i < n_tiles)
            {
                elems[i].pad[0] = (uint32_t)excl.radv;
                elems[i].pad[1] = excl.rglyphs;
                elems[i].pad[2] = excl.rbreak;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_carry = text_combine<true>(own, excl);
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) s_carry = text_identity();
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024)
    {
        const uint32_t i = base + threadIdx.x;
        const TextElem own = i < n_tiles ? text_load<RTL>(elems + i) : text_identity();
        TextElem incl = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const TextElem t = text_shfl<RTL>(incl, o, 32, true);
            if ((int)lane >= o) incl = text_combine<RTL>(t, incl);
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        TextElem before = s_carry;
        for (uint32_t w = 0; w < warp; ++w) before = text_combine<RTL>(before, s_warp[w]);
        // exclusive = everything before this warp, then the lanes before this one
        const TextElem prev = text_shfl<RTL>(incl, 1, 32, true);
        TextElem excl = lane ? text_combine<RTL>(before, prev) : before;
        const TextElem through = text_combine<RTL>(excl, own);
        if (RTL)
        {
            excl.radv = (int)own.pad[0];
            excl.rglyphs = own.pad[1];
            excl.rbreak = own.pad[2];
        }
        if (i < n_tiles) text_store<RTL>(elems + i, excl);
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = through;
        __syncthreads();
    }
    if (threadIdx.x == 0) text_store<RTL>(elems + n_tiles, s_carry);
}

template <bool RTL>
__global__ void __launch_bounds__(kTextTile) k_text_expand(const unsigned char* __restrict__ text, uint32_t n, const t3d_glyph* __restrict__ glyphs,
                                                           const __grid_constant__ TextTables T, TextParams P,
                                                           const TextElem* __restrict__ tile_elems, uint32_t n_tiles,
                                                           const TextElem* __restrict__ block_around, float* __restrict__ dst)
{
    __shared__ __align__(128) float4 s_buf[RTL ? 9 : kTextTile * 9];
    __shared__ TextTables s_t;
    __shared__ TextElem s_warp[kTextTile / 32];
    __shared__ TextElem s_rwarp[kTextTile / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    text_copy_tables(s_t, T, tid, kTextTile);
    const uint32_t i = blockIdx.x * (uint32_t)kTextTile + tid;
    const unsigned char byte = i < n ? text[i] : (unsigned char)0xff;
    // What lies before the tile (and, right to left, behind it): the scanned element of its block of tiles, then the
    // block's tiles in front of this one (behind it) -- 16 lanes fold them.
    __shared__ TextElem s_around;
    if (warp == 0)
    {
        const uint32_t block = blockIdx.x / kTextBlock, j = blockIdx.x % kTextBlock;
        const uint32_t t = block * kTextBlock + (lane & (kTextBlock - 1));
        const TextElem mine = (lane < (uint32_t)kTextBlock && t < n_tiles) ? text_load<RTL>(tile_elems + t) : text_identity();
        TextElem fwd = mine; // fold of the block's tiles [0, lane]
#pragma unroll
        for (int o = 1; o < kTextBlock; o <<= 1)
        {
            const TextElem u = text_shfl<RTL>(fwd, o, kTextBlock, true);
            if ((int)(lane & (kTextBlock - 1)) >= o) fwd = text_combine<RTL>(u, fwd);
        }
        TextElem bwd = mine; // fold of the block's tiles [lane, 15]
        if (RTL)
        {
#pragma unroll
            for (int o = 1; o < kTextBlock; o <<= 1)
            {
                const TextElem u = text_shfl<true>(bwd, o, kTextBlock, false);
                if ((int)(lane & (kTextBlock - 1)) + o < kTextBlock) bwd = text_combine<true>(bwd, u);
            }
        }
        const TextElem front = text_shfl<RTL>(fwd, 1, kTextBlock, true);   // lane j: fold of the block's tiles [0, j - 1]
        const TextElem back = text_shfl<RTL>(bwd, 1, kTextBlock, false);   // lane j: fold of the block's tiles behind j
        if (lane == j)
        {
            const TextElem blk = text_load<RTL>(block_around + block);
            TextElem a = blk; // (its right-hand fields describe what lies behind the block: set aside)
            a.radv = 0; a.rglyphs = 0; a.rbreak = 0;
            if (j > 0) a = text_combine<RTL>(a, front);
            if (RTL)
            {
                TextElem behind = text_identity();
                behind.radv = blk.radv; behind.rglyphs = blk.rglyphs; behind.rbreak = blk.rbreak;
                if (j + 1u < (uint32_t)kTextBlock) behind = text_combine<true>(back, behind);
                a.radv = behind.radv; a.rglyphs = behind.rglyphs; a.rbreak = behind.rbreak;
            }
            s_around = a;
        }
    }
    __syncthreads();
    const TextElem around = s_around;
    const TextElem own = text_elem(s_t, byte);
    TextElem incl = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const TextElem t = text_shfl<RTL>(incl, o, 32, true);
        if ((int)lane >= o) incl = text_combine<RTL>(t, incl);
    }
    if (lane == 31) s_warp[warp] = incl;
    TextElem rincl = own; // fold of [this character, end of the warp]
    if (RTL)
    {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const TextElem t = text_shfl<true>(rincl, o, 32, false);
            if ((int)lane + o < 32) rincl = text_combine<true>(rincl, t);
        }
        if (lane == 0) s_rwarp[warp] = rincl;
    }
    __syncthreads();
    TextElem before = around;
    before.radv = 0; before.rglyphs = 0; before.rbreak = 0; // (those describe what lies BEHIND the tile)
    TextElem tile_all = before;
#pragma unroll
    for (int w = 0; w < kTextTile / 32; ++w)
    {
        if ((uint32_t)w < warp) before = text_combine<RTL>(before, s_warp[w]);
        tile_all = text_combine<RTL>(tile_all, s_warp[w]);
    }
    const TextElem prev = text_shfl<RTL>(incl, 1, 32, true);
    const TextElem excl = lane ? text_combine<RTL>(before, prev) : before; // the walk before this character
    int pen_x = P.x + excl.adv;                                            // Font.cpp:272, :344-378
    uint32_t quad = excl.glyphs;
    if (RTL)
    {
        TextElem behind = text_identity(); // the warps behind this one in the tile, then what lies behind the tile
#pragma unroll
        for (int w = 0; w < kTextTile / 32; ++w)
            if ((uint32_t)w > warp) behind = text_combine<true>(behind, s_rwarp[w]);
        TextElem after_tile = text_identity();
        after_tile.radv = around.radv; after_tile.rglyphs = around.rglyphs; after_tile.rbreak = around.rbreak;
        behind = text_combine<true>(behind, after_tile);
        const TextElem next = text_shfl<true>(rincl, 1, 32, false);
        const TextElem rexcl = lane < 31 ? text_combine<true>(next, behind) : behind;
        // :283-327: the blanks the line starts with, then the characters behind this one (the line is walked backwards)
        pen_x = P.x + excl.lead + rexcl.radv;
        quad = (excl.glyphs - excl.line_glyphs) + rexcl.rglyphs;
    }
    float4 q[9];
    if (own.glyphs)
    {
        const t3d_glyph g = glyphs[byte - 32u];
        const int pen_y = (int)((uint32_t)P.y + excl.breaks * P.size);      // :348 (yPos += size, an unsigned added to an int)
        const float x = (float)(pen_x + __float2int_rz((float)g.bearing_x * P.scale)); // :369
        const float y = (float)pen_y;
        const float w = __uint2float_rn(g.width) * P.scale;                 // :371
        const float h = __uint2float_rn(P.size);                            // :372
        const float x2 = x + w, y2 = y + h;                                 // SB.cpp:493-494
        const float px[4] = { x, x, x2, x2 }, py[4] = { y, y2, y, y2 };     // SB.cpp:496-501
        if (RTL)
        {
            write_quad(q, px, py, 0.0f, g.uvs[0], g.uvs[1], g.uvs[2], g.uvs[3], P.color);
            float4* out = reinterpret_cast<float4*>(dst + (size_t)quad * kQuadFloats);
#pragma unroll
            for (int k = 0; k < 9; ++k) out[k] = q[k];
        }
        else
            write_quad(s_buf + (quad - around.glyphs) * 9, px, py, 0.0f, g.uvs[0], g.uvs[1], g.uvs[2], g.uvs[3], P.color);
    }
    if (RTL) return;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t tile_glyphs = tile_all.glyphs - around.glyphs;
    if (tid == 0 && tile_glyphs)
    {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)around.glyphs * kQuadFloats),
                     "r"((uint32_t)__cvta_generic_to_shared(s_buf)), "r"(tile_glyphs * (uint32_t)(kQuadFloats * 4)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// --------------------------------------------------------------------------------------------------
// launch wrappers
// --------------------------------------------------------------------------------------------------
uint32_t text_tiles(uint32_t n) { return (n + kTextTile - 1) / kTextTile; }
static uint32_t text_blocks(uint32_t n) { return (text_tiles(n) + kTextBlock - 1) / kTextBlock; }
// tile elements, then block elements, then the whole text's element
size_t text_scratch_bytes(uint32_t n) { return ((size_t)text_tiles(n) + text_blocks(n) + 1) * sizeof(TextElem); }

template <bool RTL>
static void launch_text_dir(cudaStream_t st, const unsigned char* t, uint32_t n, const t3d_glyph* glyphs, const TextTables& T, const TextParams& P,
                            TextElem* elems, float* dst)
{
    const uint32_t tiles = text_tiles(n), blocks = text_blocks(n);
    TextElem* block_elems = elems + tiles;
    k_text_summary<RTL><<<blocks, 256, 0, st>>>(t, n, T, elems, block_elems);
    k_text_scan<RTL><<<1, 1024, 0, st>>>(block_elems, blocks);
    k_text_expand<RTL><<<tiles, kTextTile, 0, st>>>(t, n, glyphs, T, P, elems, tiles, block_elems, dst);
}
void launch_text(void* stream, const char* text, uint32_t n, const t3d_glyph* glyphs, const TextTables& T, const TextParams& P, bool right_to_left,
                 void* scratch, float* dst)
{
    if (!n) return;
    const unsigned char* t = reinterpret_cast<const unsigned char*>(text);
    if (right_to_left) launch_text_dir<true>((cudaStream_t)stream, t, n, glyphs, T, P, static_cast<TextElem*>(scratch), dst);
    else launch_text_dir<false>((cudaStream_t)stream, t, n, glyphs, T, P, static_cast<TextElem*>(scratch), dst);
}

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
