// t3d_device.cuh -- device helpers shared by the kernels: the reference's scalar math restated with its
// evaluation order (PE.cpp / Matrix.cpp / MathUtil.h / pch.h of the reference, see t3d_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "t3d_internal.h"

namespace t3d
{

// --------------------------------------------------------------------------------------------------
// helpers
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ float u01(int32_t r)
{
    // pch.h:152: (float)rand()/RAND_MAX with RAND_MAX = 2^31-1 converted to float (= 2^31): exactly r * 2^-31.
    return (float)r / 2147483648.0f;
}
__device__ __forceinline__ float u11(int32_t r) { return (2.0f * u01(r)) - 1.0f; } // pch.h:151

// MathUtil.h:195-200, left-to-right sum of four products.
__device__ __forceinline__ void transform4(const float* __restrict__ m, float x, float y, float z, float w, float& ox,
                                           float& oy, float& oz)
{
    ox = x * m[0] + y * m[4] + z * m[8] + w * m[12];
    oy = x * m[1] + y * m[5] + z * m[9] + w * m[13];
    oz = x * m[2] + y * m[6] + z * m[10] + w * m[14];
}

// Matrix.cpp:348-406.  Only the 3x3 part is produced (m[12..14] are 0 and are added as such by callers).
struct Rot3
{
    float m0, m1, m2, m4, m5, m6, m8, m9, m10;
};
__device__ __forceinline__ Rot3 create_rotation(float x, float y, float z, float angle)
{
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
    float s, c;
    sincosf(angle, &s, &c);
    float t = 1.0f - c;
    float tx = t * x, ty = t * y, tz = t * z;
    float txy = tx * y, txz = tx * z, tyz = ty * z;
    float sx = s * x, sy = s * y, sz = s * z;
    Rot3 r;
    r.m0 = c + tx * x;
    r.m1 = txy + sz;
    r.m2 = txz - sy;
    r.m4 = txy - sz;
    r.m5 = c + ty * y;
    r.m6 = tyz + sx;
    r.m8 = txz + sy;
    r.m9 = tyz - sx;
    r.m10 = c + tz * z;
    return r;
}
// transformVector4 with a Rot3 (m[12..14] = 0): keeps the "+ w*0" term so signed zeros match.
__device__ __forceinline__ void rot_apply(const Rot3& r, float w, float& x, float& y, float& z)
{
    float ox = x * r.m0 + y * r.m4 + z * r.m8 + w * 0.0f;
    float oy = x * r.m1 + y * r.m5 + z * r.m9 + w * 0.0f;
    float oz = x * r.m2 + y * r.m6 + z * r.m10 + w * 0.0f;
    x = ox;
    y = oy;
    z = oz;
}

// The latched rotation (3x3 part), fetched once per thread.
__device__ __forceinline__ Rot3 load_frozen(const FrozenRotation* __restrict__ f)
{
    Rot3 r;
    r.m0 = f->m[0]; r.m1 = f->m[1]; r.m2 = f->m[2];
    r.m4 = f->m[4]; r.m5 = f->m[5]; r.m6 = f->m[6];
    r.m8 = f->m[8]; r.m9 = f->m[9]; r.m10 = f->m[10];
    return r;
}

// One billboard quad in SpriteBatch's vertex layout (SB.cpp:292-363 through PE.cpp:866-882): 4 vertices x
// {position, uv, colour} = 36 floats written as nine float4.  rotationPoint = (0.5, 0.5) (PE.cpp:855).
// get_frozen() yields the latched rotation; it is only called for a rotated quad in T3D_ROT_FROZEN mode.
template <class GetFrozen>
__device__ __forceinline__ void expand_quad(const float pos[3], const float color[4], float size, float angle, float u1,
                                            float v1, float u2, float v2, const float right[3], const float up[3],
                                            int rot_mode, GetFrozen get_frozen, float4* __restrict__ o)
{
    const float width = size, height = size;
    // SB.cpp:306-324
    float p0[3], p1[3], p2[3], p3[3], tfw[3];
    const float hw = width * 0.5f, hh = height * 0.5f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const float tr = right[k] * hw;
        const float tf = up[k] * hh;
        p0[k] = pos[k];
        p0[k] -= tr;
        p0[k] -= tf;
        p1[k] = pos[k];
        p1[k] += tr;
        p1[k] -= tf;
        tfw[k] = up[k] * height;
        p2[k] = p0[k] + tfw[k];
        p3[k] = p1[k] + tfw[k];
    }
    if (angle != 0) // SB.cpp:327
    {
        float rp[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const float tr = right[k] * (width * 0.5f); // rotationPoint.x = 0.5 (PE.cpp:855)
            const float tf = tfw[k] * 0.5f;
            rp[k] = p0[k];
            rp[k] += tr;
            rp[k] += tf;
        }
        Rot3 r;
        if (rot_mode == T3D_ROT_FROZEN)
            r = get_frozen(); // SB.cpp:339: the matrix latched by the first rotated quad of the process
        else
        {
            // MathUtil.h:216-225
            const float ux = (right[1] * up[2]) - (right[2] * up[1]);
            const float uy = (right[2] * up[0]) - (right[0] * up[2]);
            const float uz = (right[0] * up[1]) - (right[1] * up[0]);
            r = create_rotation(ux, uy, uz, angle);
        }
        float* ps[4] = { p0, p1, p2, p3 };
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            float* p = ps[q];
            p[0] -= rp[0];
            p[1] -= rp[1];
            p[2] -= rp[2];
            rot_apply(r, 0.0f, p[0], p[1], p[2]); // Vector3 *= Matrix: transformVector, w = 0
            p[0] += rp[0];
            p[1] += rp[1];
            p[2] += rp[2];
        }
    }
    // SB.cpp:355-359: v0=(p0,u1,v1) v1=(p1,u2,v1) v2=(p2,u1,v2) v3=(p3,u2,v2), particle colour on all four.
    o[0] = make_float4(p0[0], p0[1], p0[2], u1);
    o[1] = make_float4(v1, color[0], color[1], color[2]);
    o[2] = make_float4(color[3], p1[0], p1[1], p1[2]);
    o[3] = make_float4(u2, v1, color[0], color[1]);
    o[4] = make_float4(color[2], color[3], p2[0], p2[1]);
    o[5] = make_float4(p2[2], u1, v2, color[0]);
    o[6] = make_float4(color[1], color[2], color[3], p3[0]);
    o[7] = make_float4(p3[1], p3[2], u2, v2);
    o[8] = make_float4(color[0], color[1], color[2], color[3]);
}

// Address of particle i's word in column 0 of an emitter; its word in column c lies c * stride words further.
//   block_shift == 0  plain struct-of-arrays: stride = slab capacity.
//   block_shift == k  blocked struct-of-arrays: 2^k consecutive particles keep their 34 columns in one contiguous
//                     chunk (stride = 2^k), so a tile's working set is one contiguous region of HBM.
// Groups of 4 consecutive particles (index multiple of 4) never straddle a block.
__device__ __forceinline__ uint32_t* particle_base(uint32_t* cols, uint32_t block_shift, uint32_t i)
{
    if (block_shift == 0) return cols + i;
    const uint32_t blk = i >> block_shift, lane = i & ((1u << block_shift) - 1u);
    return cols + ((size_t)blk * T3D_NUM_COLUMNS << block_shift) + lane;
}
__device__ __forceinline__ const uint32_t* particle_base(const uint32_t* cols, uint32_t block_shift, uint32_t i)
{
    return particle_base(const_cast<uint32_t*>(cols), block_shift, i);
}

// Which emitter of the launch owns global tile `tile`: the last item with tile_begin <= tile.  Every step is a
// dependent load, so the search starts from the interpolated position (exact when emitters have equal tile
// counts: two loads) and gallops outwards before bisecting.
template <class Item>
__device__ __forceinline__ uint32_t find_item(const Item* __restrict__ items, uint32_t n_items, uint32_t tile,
                                              uint32_t total_tiles)
{
    if (n_items == 1) return 0;
    uint32_t lo, hi;
    uint32_t g = (uint32_t)(((unsigned long long)tile * n_items) / (total_tiles ? total_tiles : 1u));
    if (g >= n_items) g = n_items - 1;
    if (items[g].tile_begin <= tile)
    {
        lo = g;
        uint32_t step = 1;
        while (lo + step < n_items && items[lo + step].tile_begin <= tile)
        {
            lo += step;
            step <<= 1;
        }
        hi = lo + step < n_items ? lo + step : n_items;
    }
    else
    {
        hi = g;
        uint32_t step = 1;
        while (hi > step && items[hi - step].tile_begin > tile)
        {
            hi -= step;
            step <<= 1;
        }
        lo = hi > step ? hi - step : 0;
    }
    while (hi - lo > 1)
    {
        uint32_t mid = (lo + hi) >> 1;
        if (items[mid].tile_begin <= tile)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

} // namespace t3d
