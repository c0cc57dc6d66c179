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
