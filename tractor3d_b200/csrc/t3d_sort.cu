// t3d_sort.cu -- SURVEY 8(f4): back-to-front order for alpha-blended emitters (BLEND_ALPHA, PE.h:157-163).  The reference
// draws its particles in array order (PE.cpp:866-882), which is only right for additive blending; this produces an index
// buffer that walks the emitter's quads farthest first.
//
//   k_depth_keys                 key_i = ~order(position_i . view_dir)   (largest depth first, ascending radix sort)
//   cub::DeviceRadixSort         32-bit keys, quad numbers as values; stable, so quads of equal depth keep live-array order.
//                                This is the CUDA toolkit's library sort (like calling cuBLAS for a plain GEMM): the row is
//                                not on the per-frame hot path and a radix sort has no particle-specific structure to exploit.
//   k_sorted_triangle_indices    {4v, 4v+1, 4v+2, 4v+2, 4v+1, 4v+3} per quad v in sorted order (the triangle-list form of
//                                MeshBatch's strip, see k_triangle_indices)
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include "t3d_device.cuh"
#include "t3d_internal.h"

namespace t3d
{

__global__ void __launch_bounds__(256) k_depth_keys(const uint32_t* __restrict__ rw, size_t stride, uint32_t n, float vx, float vy, float vz,
                                                    uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const float px = __uint_as_float(__ldcs(rw + (size_t)(RW_POS + 0) * stride + i));
    const float py = __uint_as_float(__ldcs(rw + (size_t)(RW_POS + 1) * stride + i));
    const float pz = __uint_as_float(__ldcs(rw + (size_t)(RW_POS + 2) * stride + i));
    const float d = px * vx + py * vy + pz * vz; // left to right, unfused (Vector3::dot, src/math/Vector3.cpp)
    keys[i] = ~float_order(d);
    vals[i] = i;
}

__global__ void __launch_bounds__(256) k_sorted_triangle_indices(const uint32_t* __restrict__ order, uint64_t n_indices, uint32_t* __restrict__ out)
{
    const uint64_t j = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (j >= n_indices) return;
    const uint32_t v = order[j / 6u];
    const uint32_t corner = (uint32_t)(j % 6u);
    // quad v = strip {4v, 4v+1, 4v+2, 4v+3}: triangles (0,1,2) and (2,1,3)
    const uint32_t pat = corner < 3u ? corner : (corner == 3u ? 2u : (corner == 4u ? 1u : 3u));
    out[j] = 4u * v + pat;
}

size_t depth_sort_temp_bytes(uint32_t n)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (int)n);
    return bytes;
}

// scratch: 4 * n words (keys in / out, values in / out); temp: depth_sort_temp_bytes(n); out: 6 * n indices
int launch_depth_sort(void* stream, const uint32_t* rw, size_t stride, uint32_t n, const float view_dir[3], uint32_t* scratch, void* temp,
                      size_t temp_bytes, uint32_t* out)
{
    if (!n) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* keys_in = scratch;
    uint32_t* keys_out = scratch + n;
    uint32_t* vals_in = scratch + 2 * (size_t)n;
    uint32_t* vals_out = scratch + 3 * (size_t)n;
    k_depth_keys<<<(n + 255u) / 256u, 256, 0, st>>>(rw, stride, n, view_dir[0], view_dir[1], view_dir[2], keys_in, vals_in);
    const cudaError_t err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 32, st);
    if (err != cudaSuccess) return (int)err;
    const uint64_t n_indices = 6ull * n;
    k_sorted_triangle_indices<<<(unsigned)((n_indices + 255u) / 256u), 256, 0, st>>>(vals_out, n_indices, out);
    return 0;
}

} // namespace t3d
