// records.cu -- micro-benchmark for DESIGN §9(2): would keeping a particle's 19 read-only words contiguous (an 80-byte
// record, so that a moved particle costs 3 sectors instead of 19) slow the streaming path down?
// Same skeleton as the update kernel's phase B: 96 threads x 4 particles per group, cp.async ring with two groups in
// flight, 2 CTAs/SM, 15 columns written with 128-bit streaming stores.  No scan, no look-back, trivial arithmetic.
//   A  25 input columns, struct-of-arrays (what ships)                       100 B read + 60 B written per particle
//   B  10 input columns struct-of-arrays + one 80-byte record per particle   120 B read + 60 B written per particle
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o records.bin records.cu && ./records.bin
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void cp16(void* s, const void* g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int THREADS = 96, SUB = 11, WRITTEN = 15;

// MODE 0: A.  MODE 1: B, every thread copies its own four records (lane stride 320 B).  MODE 2: B, the warp copies its
// 10 KB of records cooperatively (each instruction covers 512 contiguous bytes) and every thread then reads its own
// records back from shared memory (4-way bank conflicts on those reads).
template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) k(const float* __restrict__ cols, size_t stride, const float4* __restrict__ rec, float* __restrict__ out,
                                                 uint32_t n)
{
    constexpr bool RECORD = MODE != 0;
    constexpr int SOA = RECORD ? 10 : 25, REC = RECORD ? 20 : 0, SLOTS = SOA + REC;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    extern __shared__ __align__(128) unsigned char smem[];
    float4* ring = reinterpret_cast<float4*>(smem);
    const uint32_t tid = threadIdx.x;
    const uint32_t first = blockIdx.x * (THREADS * 4 * SUB);
    auto slot = [&](int stage, int c) { return ring + ((size_t)(stage * SLOTS + c) * THREADS + tid); };
    auto issue = [&](int stage, uint32_t k0)
    {
#pragma unroll
        for (int c = 0; c < SOA; ++c) cp16(slot(stage, c), cols + (size_t)c * stride + k0);
        if (MODE == 1)
        {
#pragma unroll
            for (int j = 0; j < REC; ++j) cp16(slot(stage, SOA + j), rec + (size_t)k0 * 5 + j); // 4 particles x 80 B = 20 x 16 B, contiguous
        }
        if (MODE == 2)
        {
            // the warp's records: 32 lanes x 20 chunks, contiguous in global memory and kept in that order in shared memory
            const uint32_t kw = k0 - lane * 4; // first particle of the warp's group
            float4* const wbase = ring + ((size_t)(stage * SLOTS + SOA) * THREADS) + (size_t)warp * 32 * REC;
#pragma unroll
            for (int j = 0; j < REC; ++j) cp16(wbase + j * 32 + lane, rec + (size_t)kw * 5 + j * 32 + lane);
        }
    };
#pragma unroll
    for (int s = 0; s < 2; ++s)
    {
        const uint32_t k0 = first + s * THREADS * 4 + tid * 4;
        if (k0 < n) issue(s, k0);
        commit();
    }
#pragma unroll 1
    for (int s = 0; s < SUB; ++s)
    {
        const uint32_t k0 = first + s * THREADS * 4 + tid * 4;
        if (k0 >= n) break;
        if (s + 1 < SUB) wait<1>(); else wait<0>();
        if (MODE == 2) __syncwarp(); // the other lanes' copies of this stage have landed too
        float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < (MODE == 2 ? SOA : SLOTS); ++c)
        {
            const float4 v = *slot(s & 1, c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (MODE == 2)
        {
            const float4* const mine = ring + ((size_t)((s & 1) * SLOTS + SOA) * THREADS) + (size_t)warp * 32 * REC + lane * REC;
#pragma unroll
            for (int j = 0; j < REC; ++j)
            {
                const float4 v = mine[j];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            __syncwarp(); // everybody has read the stage before anybody refills it
        }
        if (s + 2 < SUB)
        {
            const uint32_t kn = first + (s + 2) * THREADS * 4 + tid * 4;
            if (kn < n) issue(s & 1, kn);
            commit();
        }
#pragma unroll
        for (int c = 0; c < WRITTEN; ++c) __stcs(reinterpret_cast<float4*>(out + (size_t)c * stride + k0), acc);
    }
    wait<0>();
}

template <int MODE>
static void run(const char* name, const float* cols, size_t stride, const float4* rec, float* out, uint32_t n, double bytes_per_particle)
{
    constexpr int SLOTS = MODE ? 30 : 25;
    const size_t smem = 2 * (size_t)SLOTS * THREADS * sizeof(float4);
    CHECK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t tile = THREADS * 4 * SUB, grid = (n + tile - 1) / tile;
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) k<MODE><<<grid, THREADS, smem>>>(cols, stride, rec, out, n);
    CHECK(cudaEventRecord(a));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) k<MODE><<<grid, THREADS, smem>>>(cols, stride, rec, out, n);
    CHECK(cudaEventRecord(b));
    CHECK(cudaEventSynchronize(b));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, a, b));
    ms /= reps;
    printf("%-44s %7.3f ms  %7.1f GB/s  (%.0f B/particle)\n", name, ms, bytes_per_particle * n / (ms * 1e-3) / 1e9, bytes_per_particle);
}

int main()
{
    const uint32_t n = 64u << 20;
    const size_t stride = n;
    float *cols, *out;
    float4* rec;
    CHECK(cudaMalloc(&cols, (size_t)25 * stride * sizeof(float)));
    CHECK(cudaMalloc(&out, (size_t)WRITTEN * stride * sizeof(float)));
    CHECK(cudaMalloc(&rec, (size_t)n * 80));
    CHECK(cudaMemset(cols, 0, (size_t)25 * stride * sizeof(float)));
    CHECK(cudaMemset(rec, 0, (size_t)n * 80));
    run<0>("A: 25 input columns (struct-of-arrays)", cols, stride, rec, out, n, 160.0);
    run<1>("B1: 10 columns + records, per-thread copies", cols, stride, rec, out, n, 180.0);
    run<2>("B2: 10 columns + records, warp-coalesced", cols, stride, rec, out, n, 180.0);
    run<0>("A again", cols, stride, rec, out, n, 160.0);
    CHECK(cudaDeviceSynchronize());
    return 0;
}
