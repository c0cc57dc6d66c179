// streams.cu -- micro-benchmark: what bandwidth does the update kernel's ACCESS PATTERN allow on B200?
// 34 columns of N floats (struct-of-arrays).  Each thread owns 4 consecutive particles per group; a CTA streams
// SUB groups.  Per particle it reads RO + RW columns and writes RW + WO columns (update: 15 RO, 10 RW, 5 WO).
// No scan, no look-back, trivial arithmetic: the ceiling for this many concurrent streams.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o streams streams.cu && ./streams
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int RO, int RW, int WO, int THREADS, int SUB, int MINB, bool TWO_PHASE>
__global__ void __launch_bounds__(THREADS, MINB) k_stream(float* cols, size_t stride, uint32_t n)
{
    const uint32_t first = blockIdx.x * (THREADS * 4 * SUB);
#pragma unroll 1
    for (int s = 0; s < SUB; ++s)
    {
        const uint32_t k0 = first + s * THREADS * 4 + threadIdx.x * 4;
        if (k0 >= n) return;
        float* g = cols + k0;
        float4 ro[RO ? RO : 1], rw[RW ? RW : 1];
        float4 acc = make_float4(0, 0, 0, 0);
        if (!TWO_PHASE)
        {
#pragma unroll
            for (int c = 0; c < RO; ++c) ro[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)c * stride));
#pragma unroll
            for (int c = 0; c < RW; ++c) rw[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)(RO + c) * stride));
#pragma unroll
            for (int c = 0; c < RO; ++c) { acc.x += ro[c].x; acc.y += ro[c].y; acc.z += ro[c].z; acc.w += ro[c].w; }
#pragma unroll
            for (int c = 0; c < RW; ++c)
            {
                rw[c].x += acc.x * 1e-9f; rw[c].y += acc.y * 1e-9f; rw[c].z += acc.z * 1e-9f; rw[c].w += acc.w * 1e-9f;
                __stcs(reinterpret_cast<float4*>(g + (size_t)(RO + c) * stride), rw[c]);
            }
#pragma unroll
            for (int c = 0; c < WO; ++c) __stcs(reinterpret_cast<float4*>(g + (size_t)(RO + RW + c) * stride), acc);
        }
        else
        {
            // two halves with a compiler fence in between, like the shipped kernel
            constexpr int RO1 = RO / 2, RW1 = RW / 2;
#pragma unroll
            for (int c = 0; c < RO1; ++c) ro[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)c * stride));
#pragma unroll
            for (int c = 0; c < RW1; ++c) rw[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)(RO + c) * stride));
#pragma unroll
            for (int c = 0; c < RO1; ++c) { acc.x += ro[c].x; acc.y += ro[c].y; acc.z += ro[c].z; acc.w += ro[c].w; }
#pragma unroll
            for (int c = 0; c < RW1; ++c)
            {
                rw[c].x += acc.x * 1e-9f;
                __stcs(reinterpret_cast<float4*>(g + (size_t)(RO + c) * stride), rw[c]);
            }
            asm volatile("" ::: "memory");
#pragma unroll
            for (int c = RO1; c < RO; ++c) ro[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)c * stride));
#pragma unroll
            for (int c = RW1; c < RW; ++c) rw[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)(RO + c) * stride));
#pragma unroll
            for (int c = RO1; c < RO; ++c) { acc.x += ro[c].x; acc.y += ro[c].y; acc.z += ro[c].z; acc.w += ro[c].w; }
#pragma unroll
            for (int c = RW1; c < RW; ++c)
            {
                rw[c].x += acc.x * 1e-9f;
                __stcs(reinterpret_cast<float4*>(g + (size_t)(RO + c) * stride), rw[c]);
            }
#pragma unroll
            for (int c = 0; c < WO; ++c) __stcs(reinterpret_cast<float4*>(g + (size_t)(RO + RW + c) * stride), acc);
            asm volatile("" ::: "memory");
        }
    }
}

// k_stream + the update kernel's control skeleton, added feature by feature:
//   PHASEA  load column 0 of all SUB groups first, reduce a flag count over the CTA (warp shuffles + one barrier)
//   LOOK    decoupled look-back over per-tile status words (publish aggregate, poll predecessors, publish prefix) + barrier
//   ROLLED  the group loop is not unrolled
template <int THREADS, int SUB, int MINB, bool PHASEA, bool LOOK, bool ROLLED>
__global__ void __launch_bounds__(THREADS, MINB) k_skel(float* cols, size_t stride, uint32_t n, unsigned long long* status, uint32_t epoch)
{
    constexpr int RO = 15, RW = 10, WO = 5;
    __shared__ uint32_t s_warp[THREADS / 32];
    __shared__ uint32_t s_excl;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t first = tile * (THREADS * 4 * SUB);
    float bias = 0.f;
    if (PHASEA)
    {
        float4 e[SUB];
#pragma unroll
        for (int s = 0; s < SUB; ++s)
        {
            const uint32_t k0 = first + s * THREADS * 4 + tid * 4;
            e[s] = k0 < n ? __ldcs(reinterpret_cast<const float4*>(cols + (size_t)(RO) * stride + k0)) : make_float4(1, 1, 1, 1);
        }
        uint32_t cnt = 0;
#pragma unroll
        for (int s = 0; s < SUB; ++s) cnt += (e[s].x < -1e30f) + (e[s].y < -1e30f) + (e[s].z < -1e30f) + (e[s].w < -1e30f);
        uint32_t v = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        uint32_t tile_dead = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) tile_dead += s_warp[w];
        if (LOOK)
        {
            const unsigned long long tag = (unsigned long long)epoch << 34;
            if (tile == 0) { if (tid == 0) { status[0] = tag | (2ull << 32) | tile_dead; s_excl = 0; } }
            else if (warp == 0)
            {
                if (lane == 0) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(status + tile), "l"(tag | (1ull << 32) | tile_dead) : "memory"); }
                uint32_t excl = 0; int look = (int)tile - 1;
                while (true)
                {
                    const int idx = look - (int)lane;
                    unsigned long long st = tag | (2ull << 32);
                    if (idx >= 0)
                    {
                        do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(st) : "l"(status + idx) : "memory"); }
                        while ((st >> 34) != epoch || ((st >> 32) & 3ull) == 0);
                    }
                    const unsigned pm = __ballot_sync(0xffffffffu, ((st >> 32) & 3ull) == 2ull);
                    const int fp = __ffs((int)pm) - 1;
                    uint32_t c = (pm == 0 || (int)lane <= fp) ? (uint32_t)st : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    excl += c;
                    if (pm) break;
                    look -= 32;
                }
                if (lane == 0) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(status + tile), "l"(tag | (2ull << 32) | (excl + tile_dead)) : "memory"); s_excl = excl; }
            }
            __syncthreads();
            bias = (float)s_excl;
        }
        else
            bias = (float)tile_dead;
    }
#pragma unroll(ROLLED ? 1 : SUB)
    for (int s = 0; s < SUB; ++s)
    {
        const uint32_t k0 = first + s * THREADS * 4 + tid * 4;
        if (k0 >= n) return;
        float* g = cols + k0;
        float4 ro[RO], rw[RW];
        float4 acc = make_float4(bias, 0, 0, 0);
        constexpr int RO1 = RO / 2, RW1 = RW / 2;
#pragma unroll
        for (int c = 0; c < RO1; ++c) ro[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)c * stride));
#pragma unroll
        for (int c = 0; c < RW1; ++c) rw[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)(RO + c) * stride));
#pragma unroll
        for (int c = 0; c < RO1; ++c) { acc.x += ro[c].x; acc.y += ro[c].y; acc.z += ro[c].z; acc.w += ro[c].w; }
#pragma unroll
        for (int c = 0; c < RW1; ++c) { rw[c].x += acc.x * 1e-9f; __stcs(reinterpret_cast<float4*>(g + (size_t)(RO + c) * stride), rw[c]); }
        asm volatile("" ::: "memory");
#pragma unroll
        for (int c = RO1; c < RO; ++c) ro[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)c * stride));
#pragma unroll
        for (int c = RW1; c < RW; ++c) rw[c] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)(RO + c) * stride));
#pragma unroll
        for (int c = RO1; c < RO; ++c) { acc.x += ro[c].x; acc.y += ro[c].y; acc.z += ro[c].z; acc.w += ro[c].w; }
#pragma unroll
        for (int c = RW1; c < RW; ++c) { rw[c].x += acc.x * 1e-9f; __stcs(reinterpret_cast<float4*>(g + (size_t)(RO + c) * stride), rw[c]); }
#pragma unroll
        for (int c = 0; c < WO; ++c) __stcs(reinterpret_cast<float4*>(g + (size_t)(RO + RW + c) * stride), acc);
        asm volatile("" ::: "memory");
    }
}

template <int THREADS, int SUB, int MINB, bool PHASEA, bool LOOK, bool ROLLED>
static void run_skel(float* cols, size_t stride, uint32_t n, unsigned long long* status)
{
    const uint32_t tile = THREADS * 4 * SUB;
    const uint32_t grid = (n + tile - 1) / tile;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    static uint32_t epoch = 1;
    for (int i = 0; i < 3; ++i) k_skel<THREADS, SUB, MINB, PHASEA, LOOK, ROLLED><<<grid, THREADS>>>(cols, stride, n, status, epoch++);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) k_skel<THREADS, SUB, MINB, PHASEA, LOOK, ROLLED><<<grid, THREADS>>>(cols, stride, n, status, epoch++);
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double bytes = (double)n * 4.0 * 40 + (PHASEA ? (double)n * 4.0 : 0.0);
    printf("skeleton thr %3d sub %d minb %d phaseA %d lookback %d rolled %d : %.3f ms  %.0f GB/s (incl. the %s)\n", THREADS, SUB, MINB,
           PHASEA, LOOK, ROLLED, ms, bytes / ms / 1e6, PHASEA ? "extra 4 B/particle phase-A re-read" : "160 B/particle");
}

__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n4)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <int RO, int RW, int WO, int THREADS, int SUB, int MINB, bool TWO>
static void run(const char* name, float* cols, size_t stride, uint32_t n)
{
    const uint32_t tile = THREADS * 4 * SUB;
    const uint32_t grid = (n + tile - 1) / tile;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k_stream<RO, RW, WO, THREADS, SUB, MINB, TWO><<<grid, THREADS>>>(cols, stride, n);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) k_stream<RO, RW, WO, THREADS, SUB, MINB, TWO><<<grid, THREADS>>>(cols, stride, n);
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double bytes = (double)n * 4.0 * (RO + 2 * RW + WO);
    printf("%-34s RO %2d RW %2d WO %2d  thr %3d sub %d minb %d %s : %.3f ms  %.0f GB/s\n", name, RO, RW, WO, THREADS, SUB, MINB,
           TWO ? "2ph" : "1ph", ms, bytes / ms / 1e6);
}

int main()
{
    const uint32_t n = 64u << 20;
    const size_t stride = n;
    float* cols;
    CHECK(cudaMalloc(&cols, (size_t)34 * stride * sizeof(float)));
    CHECK(cudaMemset(cols, 0, (size_t)34 * stride * sizeof(float)));
    {
        float4 *a = (float4*)cols, *b = (float4*)(cols + (size_t)17 * stride);
        size_t n4 = (size_t)16 * stride / 4;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 2; ++i) k_copy<<<148 * 16, 512>>>(a, b, n4);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) k_copy<<<148 * 16, 512>>>(a, b, n4);
        cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("grid-stride float4 copy of %.1f GB: %.3f ms  %.0f GB/s (read+write)\n", n4 * 16.0 / 1e9, ms, 2.0 * n4 * 16 / ms / 1e6);
        CHECK(cudaMemcpy(b, a, n4 * 16, cudaMemcpyDeviceToDevice));
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) cudaMemcpyAsync(b, a, n4 * 16, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("cudaMemcpy D2D: %.3f ms  %.0f GB/s (read+write)\n", ms, 2.0 * n4 * 16 / ms / 1e6);
    }
    // the update kernel's pattern: 15 RO + 10 RW + 5 WO = 160 B/particle
    run<15, 10, 5, 128, 4, 4, false>("update pattern", cols, stride, n);
    run<15, 10, 5, 128, 4, 4, true>("update pattern", cols, stride, n);
    run<15, 10, 5, 128, 4, 3, false>("update pattern", cols, stride, n);
    run<15, 10, 5, 256, 2, 2, true>("update pattern", cols, stride, n);
    run<15, 10, 5, 256, 4, 2, true>("update pattern", cols, stride, n);
    run<15, 10, 5, 128, 1, 4, true>("update pattern", cols, stride, n);
    run<15, 10, 5, 128, 8, 5, true>("update pattern", cols, stride, n);
    run<15, 10, 5, 64, 4, 8, true>("update pattern", cols, stride, n);
    {
        unsigned long long* status;
        CHECK(cudaMalloc(&status, sizeof(unsigned long long) * (1 << 20)));
        CHECK(cudaMemset(status, 0, sizeof(unsigned long long) * (1 << 20)));
        run_skel<128, 4, 4, false, false, false>(cols, stride, n, status);
        run_skel<128, 4, 4, false, false, true>(cols, stride, n, status);
        run_skel<128, 4, 4, true, false, false>(cols, stride, n, status);
        run_skel<128, 4, 4, true, true, false>(cols, stride, n, status);
        run_skel<128, 4, 4, true, true, true>(cols, stride, n, status);
        run_skel<128, 1, 4, true, true, false>(cols, stride, n, status);
        run_skel<128, 2, 4, true, true, false>(cols, stride, n, status);
        run_skel<128, 8, 4, true, true, true>(cols, stride, n, status);
        run_skel<256, 4, 2, true, true, false>(cols, stride, n, status);
        run_skel<64, 8, 8, true, true, true>(cols, stride, n, status);
    }
    // fewer, fatter streams with the same bytes per particle
    run<16, 0, 16, 128, 4, 4, false>("16 in / 16 out (copy-like)", cols, stride, n);
    run<8, 0, 8, 128, 4, 8, false>("8 in / 8 out", cols, stride, n);
    run<2, 0, 2, 128, 4, 8, false>("2 in / 2 out", cols, stride, n);
    run<1, 0, 1, 256, 4, 8, false>("1 in / 1 out", cols, stride, n);
    run<25, 0, 0, 128, 4, 4, false>("25 in / 0 out (read only)", cols, stride, n);
    run<0, 0, 15, 128, 4, 8, false>("0 in / 15 out (write only)", cols, stride, n);
    run<0, 15, 0, 128, 4, 8, false>("15 in-place rw", cols, stride, n);
    run<10, 0, 1, 128, 4, 8, false>("10 in / 1 out (draw reads)", cols, stride, n);
    return 0;
}
