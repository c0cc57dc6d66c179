// t3d_device.cuh -- device helpers shared by the kernels: the reference's scalar math restated with its
// evaluation order (PE.cpp / Matrix.cpp / MathUtil.h / pch.h of the reference, see t3d_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstring>

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

// The four corners of one billboard (SB.cpp:292-352 through PE.cpp:866-882), in world space.
// rotationPoint = (0.5, 0.5) (PE.cpp:855).  get_frozen() yields the latched rotation; it is only called for a rotated
// quad in T3D_ROT_FROZEN mode.
template <class GetFrozen>
__device__ __forceinline__ void quad_corners(const float pos[3], float size, float angle, const float right[3], const float up[3],
                                             int rot_mode, GetFrozen get_frozen, float p0[3], float p1[3], float p2[3], float p3[3])
{
    const float width = size, height = size;
    // SB.cpp:306-324
    float tfw[3];
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
}

// One billboard quad in SpriteBatch's vertex layout: 4 vertices x {position, uv, colour} = 36 floats written as nine
// float4.  SB.cpp:355-359: v0=(p0,u1,v1) v1=(p1,u2,v1) v2=(p2,u1,v2) v3=(p3,u2,v2), particle colour on all four.
template <class GetFrozen>
__device__ __forceinline__ void expand_quad(const float pos[3], const float color[4], float size, float angle, float u1,
                                            float v1, float u2, float v2, const float right[3], const float up[3],
                                            int rot_mode, GetFrozen get_frozen, float4* __restrict__ o)
{
    float p0[3], p1[3], p2[3], p3[3];
    quad_corners(pos, size, angle, right, up, rot_mode, get_frozen, p0, p1, p2, p3);
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

// The same quad with the vertex shader's projection already applied (res/shaders/sprite.vert:19:
// gl_Position = u_projectionMatrix * vec4(a_position, 1), the matrix being the node's view-projection matrix,
// PE.cpp:846-849): 4 vertices x {x, y, z, w, u, v, r, g, b, a} = 40 floats written as ten float4.  Each component is
// the left-to-right sum of MathUtil::transformVector4 (MathUtil.h:195-200) with w = 1.
template <class GetFrozen>
__device__ __forceinline__ void expand_quad_clip(const float pos[3], const float color[4], float size, float angle, float u1,
                                                 float v1, float u2, float v2, const float right[3], const float up[3],
                                                 int rot_mode, GetFrozen get_frozen, const float* __restrict__ vp,
                                                 float4* __restrict__ o)
{
    float p[4][3];
    quad_corners(pos, size, angle, right, up, rot_mode, get_frozen, p[0], p[1], p[2], p[3]);
    float c[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            c[q][r] = p[q][0] * vp[r] + p[q][1] * vp[4 + r] + p[q][2] * vp[8 + r] + 1.0f * vp[12 + r];
    }
    o[0] = make_float4(c[0][0], c[0][1], c[0][2], c[0][3]);
    o[1] = make_float4(u1, v1, color[0], color[1]);
    o[2] = make_float4(color[2], color[3], c[1][0], c[1][1]);
    o[3] = make_float4(c[1][2], c[1][3], u2, v1);
    o[4] = make_float4(color[0], color[1], color[2], color[3]);
    o[5] = make_float4(c[2][0], c[2][1], c[2][2], c[2][3]);
    o[6] = make_float4(u1, v2, color[0], color[1]);
    o[7] = make_float4(color[2], color[3], c[3][0], c[3][1]);
    o[8] = make_float4(c[3][2], c[3][3], u2, v2);
    o[9] = make_float4(color[0], color[1], color[2], color[3]);
}

// --------------------------------------------------------------------------------------------------
// spawn: one particle of ParticleEmitter::emitOnce (PE.cpp:312-364 with the generators of :611-679), from a source of
// raw draws.  Shared by k_spawn and by the update kernel's in-launch spawn, so both produce the same bits.
// --------------------------------------------------------------------------------------------------
struct SpawnedParticle
{
    float cs[4], ce[4];
    int energy;
    float size_start, size_end, spin, angle, rot_speed;
    float pos[3], vel[3], acc[3], axis[3];
    uint32_t frame;
};
// Src::next() yields the raw draws (integers in [0, 2^31-1]) in the order the reference consumes them.
template <class Src>
__device__ __forceinline__ SpawnedParticle spawn_particle(const SpawnParams& P, const float* __restrict__ world, Src& src)
{
    SpawnedParticle o;
    // PE.cpp:312-313 -> :669-679: x, y, z, w, one draw each.
#pragma unroll
    for (int k = 0; k < 4; ++k) o.cs[k] = P.color_start[k] + P.color_start_var[k] * u11(src.next());
#pragma unroll
    for (int k = 0; k < 4; ++k) o.ce[k] = P.color_end[k] + P.color_end_var[k] * u11(src.next());
    // PE.cpp:316: float overload (the bounds are float members), truncated into the integer field.
    o.energy = __float2int_rz(P.energy_min + (P.energy_max - P.energy_min) * u01(src.next()));
    o.size_start = P.size_start_min + (P.size_start_max - P.size_start_min) * u01(src.next()); // :317
    o.size_end = P.size_end_min + (P.size_end_max - P.size_end_min) * u01(src.next());         // :318
    o.spin = P.spin_min + (P.spin_max - P.spin_min) * u01(src.next());                          // :319-320
    o.angle = 0.0f + (o.spin - 0.0f) * u01(src.next());                                         // :321
    o.rot_speed = P.rot_speed_min + (P.rot_speed_max - P.rot_speed_min) * u01(src.next());      // :322
    if (P.ellipsoid)
    {
        // PE.cpp:629-650: rejection-sample the unit ball, scale, translate.
        float x, y, z;
        do
        {
            x = u11(src.next());
            y = u11(src.next());
            z = u11(src.next());
            // Vector3::length() > 1.0f (PE.cpp:640).  The correctly rounded square root of s exceeds 1 exactly when s exceeds
            // 1 + 2^-23 (sqrt(1 + 2^-23) = 1 + 2^-24 - ... rounds down to 1), so the comparison needs no square root
            // (checked against sqrtf over every float in [0.5, 2) by tests/test_abi_and_host.py).
        } while ((x * x + y * y + z * z) > 1.00000011920928955078125f);
        o.pos[0] = x * P.position_var[0];
        o.pos[1] = y * P.position_var[1];
        o.pos[2] = z * P.position_var[2];
        o.pos[0] += P.position[0];
        o.pos[1] += P.position[1];
        o.pos[2] += P.position[2];
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) o.pos[k] = P.position[k] + P.position_var[k] * u11(src.next()); // :617-626
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) o.vel[k] = P.velocity[k] + P.velocity_var[k] * u11(src.next());
#pragma unroll
    for (int k = 0; k < 3; ++k) o.acc[k] = P.acceleration[k] + P.acceleration_var[k] * u11(src.next());
#pragma unroll
    for (int k = 0; k < 3; ++k) o.axis[k] = P.rotation_axis[k] + P.rotation_axis_var[k] * u11(src.next());

    // PE.cpp:298-305, 332-351: orbiting properties are rotated by the node's world matrix with its translation removed.
    float w[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = world[k];
    const float tx = w[12], ty = w[13], tz = w[14];
    w[12] = 0.0f;
    w[13] = 0.0f;
    w[14] = 0.0f;
    if (P.orbit_position) transform4(w, o.pos[0], o.pos[1], o.pos[2], 1.0f, o.pos[0], o.pos[1], o.pos[2]);
    if (P.orbit_velocity) transform4(w, o.vel[0], o.vel[1], o.vel[2], 1.0f, o.vel[0], o.vel[1], o.vel[2]);
    if (P.orbit_acceleration) transform4(w, o.acc[0], o.acc[1], o.acc[2], 1.0f, o.acc[0], o.acc[1], o.acc[2]);
    if (o.rot_speed != 0.0f && !(o.axis[0] == 0.0f && o.axis[1] == 0.0f && o.axis[2] == 0.0f))
        transform4(w, o.axis[0], o.axis[1], o.axis[2], 1.0f, o.axis[0], o.axis[1], o.axis[2]);
    o.pos[0] += tx; // :354
    o.pos[1] += ty;
    o.pos[2] += tz;

    o.frame = 0; // :357-364
    if (P.frame_random_offset > 0) o.frame = (uint32_t)src.next() % P.frame_random_offset;
    return o;
}

// Cache hints of the streaming accesses.  Stores are evict-first (st.global.cs: every byte is written once per frame); loads are
// plain: evict-first loads (ld.global.cs) measured 0.7 % slower with nothing replaced per frame and 2-4 % slower under churn,
// where they push the movers' sectors out of L2 before their neighbours are asked for (profiles/r2_cache_hints.txt).  The
// -DT3D_EXP_* switches build the library with other choices (experiments, scripts/gpu/knobs.sh).
#if defined(T3D_EXP_ST_PLAIN)
#define T3D_STCS(p, v) (*(p) = (v))
#elif defined(T3D_EXP_ST_WT)
#define T3D_STCS(p, v) __stwt((p), (v))
#else
#define T3D_STCS(p, v) __stcs((p), (v))
#endif
#if defined(T3D_EXP_LD_QUAL) // a PTX load with the given qualifiers, e.g. -DT3D_EXP_LD_QUAL="ld.global.L2::256B"
__device__ __forceinline__ uint4 ldq_raw(const uint4* p)
{
    uint4 v;
    asm volatile(T3D_EXP_LD_QUAL ".v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(__cvta_generic_to_global(p)));
    return v;
}
__device__ __forceinline__ uint2 ldq_raw(const uint2* p)
{
    uint2 v;
    asm volatile(T3D_EXP_LD_QUAL ".v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(__cvta_generic_to_global(p)));
    return v;
}
__device__ __forceinline__ uint32_t ldq_raw(const uint32_t* p)
{
    uint32_t v;
    asm volatile(T3D_EXP_LD_QUAL ".u32 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(p)));
    return v;
}
__device__ __forceinline__ float4 ldq(const float4* p)
{
    const uint4 v = ldq_raw(reinterpret_cast<const uint4*>(p));
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
__device__ __forceinline__ int4 ldq(const int4* p)
{
    const uint4 v = ldq_raw(reinterpret_cast<const uint4*>(p));
    return make_int4((int)v.x, (int)v.y, (int)v.z, (int)v.w);
}
__device__ __forceinline__ float2 ldq(const float2* p)
{
    const uint2 v = ldq_raw(reinterpret_cast<const uint2*>(p));
    return make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
}
__device__ __forceinline__ int2 ldq(const int2* p)
{
    const uint2 v = ldq_raw(reinterpret_cast<const uint2*>(p));
    return make_int2((int)v.x, (int)v.y);
}
__device__ __forceinline__ uint32_t ldq(const uint32_t* p) { return ldq_raw(p); }
#define T3D_LDCS(p) ldq(p)
#elif defined(T3D_EXP_LD_CS)
#define T3D_LDCS(p) __ldcs(p)
#elif defined(T3D_EXP_LD_CG)
#define T3D_LDCS(p) __ldcg(p)
#else
#define T3D_LDCS(p) (*(p))
#endif
// V x 32-bit streaming accesses.
template <int V>
struct Vec;
template <>
struct Vec<4>
{
    static __device__ __forceinline__ void ld(const uint32_t* p, float* a)
    {
        const float4 v = T3D_LDCS(reinterpret_cast<const float4*>(p));
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    }
    static __device__ __forceinline__ void ld(const uint32_t* p, int* a)
    {
        const int4 v = T3D_LDCS(reinterpret_cast<const int4*>(p));
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    }
    static __device__ __forceinline__ void st(uint32_t* p, const float* a)
    {
        T3D_STCS(reinterpret_cast<float4*>(p), make_float4(a[0], a[1], a[2], a[3]));
    }
    static __device__ __forceinline__ void st(uint32_t* p, const int* a)
    {
        T3D_STCS(reinterpret_cast<int4*>(p), make_int4(a[0], a[1], a[2], a[3]));
    }
};
template <>
struct Vec<2>
{
    static __device__ __forceinline__ void ld(const uint32_t* p, float* a)
    {
        const float2 v = T3D_LDCS(reinterpret_cast<const float2*>(p));
        a[0] = v.x; a[1] = v.y;
    }
    static __device__ __forceinline__ void ld(const uint32_t* p, int* a)
    {
        const int2 v = T3D_LDCS(reinterpret_cast<const int2*>(p));
        a[0] = v.x; a[1] = v.y;
    }
    static __device__ __forceinline__ void st(uint32_t* p, const float* a) { T3D_STCS(reinterpret_cast<float2*>(p), make_float2(a[0], a[1])); }
    static __device__ __forceinline__ void st(uint32_t* p, const int* a) { T3D_STCS(reinterpret_cast<int2*>(p), make_int2(a[0], a[1])); }
};
__device__ __forceinline__ float ld_f(const uint32_t* p) { return __uint_as_float(T3D_LDCS(p)); }
__device__ __forceinline__ int ld_i(const uint32_t* p) { return (int)T3D_LDCS(p); }

// ---- asynchronous copies -------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
#if defined(T3D_EXP_CPA_QUAL) // e.g. -DT3D_EXP_CPA_QUAL="cp.async.cg.shared.global.L2::256B"
    asm volatile(T3D_EXP_CPA_QUAL " [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Life records in shared memory.  A warp owns 32 * V consecutive particles; lane t's V records form row t of the
// warp's stage (V * 64 bytes).  The warp fetches the 32 * V * 64 contiguous bytes cooperatively (chunk g = 16 bytes, lane j copies
// chunks j, j + 32, ...: 512 contiguous bytes per cp.async instruction) and every lane reads its own row back.  Rows are
// a multiple of the 128-byte bank period apart, so chunk p of a row is stored at position p ^ (row & 7): the eight lanes
// of a quarter-warp then hit eight different bank groups on both sides.
template <int V>
__device__ __forceinline__ uint32_t rec_chunk_offset(uint32_t row, uint32_t p)
{
    return row * (V * 64u) + ((p ^ (row & 7u)) << 4);
}
// issue: the warp's `n` records starting at `rec_g` (n <= 32 * V; records past n are not fetched)
template <int V>
__device__ __forceinline__ void rec_issue(unsigned char* stage_warp, const uint32_t* rec_g, uint32_t n, uint32_t lane)
{
#pragma unroll
    for (int i = 0; i < V * 4; ++i)
    {
        const uint32_t g = (uint32_t)i * 32u + lane; // chunk of the warp's run
        if ((g >> 2) < n) cp_async16(stage_warp + rec_chunk_offset<V>(g / (V * 4u), g % (V * 4u)), rec_g + g * 4u);
    }
}
// The same copy when some slots of the group receive a mover (PE.cpp:825-828): idx[l] = the particle whose record belongs
// in slot l of this lane's row -- the slot's own index, or the source of the move.  Every chunk of the stage is still
// written by exactly one copy, so the mover's record simply arrives in place of the dead particle's.
template <int V>
__device__ __forceinline__ void rec_issue_indexed(unsigned char* stage_warp, const uint32_t* rec, uint32_t n, uint32_t lane,
                                                  const uint32_t* idx)
{
#pragma unroll
    for (int i = 0; i < V * 4; ++i)
    {
        const uint32_t g = (uint32_t)i * 32u + lane; // chunk of the warp's run
        const uint32_t row = g / (V * 4u), p = g % (V * 4u);
        uint32_t src = 0;
#pragma unroll
        for (int l = 0; l < V; ++l)
        {
            const uint32_t t = __shfl_sync(0xffffffffu, idx[l], (int)row);
            if ((uint32_t)l == (p >> 2)) src = t;
        }
        if ((g >> 2) < n) cp_async16(stage_warp + rec_chunk_offset<V>(row, p), rec + (size_t)src * REC_WORDS + (p & 3u) * 4u);
    }
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
// read back: record l of this lane's row, as 16 words
template <int V>
__device__ __forceinline__ void rec_read(const unsigned char* stage_warp, uint32_t lane, int l, uint32_t* w)
{
#pragma unroll
    for (int c = 0; c < 4; ++c)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(stage_warp + rec_chunk_offset<V>(lane, (uint32_t)(l * 4 + c)));
        w[c * 4 + 0] = v.x; w[c * 4 + 1] = v.y; w[c * 4 + 2] = v.z; w[c * 4 + 3] = v.w;
    }
}

// Order-preserving map of a float onto an unsigned integer (for atomic min / max of bounding boxes).
__device__ __host__ __forceinline__ uint32_t float_order(float f)
{
#if defined(__CUDA_ARCH__)
    const uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __host__ __forceinline__ float float_unorder(uint32_t u)
{
    const uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// Which emitter of the launch owns global tile `tile`: the last item with tile_begin <= tile.  Every step is a
// dependent load, so the search starts from the interpolated position (exact when emitters have equal tile
// counts: two loads) and gallops outwards before bisecting.
// the interpolated position: exact when the emitters of the launch have equal tile counts
__device__ __forceinline__ uint32_t guess_item(uint32_t n_items, uint32_t tile, uint32_t total_tiles)
{
    if (n_items == 1) return 0;
    const uint32_t g = (uint32_t)(((unsigned long long)tile * n_items) / (total_tiles ? total_tiles : 1u));
    return g >= n_items ? n_items - 1 : g;
}
template <class Item>
__device__ __forceinline__ uint32_t find_item(const Item* __restrict__ items, uint32_t n_items, uint32_t tile,
                                              uint32_t total_tiles)
{
    if (n_items == 1) return 0;
    uint32_t lo, hi;
    const uint32_t g = guess_item(n_items, tile, total_tiles);
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
