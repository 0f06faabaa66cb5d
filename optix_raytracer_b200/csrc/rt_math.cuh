// rt_math.cuh — device-side arithmetic contract (DESIGN.md "arithmetic contract").
//
// Every floating-point operation on the hot path is an explicitly named IEEE-754 binary32 operation:
// products and sums are plain (the library is compiled with -fmad=false, so nvcc never fuses on its
// own), fused multiply-adds are written __fmaf_rn, division and square root are the correctly
// rounded __fdiv_rn / __fsqrt_rn.  The scalar oracle (oracle/oracle.cpp, compiled
// -ffp-contract=off) performs the same sequence, which is what makes hit records and accumulated
// radiance bit-comparable between the two.  Vector semantics follow the reference's
// SDK/sutil/vec_math.h:500-570 (normalize = v * (1/sqrt(dot)), faceforward, lerp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200rt {

__device__ __forceinline__ float fm(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 neg(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return fm(a.z, b.z, fm(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float3 cross(float3 a, float3 b)
{
    return f3(fm(a.y, b.z, -(a.z * b.y)), fm(a.z, b.x, -(a.x * b.z)), fm(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ float length(float3 v) { return fsqrt(dot(v, v)); }
__device__ __forceinline__ float3 normalize(float3 v) { return v * fdiv(1.0f, fsqrt(dot(v, v))); }
__device__ __forceinline__ float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }
__device__ __forceinline__ float clampf(float x, float a, float b) { return fmaxf(a, fminf(x, b)); }
__device__ __forceinline__ float sel3(float3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// ---- RNG: reference SDK/cuda/random.h:30-67 (integer arithmetic, bit exact) -------------------
__device__ __forceinline__ uint32_t tea4(uint32_t val0, uint32_t val1)
{
    uint32_t a = val0, b = val1, sum = 0;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
        sum += 0x9e3779b9u;
        a += ((b << 4) + 0xa341316cu) ^ (b + sum) ^ ((b >> 5) + 0xc8013ea4u);
        b += ((a << 4) + 0xad90777du) ^ (a + sum) ^ ((a >> 5) + 0x7e95761eu);
    }
    return a;
}
__device__ __forceinline__ uint32_t lcg(uint32_t& state)
{
    state = 1664525u * state + 1013904223u;
    return state & 0x00FFFFFFu;
}
// (float)lcg / 2^24 — both the conversion and the division by a power of two are exact
__device__ __forceinline__ float rnd(uint32_t& state) { return (float)lcg(state) * (1.0f / 16777216.0f); }

// ---- deterministic sin/cos on [0, 2*pi]: Cody–Waite by pi/2 + minimax polynomials, fma only ----
__device__ __forceinline__ void det_sincos(float phi, float& s, float& c)
{
    const float P1 = 1.5703125f, P2 = 4.837512969970703125e-4f, P3 = 7.54978995489188e-8f;
    const int k = (int)fm(phi, 0.636619772f, 0.5f);
    const float fk = (float)k;
    float r = fm(-fk, P1, phi);
    r = fm(-fk, P2, r);
    r = fm(-fk, P3, r);
    const float z = r * r;
    float sp = fm(z, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fm(sp, z, -1.6666654611e-1f);
    const float sr = fm(sp * z, r, r);
    float cp = fm(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fm(cp, z, 4.166664568298827e-2f);
    const float cr = fm(cp * z, z, fm(-0.5f, z, 1.0f));
    const int q = k & 3;
    s = (q == 0) ? sr : (q == 1) ? cr : (q == 2) ? -sr : -cr;
    c = (q == 0) ? cr : (q == 1) ? -sr : (q == 2) ? -cr : sr;
}

// ---- sRGB quantisation: reference SDK/cuda/helpers.h:36-64.  The reference build uses fast-math
// powf; __powf is the same approximation.  Compared against the oracle with a 1-LSB tolerance. ----
__device__ __forceinline__ float to_srgb1(float c)
{
    const float powed = __powf(c, 1.0f / 2.4f);
    return c < 0.0031308f ? 12.92f * c : fm(1.055f, powed, -0.055f);
}
__device__ __forceinline__ unsigned char quant8(float x)
{
    x = clampf(x, 0.0f, 1.0f);
    const unsigned int q = (unsigned int)(x * 256.0f);
    return (unsigned char)(q < 255u ? q : 255u);
}
__device__ __forceinline__ uchar4 make_color(float3 c)
{
    return make_uchar4(quant8(to_srgb1(clampf(c.x, 0.f, 1.f))), quant8(to_srgb1(clampf(c.y, 0.f, 1.f))),
                       quant8(to_srgb1(clampf(c.z, 0.f, 1.f))), 255u);
}

// ---- affine 3x4 helpers (row major) ------------------------------------------------------------
__device__ __forceinline__ float3 xform_point(const float* m, float3 p)
{
    return f3(fm(m[2], p.z, fm(m[1], p.y, m[0] * p.x)) + m[3], fm(m[6], p.z, fm(m[5], p.y, m[4] * p.x)) + m[7],
              fm(m[10], p.z, fm(m[9], p.y, m[8] * p.x)) + m[11]);
}
__device__ __forceinline__ float3 xform_vec(const float* m, float3 v)
{
    return f3(fm(m[2], v.z, fm(m[1], v.y, m[0] * v.x)), fm(m[6], v.z, fm(m[5], v.y, m[4] * v.x)),
              fm(m[10], v.z, fm(m[9], v.y, m[8] * v.x)));
}
// normal object->world = transpose(inverse) * n   (optixTransformNormalFromObjectToWorldSpace)
__device__ __forceinline__ float3 xform_normal(const float* inv, float3 n)
{
    return f3(fm(inv[8], n.z, fm(inv[4], n.y, inv[0] * n.x)), fm(inv[9], n.z, fm(inv[5], n.y, inv[1] * n.x)),
              fm(inv[10], n.z, fm(inv[6], n.y, inv[2] * n.x)));
}
// inverse of an affine 3x4: adjugate / det, then -Ainv*t (same op order as the oracle)
__host__ __device__ inline void invert34(const float* m, float* inv)
{
#ifdef __CUDA_ARCH__
#define B2_FM(a, b, c) __fmaf_rn(a, b, c)
#define B2_DIV(a, b) __fdiv_rn(a, b)
#else
#define B2_FM(a, b, c) __builtin_fmaf(a, b, c)
#define B2_DIV(a, b) ((a) / (b))
#endif
    const float a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const float c00 = B2_FM(e, i, -(f * h)), c01 = B2_FM(f, g, -(d * i)), c02 = B2_FM(d, h, -(e * g));
    const float det = B2_FM(c, c02, B2_FM(b, c01, a * c00));
    const float r = B2_DIV(1.0f, det);
    inv[0] = c00 * r; inv[1] = B2_FM(c, h, -(b * i)) * r; inv[2] = B2_FM(b, f, -(c * e)) * r;
    inv[4] = c01 * r; inv[5] = B2_FM(a, i, -(c * g)) * r; inv[6] = B2_FM(c, d, -(a * f)) * r;
    inv[8] = c02 * r; inv[9] = B2_FM(b, g, -(a * h)) * r; inv[10] = B2_FM(a, e, -(b * d)) * r;
    const float tx = m[3], ty = m[7], tz = m[11];
    inv[3] = -B2_FM(inv[2], tz, B2_FM(inv[1], ty, inv[0] * tx));
    inv[7] = -B2_FM(inv[6], tz, B2_FM(inv[5], ty, inv[4] * tx));
    inv[11] = -B2_FM(inv[10], tz, B2_FM(inv[9], ty, inv[8] * tx));
#undef B2_FM
#undef B2_DIV
}

}  // namespace b200rt
