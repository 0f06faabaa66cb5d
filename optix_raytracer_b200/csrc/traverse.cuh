// traverse.cuh — software traversal of the 8-wide compressed BVH (accel.h) on sm_100a.
//
// Replaces the closed optixTrace / optixTraverse of the reference (device call sites:
// SDK/optixPathTracer/optixPathTracer.cu:184,227; SDK/optixRaycasting/optixRaycasting.cu:53;
// SDK/optixMultiGPU/optixMultiGPU.cu:158,183).  B200 has no RT cores, so this *is* the hot loop.
//
// Semantics (the contract the oracle restates):
//   * a triangle is hit when the watertight test (fixed IEEE op order, original object-space
//     vertices) yields tmin < t < tmax;
//   * closest hit = minimum (t, instance index, triangle ordinal) — strictly smaller t wins, equal t
//     goes to the lower ordinal — so the result does not depend on BVH topology or visit order;
//   * any-hit (TERMINATE_ON_FIRST_HIT) returns whether any triangle is hit;
//   * box tests are conservative: child boxes are padded at build time and the slab comparison
//     carries a relative slack, so a triangle that the exact test would accept is never culled.
// One thread per ray, traversal stack of (node group | triangle group) words as in the CWBVH
// paper; octant-ordered child visiting via the per-node slot assignment.
#pragma once
#include "accel.h"
#include "rt_math.cuh"

namespace b200rt {

// TRAV_STACK (entries of the per-lane traversal stack): accel.h — the builder checks the tree depth against it
constexpr float BOX_SLACK = 1.0000038f;   // 1 + 2^-18 on the far side of every slab comparison
constexpr float DIR_EPS = 8.27180613e-25f;  // 2^-80: |d| below this is clamped for the *box* tests only

struct RayHit {
    float t;         // in: tmax / out: hit distance
    float b1, b2;    // OptiX barycentrics (weights of vertices 1 and 2)
    uint32_t prim;   // primitive index (build-input local + primitiveIndexOffset)
    uint32_t sbt;    // GAS-local SBT index | flags << 24 (+ instance sbtOffset added by the caller)
    uint32_t ord;    // GAS-global triangle ordinal
    uint32_t inst;   // instance index (0 when a GAS is traced directly)
};

struct TravStats { uint32_t nodes, tris; };

__device__ __forceinline__ uint32_t ldg_u32(const void* p) { return __ldg((const uint32_t*)p); }

// per-ray constants of the watertight test
struct TriRay {
    float3 o;
    int kx, ky, kz;
    float Sx, Sy, Sz;
};
__device__ __forceinline__ TriRay make_tri_ray(float3 o, float3 d)
{
    TriRay r;
    r.o = o;
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0;
    float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    const float dz = sel3(d, kz);
    if (dz < 0.0f) { const int tmp = kx; kx = ky; ky = tmp; }
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = fdiv(sel3(d, kx), dz);
    r.Sy = fdiv(sel3(d, ky), dz);
    r.Sz = fdiv(1.0f, dz);
    return r;
}

// cull_word(ray_flags, instance_flags): accel.h.
// does the triangle (geometry flags gflags) run any-hit programs under this cull word?
__device__ __forceinline__ bool anyhit_off(uint32_t gflags, uint32_t cull) { return (cull & 3u) ? (cull & 1u) != 0u : (gflags & 1u) != 0u; }

// Watertight test; identical op sequence to oracle.cpp:tri_hit.  Accepts tmin < t and
// (t < best.t  or  t == best.t with lower ordinal once something was found).
template <bool ANY>
__device__ __forceinline__ bool tri_test(const TriRay& r, const float4 q0, const float4 q1, const float4 q2, float tmin,
                                         RayHit& best, bool& found, uint32_t cull)
{
    const float3 A = xyz(q0) - r.o, B = xyz(q1) - r.o, C = xyz(q2) - r.o;
    const float Akz = sel3(A, r.kz), Bkz = sel3(B, r.kz), Ckz = sel3(C, r.kz);
    const float Ax = fm(-r.Sx, Akz, sel3(A, r.kx)), Ay = fm(-r.Sy, Akz, sel3(A, r.ky));
    const float Bx = fm(-r.Sx, Bkz, sel3(B, r.kx)), By = fm(-r.Sy, Bkz, sel3(B, r.ky));
    const float Cx = fm(-r.Sx, Ckz, sel3(C, r.kx)), Cy = fm(-r.Sy, Ckz, sel3(C, r.ky));
    // edge functions: two rounded products and one subtraction (NOT fused) — this is what makes the value for a
    // shared edge exactly antisymmetric between the two triangles, i.e. watertight (Woop et al. 2013, sec. 3); the library is built with -fmad=false
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
        V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
        W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = (U + V) + W;
    if (det == 0.0f) return false;
    const float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
    const float T = fm(W, Cz, fm(V, Bz, U * Az));
    const float t = fdiv(T, det);
    if (!(t > tmin && t <= best.t)) return false;
    if (cull) {
        const uint32_t gflags = __float_as_uint(q1.w) >> TRI_FLAG_SHIFT;
        if (!(gflags & 4u)) {  // OPTIX_GEOMETRY_FLAG_DISABLE_TRIANGLE_FACE_CULLING
            if ((cull & 16u) && det < 0.0f) return false;
            if ((cull & 32u) && det > 0.0f) return false;
        }
        // OPTIX_RAY_FLAG_CULL_DISABLED_ANYHIT (1<<6) / CULL_ENFORCED_ANYHIT (1<<7) against the effective any-hit state of the triangle
        if (cull & 0xc0u) {
            const bool off = anyhit_off(gflags, cull);
            if ((cull & 64u) && off) return false;
            if ((cull & 128u) && !off) return false;
        }
    }
    const uint32_t ord = __float_as_uint(q2.w);
    if (t == best.t && !(found && ord < best.ord)) return false;
    best.t = t;
    if (!ANY) {
        best.b1 = fdiv(V, det);
        best.b2 = fdiv(W, det);
        best.prim = __float_as_uint(q0.w);
        best.sbt = __float_as_uint(q1.w);
        best.ord = ord;
    }
    found = true;
    return true;
}

// bytes with bit 4 set -> 0xff, others -> 0x00 (x has at most bit 4 of each byte set)
__device__ __forceinline__ uint32_t byte_mask_from_bit4(uint32_t x) { return (x >> 4) * 0xffu; }
// Bytes of w as floats, exactly and without the quarter-rate I2F (XU pipe) that (float)((w >> 8j) & 0xff) compiles to — ncu
// showed XU as the busiest pipe of the node test (profiles/r01_trace_v1.md: xu 42 %, fma 16 %), then, after a PRMT + FADD
// version, the ALU pipe (64 %).  One PRMT (ALU pipe) builds two fp16 bit patterns 0x64qq = 1024 + q at once, and sm_100's
// mixed-precision add (PTX add.rn.f32.f16 -> SASS FHADD with a .H0/.H1 operand select, FMA pipe) subtracts the 1024 in fp32:
// 24 ALU + 48 FMA-pipe instructions per node visit for its 48 plane bytes, every value exact.
struct BytePair { uint32_t h; };  // half2 bit pattern {1024 + b_lo, 1024 + b_hi}
__device__ __forceinline__ BytePair byte_pair01(uint32_t w) { return BytePair{__byte_perm(w, 0x64646464u, 0x4140u)}; }
__device__ __forceinline__ BytePair byte_pair23(uint32_t w) { return BytePair{__byte_perm(w, 0x64646464u, 0x4342u)}; }
#ifndef B200RT_DECODE_FADD
#define B200RT_DECODE_FADD 0
#endif
__device__ __forceinline__ float pair_lo(BytePair p)
{
#if B200RT_DECODE_FADD
    return __uint_as_float(0x4B000000u | (p.h & 0xffu)) - 8388608.0f;
#else
    float f;
    asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(f) : "h"((unsigned short)(p.h & 0xffffu)), "f"(-1024.0f));
    return f;
#endif
}
__device__ __forceinline__ float pair_hi(BytePair p)
{
#if B200RT_DECODE_FADD
    return __uint_as_float(0x4B000000u | ((p.h >> 16) & 0xffu)) - 8388608.0f;
#else
    float f;
    asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(f) : "h"((unsigned short)(p.h >> 16)), "f"(-1024.0f));
    return f;
#endif
}
// single byte (compile-time j) through the same path; used by the one-ray-per-thread reference traversal below
__device__ __forceinline__ float byte_f(uint32_t w, int j)
{
    const BytePair p = (j < 2) ? byte_pair01(w) : byte_pair23(w);
    return (j & 1) ? pair_hi(p) : pair_lo(p);
}

// Trace one ray through one GAS.  `best.t` carries tmax in and the closest t out.
template <bool ANY, bool STATS>
__device__ __forceinline__ bool trace_gas(const AccelHeader* __restrict__ gas, float3 o, float3 d, float tmin, RayHit& best,
                                          uint32_t cull, TravStats* st)
{
    const char* base = (const char*)gas;
    const uint4* __restrict__ nodes = (const uint4*)(base + gas->nodes_off);
    const float4* __restrict__ tris = (const float4*)(base + gas->tris_off);
    if (gas->num_tris == 0) return false;

    const TriRay tr = make_tri_ray(o, d);
    // box-test direction: clamp tiny components (keeps the sign) so 1/d stays finite
    const float bx = fabsf(d.x) < DIR_EPS ? copysignf(DIR_EPS, d.x) : d.x;
    const float by = fabsf(d.y) < DIR_EPS ? copysignf(DIR_EPS, d.y) : d.y;
    const float bz = fabsf(d.z) < DIR_EPS ? copysignf(DIR_EPS, d.z) : d.z;
    const float idx = fdiv(1.0f, bx), idy = fdiv(1.0f, by), idz = fdiv(1.0f, bz);
    const uint32_t oct = (bx < 0.0f ? 4u : 0u) | (by < 0.0f ? 2u : 0u) | (bz < 0.0f ? 1u : 0u);
    const uint32_t octinv4 = (7u - oct) * 0x01010101u;

    uint2 stack[TRAV_STACK];
    int sp = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u);  // root: node index base 0, one pending bit at priority 7
    uint2 tgroup = make_uint2(0u, 0u);
    bool found = false;

    for (;;) {
        if (ngroup.y & 0xff000000u) {
            // pop the highest-priority pending child of the current node group
            const uint32_t hits_imask = ngroup.y;
            const uint32_t bit = 31u - __clz(hits_imask);
            const uint32_t child_base = ngroup.x;
            ngroup.y &= ~(1u << bit);
            if (ngroup.y & 0xff000000u) {
                if (sp < TRAV_STACK) stack[sp++] = ngroup;
            }
            const uint32_t slot = (bit - 24u) ^ (octinv4 & 0xffu);
            const uint32_t rel = __popc(hits_imask & ~(0xffffffffu << slot));
            const uint32_t node_index = child_base + rel;
            const uint4* np = nodes + (size_t)node_index * 5u;
            const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (STATS) st->nodes++;

            const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
            const uint32_t e_imask = n0.w;
            const float sx = __uint_as_float((e_imask & 0xffu) << 23);
            const float sy = __uint_as_float(((e_imask >> 8) & 0xffu) << 23);
            const float sz = __uint_as_float(((e_imask >> 16) & 0xffu) << 23);
            const float aix = sx * idx, aiy = sy * idy, aiz = sz * idz;
            const float aox = (px - o.x) * idx, aoy = (py - o.y) * idy, aoz = (pz - o.z) * idz;
            const float tfar = best.t;
#ifdef B200RT_DEBUG_TRAVERSAL
            printf("node %u: P %g %g %g s %g %g %g id %g %g %g ai %g %g %g ao %g %g %g tfar %g oct %u\n", node_index, px, py, pz, sx, sy, sz, idx, idy, idz, aix, aiy, aiz, aox, aoy, aoz, tfar, oct);
#endif

            uint32_t hitmask = 0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t meta4 = half ? n1.w : n1.z;
                const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                const uint32_t inner_mask4 = byte_mask_from_bit4(is_inner4);
                const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
                const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
                const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
                const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
                const uint32_t xn = bx < 0.0f ? qhix : qlox, xf = bx < 0.0f ? qlox : qhix;
                const uint32_t yn = by < 0.0f ? qhiy : qloy, yf = by < 0.0f ? qloy : qhiy;
                const uint32_t zn = bz < 0.0f ? qhiz : qloz, zf = bz < 0.0f ? qloz : qhiz;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float tnx = fm(byte_f(xn, j), aix, aox), tfx = fm(byte_f(xf, j), aix, aox);
                    const float tny = fm(byte_f(yn, j), aiy, aoy), tfy = fm(byte_f(yf, j), aiy, aoy);
                    const float tnz = fm(byte_f(zn, j), aiz, aoz), tfz = fm(byte_f(zf, j), aiz, aoz);
                    const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
                    const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tfar));
                    if (cmin <= cmax * BOX_SLACK) {
                        const uint32_t cb = (child_bits4 >> (8 * j)) & 0xffu;
                        const uint32_t bi = (bit_index4 >> (8 * j)) & 0xffu;
                        hitmask |= cb << bi;
                    }
                }
            }
#ifdef B200RT_DEBUG_TRAVERSAL
            printf("   hitmask %08x\n", hitmask);
#endif
            ngroup = make_uint2(n1.x, (hitmask & 0xff000000u) | (e_imask >> 24));
            tgroup = make_uint2(n1.y, hitmask & 0x00ffffffu);
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }

        while (tgroup.y) {
            const uint32_t ti = 31u - __clz(tgroup.y);
            tgroup.y &= ~(1u << ti);
            const float4* tp = tris + (size_t)(tgroup.x + ti) * 3u;
            const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
            if (STATS) st->tris++;
            if (tri_test<ANY>(tr, q0, q1, q2, tmin, best, found, cull)) {
                if (ANY) return true;
            }
        }

        if ((ngroup.y & 0xff000000u) == 0u) {
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
    return found;
}

// Trace through a traversable handle: a GAS directly, or an IAS (flat instance list, ray carried to
// object space with the stored inverse; `t` is shared between spaces as in OptiX).
template <bool ANY, bool STATS>
__device__ __forceinline__ bool trace_handle(const AccelHeader* __restrict__ h, float3 o, float3 d, float tmin, float tmax,
                                             uint32_t ray_flags, RayHit& hit, TravStats* st)
{
    hit.t = tmax;
    hit.inst = 0;
    const uint32_t cull = cull_word(ray_flags, 0u);
    if (h->kind == ACCEL_KIND_GAS) {
        // initial acceptance must be strict (t < tmax): run with found=false semantics
        return trace_gas<ANY, STATS>(h, o, d, tmin, hit, cull, st);
    }
    bool any = false;
    const InstanceRecord* recs = (const InstanceRecord*)((const char*)h + h->inst_off);
    const uint32_t n = h->num_instances;
    for (uint32_t k = 0; k < n; ++k) {
        const InstanceRecord* ir = recs + k;
        if (!(ir->mask & ray_visibility(ray_flags))) continue;
        const float3 oo = xform_point(ir->inv, o), dd = xform_vec(ir->inv, d);
        RayHit cand = hit;
        const uint32_t c = cull_word(ray_flags, ir->flags);
        if (trace_gas<ANY, STATS>((const AccelHeader*)ir->gas, oo, dd, tmin, cand, c, st)) {
            // a later instance only replaces on strictly smaller t (lower instance index wins ties)
            if (!any || cand.t < hit.t) {
                hit = cand;
                hit.inst = k;
                any = true;
                if (ANY) return true;
            }
        }
    }
    return any;
}

}  // namespace b200rt
