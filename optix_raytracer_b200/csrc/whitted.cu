// whitted.cu — wavefront restatement of the glTF viewer's shading path (optixMeshViewer, BASELINE.json configs[2]).
//
// Replaces   optixLaunch(scene.pipeline(), 0, d_params, sizeof(whitted::LaunchParams), scene.sbt(), width, height, 1)
// (SDK/optixMeshViewer/optixMeshViewer.cpp:283-308) and the device programs of SDK/cuda/whitted.cu:
//   __raygen__pinhole (44-98)            pixel-centre ray at subframe 0, tea<4>/rnd jitter afterwards; running-mean accumulation
//   __miss__constant_radiance (139-142)  params.miss_color
//   __closesthit__radiance (149-289)     glTF PBR direct lighting: base colour / metallic-roughness / emissive / normal textures
//                                        (sampleTexture, SDK/cuda/LocalShading.h:37-53), GGX D * Smith Vis * Schlick F per point light
//                                        (whitted_cuda.h:48-80) behind one shadow ray each (tmin 0.001, tmax L_dist - 0.001)
//   getLocalGeometry (SDK/cuda/LocalGeometry.h:59-176)
// It consumes the reference's 128-byte whitted::LaunchParams, its Light records (36 B) and its SBT whose hit-group records carry
// whitted::HitGroupData {GeometryData 112 B, MaterialData 240 B} with real cudaTextureObject_t handles, two records per primitive
// group (radiance, occlusion: SDK/sutil/Scene.cpp:1405-1433) — layouts pinned in tests/golden/kat.json ("whitted_layout").
// Texture fetches are hardware tex2D<float4> on the caller's texture objects, exactly what the reference's programs execute.
//
// Scope: OPAQUE materials.  MASK / BLEND materials need the any-hit programs (whitted.cu:100-137, alpha cut-outs and the
// pending/committed occlusion attenuation) inside traversal — SURVEY.md 8(f) rank 1; a launch that meets one fails with
// B200RT_ERROR_NOT_SUPPORTED instead of rendering it wrongly.
//
// Stages (one sample per pixel and launch, as the reference):
//   RAYGEN  -> TRACE closest (CULL_BACK_FACING_TRIANGLES, whitted_cuda.h:110) -> SHADE (material, per-light BRDF terms, shadow rays)
//   -> TRACE any -> RESOLVE (sum the unoccluded terms in light order, accumulate, make_color)
#include <string.h>

#include <algorithm>

#include "accel.h"
#include "internal.h"
#include "rt_math.cuh"

namespace b200rt {

struct WBufView { uint64_t data; uint32_t count; uint16_t byte_stride; uint16_t elmt; };  // SDK/cuda/BufferView.h:32-38
struct WTexture { int texcoord; int pad; cudaTextureObject_t tex; float2 offset, rotation, scale; };  // MaterialData::Texture (40 B)
static_assert(sizeof(WTexture) == 40, "MaterialData::Texture");
struct WMaterial {  // MaterialData (240 B), PBR view of the union
    int type; int pad0;
    WTexture normal_tex;
    int alpha_mode; float alpha_cutoff;
    float emissive_factor[3]; int pad1;
    WTexture emissive_tex;
    unsigned char double_sided; unsigned char pad2[15];
    float base_color[4];
    float metallic, roughness;
    WTexture base_color_tex;
    WTexture metallic_roughness_tex;
    char tail[8];
};
static_assert(sizeof(WMaterial) == 240 && offsetof(WMaterial, alpha_mode) == 48 && offsetof(WMaterial, emissive_factor) == 56 &&
                  offsetof(WMaterial, emissive_tex) == 72 && offsetof(WMaterial, double_sided) == 112 && offsetof(WMaterial, base_color) == 128 &&
                  offsetof(WMaterial, metallic) == 144 && offsetof(WMaterial, base_color_tex) == 152 && offsetof(WMaterial, metallic_roughness_tex) == 192,
              "MaterialData layout");
struct WLight { int type; float color[3]; float intensity; float position[3]; int falloff; };  // Light (36 B), Point view
static_assert(sizeof(WLight) == 36, "Light layout");
struct WParams {  // whitted::LaunchParams (128 B)
    unsigned int width, height, subframe_index;
    float4* accum_buffer;
    uchar4* frame_buffer;
    int max_depth; float scene_epsilon;
    float3 eye, U, V, W;
    WBufView lights;
    float3 miss_color;
    uint64_t handle;
};
static_assert(sizeof(WParams) == 128 && offsetof(WParams, accum_buffer) == 16 && offsetof(WParams, eye) == 40 && offsetof(WParams, lights) == 88 &&
                  offsetof(WParams, miss_color) == 104 && offsetof(WParams, handle) == 120,
              "whitted::LaunchParams layout");

constexpr uint32_t W_RAY_TYPES = 2;  // whitted::RAY_TYPE_COUNT
constexpr uint32_t W_MAX_TRACE_DEPTH = 8;
struct WCounters { unsigned int nhit; unsigned int pad[3]; };
// asynchronous launch errors, written by the kernels into the context's pinned host block and reported by the NEXT launch
// (like CUDA's own asynchronous errors): [0] a MASK / BLEND material was hit, [1] more lights than the workspace was sized for
struct WAsyncFlags { unsigned int unsupported_material; unsigned int too_many_lights; };

// ---- RAYGEN (whitted.cu:44-80) -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) w_raygen_kernel(const WParams* __restrict__ params, uint32_t width, uint32_t height, float4* __restrict__ rays,
                                                        WCounters* __restrict__ counters)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) counters->nhit = 0;
    if (i >= width * height) return;
    const WParams P = *params;
    const uint32_t ix = i % width, iy = i / width;
    uint32_t seed = tea4(iy * width + ix, P.subframe_index);
    float jx = 0.5f, jy = 0.5f;
    if (P.subframe_index != 0) { jx = rnd(seed); jy = rnd(seed); }
    const float dx = fm(2.0f, fdiv((float)ix + jx, (float)width), -1.0f), dy = fm(2.0f, fdiv((float)iy + jy, (float)height), -1.0f);
    const float3 dir = normalize(f3(fm(dy, P.V.x, dx * P.U.x) + P.W.x, fm(dy, P.V.y, dx * P.U.y) + P.W.y, fm(dy, P.V.z, dx * P.U.z) + P.W.z));
    rays[2 * (size_t)i] = make_float4(P.eye.x, P.eye.y, P.eye.z, 0.0f);
    rays[2 * (size_t)i + 1] = make_float4(dir.x, dir.y, dir.z, 1e16f);
}

__device__ __forceinline__ float3 ldf3(uint64_t base, uint32_t idx, uint32_t stride)
{
    const float* p = (const float*)(base + (uint64_t)idx * (stride ? stride : 12u));
    return f3(p[0], p[1], p[2]);
}
__device__ __forceinline__ float2 ldf2(uint64_t base, uint32_t idx, uint32_t stride)
{
    const float* p = (const float*)(base + (uint64_t)idx * (stride ? stride : 8u));
    return make_float2(p[0], p[1]);
}
__device__ __forceinline__ float4 ldf4(uint64_t base, uint32_t idx, uint32_t stride)
{
    const float* p = (const float*)(base + (uint64_t)idx * (stride ? stride : 16u));
    return make_float4(p[0], p[1], p[2], p[3]);
}
__device__ __forceinline__ float3 bary3(float b0, float b1, float b2, float3 a, float3 b, float3 c)
{
    return f3(fm(b2, c.x, fm(b1, b.x, b0 * a.x)), fm(b2, c.y, fm(b1, b.y, b0 * a.y)), fm(b2, c.z, fm(b1, b.z, b0 * a.z)));
}

struct WGeom { float3 P, N, Ng; float2 UV[2]; float3 dpdu[2], dpdv[2]; float4 color; };

// getLocalGeometry (LocalGeometry.h:59-163) for a triangle mesh
__device__ __forceinline__ WGeom w_local_geometry(const char* __restrict__ rec, uint32_t prim, float b1, float b2, const InstanceRecord* ir)
{
    const WBufView vi = *(const WBufView*)(rec + 16), vp = *(const WBufView*)(rec + 32), vn = *(const WBufView*)(rec + 48),
                   vt0 = *(const WBufView*)(rec + 64), vt1 = *(const WBufView*)(rec + 80), vc = *(const WBufView*)(rec + 96);
    uint32_t i0, i1, i2;
    if (vi.elmt == 4) { const uint32_t* ip = (const uint32_t*)vi.data + 3 * (size_t)prim; i0 = ip[0]; i1 = ip[1]; i2 = ip[2]; }
    else if (vi.elmt == 2) { const uint16_t* ip = (const uint16_t*)vi.data + 3 * (size_t)prim; i0 = ip[0]; i1 = ip[1]; i2 = ip[2]; }
    else { i0 = 3 * prim; i1 = i0 + 1; i2 = i0 + 2; }
    const float b0 = (1.0f - b1) - b2;
    WGeom g;
    const float3 P0 = ldf3(vp.data, i0, vp.byte_stride), P1 = ldf3(vp.data, i1, vp.byte_stride), P2 = ldf3(vp.data, i2, vp.byte_stride);
    g.P = bary3(b0, b1, b2, P0, P1, P2);
    if (ir) g.P = xform_point(ir->m, g.P);
    g.color = make_float4(1.f, 1.f, 1.f, 1.f);
    if (vc.data) {
        const float4 c0 = ldf4(vc.data, i0, vc.byte_stride), c1 = ldf4(vc.data, i1, vc.byte_stride), c2 = ldf4(vc.data, i2, vc.byte_stride);
        g.color = make_float4(fm(b2, c2.x, fm(b1, c1.x, b0 * c0.x)), fm(b2, c2.y, fm(b1, c1.y, b0 * c0.y)), fm(b2, c2.z, fm(b1, c1.z, b0 * c0.z)),
                              fm(b2, c2.w, fm(b1, c1.w, b0 * c0.w)));
    }
    float3 Ng = cross(P1 - P0, P2 - P0);
    if (ir) Ng = xform_normal(ir->inv, Ng);
    g.Ng = normalize(Ng);
    float3 N0, N1, N2;
    if (vn.data) {
        N0 = ldf3(vn.data, i0, vn.byte_stride); N1 = ldf3(vn.data, i1, vn.byte_stride); N2 = ldf3(vn.data, i2, vn.byte_stride);
        float3 N = bary3(b0, b1, b2, N0, N1, N2);
        if (ir) N = xform_normal(ir->inv, N);
        g.N = normalize(N);
    } else {
        g.N = N0 = N1 = N2 = g.Ng;
    }
    const float3 dp1 = P0 - P2, dp2 = P1 - P2;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const WBufView vt = j ? vt1 : vt0;
        if (vt.data) {
            const float2 u0 = ldf2(vt.data, i0, vt.byte_stride), u1 = ldf2(vt.data, i1, vt.byte_stride), u2 = ldf2(vt.data, i2, vt.byte_stride);
            g.UV[j] = make_float2(fm(b2, u2.x, fm(b1, u1.x, b0 * u0.x)), fm(b2, u2.y, fm(b1, u1.y, b0 * u0.y)));
            const float du1 = u0.x - u2.x, du2 = u1.x - u2.x, dv1 = u0.y - u2.y, dv2 = u1.y - u2.y;
            const float det = fm(du1, dv2, -(dv1 * du2));
            const float invdet = fdiv(1.0f, det);
            g.dpdu[j] = f3(fm(dv2, dp1.x, -(dv1 * dp2.x)) * invdet, fm(dv2, dp1.y, -(dv1 * dp2.y)) * invdet, fm(dv2, dp1.z, -(dv1 * dp2.z)) * invdet);
            g.dpdv[j] = f3(fm(du1, dp2.x, -(du2 * dp1.x)) * invdet, fm(du1, dp2.y, -(du2 * dp1.y)) * invdet, fm(du1, dp2.z, -(du2 * dp1.z)) * invdet);
        } else {
            g.UV[j] = make_float2(b1, b2);
            g.dpdu[j] = neg(dp1);
            g.dpdv[j] = f3(dp2.x - dp1.x, dp2.y - dp1.y, dp2.z - dp1.z);
        }
    }
    return g;
}

// sampleTexture<float4> (LocalShading.h:37-53): hardware bilinear fetch on the caller's texture object
__device__ __forceinline__ float4 w_sample(const WTexture& t, const WGeom& g)
{
    const float2 uv = g.UV[t.texcoord & 1];
    const float ux = uv.x * t.scale.x, uy = uv.y * t.scale.y;
    const float tx = fm(uy, t.rotation.x, ux * t.rotation.y) + t.offset.x;
    const float ty = fm(uy, t.rotation.y, ux * -t.rotation.x) + t.offset.y;
    return tex2D<float4>(t.tex, tx, ty);
}

// ---- SHADE: __miss__constant_radiance and __closesthit__radiance up to the shadow rays ------------------------------------------------
__global__ void __launch_bounds__(128) w_shade_kernel(const WParams* __restrict__ params, uint32_t width, uint32_t height, const float4* __restrict__ rays,
                                                       const ExtHit* __restrict__ hits, const uint32_t* __restrict__ hit_sbt,
                                                       const char* __restrict__ hg_base, uint32_t hg_stride, uint32_t hg_count, float4* __restrict__ base,
                                                       int* __restrict__ slot, float4* __restrict__ probes, float4* __restrict__ terms,
                                                       WCounters* __restrict__ counters, uint32_t nl_cap, WAsyncFlags* __restrict__ async_flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t npix = width * height;
    const WParams P = *params;
    if (i == 0 && P.lights.count > nl_cap) async_flags->too_many_lights = P.lights.count;
    const uint32_t nl = min(P.lights.count, nl_cap);
    bool is_hit = false;
    ExtHit h;
    if (i < npix) {
        h = hits[i];
        is_hit = h.t >= 0.0f;
        if (!is_hit) { base[i] = make_float4(P.miss_color.x, P.miss_color.y, P.miss_color.z, 0.f); slot[i] = -1; }
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, is_hit);
    uint32_t kb = 0;
    const uint32_t lane = threadIdx.x & 31u;
    if (mask && lane == (uint32_t)(__ffs(mask) - 1)) kb = atomicAdd(&counters->nhit, (unsigned)__popc(mask));
    kb = __shfl_sync(0xffffffffu, kb, mask ? __ffs(mask) - 1 : 0);
    if (!is_hit) return;
    const uint32_t k = kb + __popc(mask & ((1u << lane) - 1u));
    slot[i] = (int)k;
    const float4 ro = rays[2 * (size_t)i], rd = rays[2 * (size_t)i + 1];
    const float3 dir = f3(rd.x, rd.y, rd.z);
    const AccelHeader* handle = (const AccelHeader*)P.handle;
    const InstanceRecord* ir = nullptr;
    // SBT index = instance.sbtOffset + GAS-local index * RAY_TYPE_COUNT + RAY_TYPE_RADIANCE (Scene.cpp:1147-1154: sbtOffset advances by
    // primitive groups * ray types)
    uint32_t rec_idx = hit_sbt[i] * W_RAY_TYPES;
    if (handle->kind == ACCEL_KIND_IAS) {
        ir = (const InstanceRecord*)((const char*)handle + handle->inst_off) + h.inst;
        rec_idx += ir->sbt_offset;
    }
    if (rec_idx >= hg_count) rec_idx = hg_count - 1;
    const char* rec = hg_base + (size_t)rec_idx * hg_stride + B200RT_SBT_RECORD_HEADER_SIZE;
    const WGeom g = w_local_geometry(rec, h.prim, h.b1, h.b2, ir);
    const WMaterial& m = *(const WMaterial*)(rec + 112);
    if (m.alpha_mode != 0) async_flags->unsupported_material = 1;  // MASK / BLEND need the any-hit programs (not built yet): flagged, rendered opaque

    // material (whitted.cu:157-186)
    float4 bc = make_float4(m.base_color[0] * g.color.x, m.base_color[1] * g.color.y, m.base_color[2] * g.color.z, m.base_color[3] * g.color.w);
    if (m.base_color_tex.tex) {
        const float4 t = w_sample(m.base_color_tex, g);
        bc = make_float4(bc.x * __powf(t.x, 2.2f), bc.y * __powf(t.y, 2.2f), bc.z * __powf(t.z, 2.2f), bc.w * t.w);
    }
    float metallic = m.metallic, roughness = m.roughness;
    if (m.metallic_roughness_tex.tex) {
        const float4 t = w_sample(m.metallic_roughness_tex, g);
        roughness *= t.y;
        metallic *= t.z;
    }
    const float F0 = 0.04f;
    const float km = 1.0f - metallic;
    const float3 diff_color = f3((bc.x * (1.0f - F0)) * km, (bc.y * (1.0f - F0)) * km, (bc.z * (1.0f - F0)) * km);
    // lerp(F0, base_color, metallic) = a + t * (b - a)
    const float3 spec_color = f3(fm(metallic, bc.x - F0, F0), fm(metallic, bc.y - F0, F0), fm(metallic, bc.z - F0, F0));
    const float alpha = roughness * roughness;
    float3 result = f3(0.f, 0.f, 0.f);
    float4 et = make_float4(1.f, 1.f, 1.f, 1.f);
    if (m.emissive_tex.tex) et = w_sample(m.emissive_tex, g);
    result = f3(fm(m.emissive_factor[0], et.x, result.x), fm(m.emissive_factor[1], et.y, result.y), fm(m.emissive_factor[2], et.z, result.z));
    float3 N = g.N;
    if (m.normal_tex.tex) {
        const int tc = m.normal_tex.texcoord & 1;
        const float4 t = w_sample(m.normal_tex, g);
        const float nx = fm(2.0f, t.x, -1.0f), ny = fm(2.0f, t.y, -1.0f), nz = fm(2.0f, t.z, -1.0f);
        const float2 rot = m.normal_tex.rotation;
        const float tx = fm(ny, -rot.x, nx * rot.y), ty = fm(ny, rot.y, nx * rot.x);
        const float3 du = normalize(g.dpdu[tc]), dv = normalize(g.dpdv[tc]);
        N = normalize(f3(fm(nz, g.N.x, fm(ty, dv.x, tx * du.x)), fm(nz, g.N.y, fm(ty, dv.y, tx * du.y)), fm(nz, g.N.z, fm(ty, dv.z, tx * du.z))));
    }
    if (dot(N, dir) > 0.0f) N = neg(N);
    // lights (whitted.cu:222-262): depth = payload depth (0) + 1 < MAX_TRACE_DEPTH always holds for the primary hit
    const float3 V = neg(normalize(dir));
    for (uint32_t li = 0; li < nl; ++li) {
        const WLight L = *(const WLight*)(P.lights.data + (uint64_t)li * (P.lights.byte_stride ? P.lights.byte_stride : 36u));
        float4 term = make_float4(0.f, 0.f, 0.f, 0.f);                    // w: 0 = nothing, 1 = add if the probe is unoccluded, 2 = add
        float4 po = make_float4(g.P.x, g.P.y, g.P.z, 1.0f), pd = make_float4(0.f, 0.f, 1.f, -1.0f);  // tmin > tmax: a null probe
        if (L.type == 0) {
            const float3 Lv = f3(L.position[0] - g.P.x, L.position[1] - g.P.y, L.position[2] - g.P.z);
            const float L_dist = length(Lv);
            const float3 Ld = f3(fdiv(Lv.x, L_dist), fdiv(Lv.y, L_dist), fdiv(Lv.z, L_dist));
            const float3 H = normalize(Ld + V);
            const float N_dot_L = dot(N, Ld), N_dot_V = dot(N, V), N_dot_H = dot(N, H), V_dot_H = dot(V, H);
            if (N_dot_L > 0.0f && N_dot_V > 0.0f) {
                // schlick / vis / ggxNormal (whitted_cuda.h:48-72); pow(x, 5) as exact products
                const float x1 = 1.0f - V_dot_H, x2 = x1 * x1, x5 = (x2 * x2) * x1;
                const float3 F = f3(fm(1.0f - spec_color.x, x5, spec_color.x), fm(1.0f - spec_color.y, x5, spec_color.y), fm(1.0f - spec_color.z, x5, spec_color.z));
                const float a2 = alpha * alpha;
                const float ggx0 = N_dot_L * fsqrt(fm(N_dot_V * N_dot_V, 1.0f - a2, a2));
                const float ggx1 = N_dot_V * fsqrt(fm(N_dot_L * N_dot_L, 1.0f - a2, a2));
                const float G_vis = fdiv((2.0f * N_dot_L) * N_dot_V, ggx0 + ggx1);
                const float xx = fm(N_dot_H * N_dot_H, a2 - 1.0f, 1.0f);
                const float D = fdiv(a2, (3.14159265358979323846f * xx) * xx);
                const float3 diff = f3(fdiv((1.0f - F.x) * diff_color.x, 3.14159265358979323846f), fdiv((1.0f - F.y) * diff_color.y, 3.14159265358979323846f),
                                       fdiv((1.0f - F.z) * diff_color.z, 3.14159265358979323846f));
                const float3 spec = f3((F.x * G_vis) * D, (F.y * G_vis) * D, (F.z * G_vis) * D);
                // light.color * attenuation(=1) * intensity * N_dot_L * (diff + spec)
                const float sx = (L.color[0] * L.intensity) * N_dot_L, sy = (L.color[1] * L.intensity) * N_dot_L, sz = (L.color[2] * L.intensity) * N_dot_L;
                term = make_float4(sx * (diff.x + spec.x), sy * (diff.y + spec.y), sz * (diff.z + spec.z), 1.0f);
                po.w = 0.001f;
                pd = make_float4(Ld.x, Ld.y, Ld.z, L_dist - 0.001f);
            }
        } else if (L.type == 1) {
            term = make_float4(L.color[0] * bc.x, L.color[1] * bc.y, L.color[2] * bc.z, 2.0f);
        }
        const size_t s = (size_t)k * nl + li;
        terms[s] = term;
        probes[2 * s] = po;
        probes[2 * s + 1] = pd;
    }
    base[i] = make_float4(result.x, result.y, result.z, 0.f);
    (void)ro;
}

// ---- RESOLVE: rest of __closesthit__radiance + tail of __raygen__pinhole (whitted.cu:84-97) -----------------------------------------
__global__ void __launch_bounds__(256) w_resolve_kernel(const WParams* __restrict__ params, uint32_t width, uint32_t height, const float4* __restrict__ base,
                                                         const int* __restrict__ slot, const float4* __restrict__ terms,
                                                         const uint32_t* __restrict__ occluded, uint32_t nl_cap)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width * height) return;
    const WParams P = *params;
    const uint32_t nl = min(P.lights.count, nl_cap);
    const float4 b = base[i];
    float3 result = f3(b.x, b.y, b.z);
    const int k = slot[i];
    if (k >= 0) {
        for (uint32_t li = 0; li < nl; ++li) {
            const size_t s = (size_t)k * nl + li;
            const float4 t = terms[s];
            if (t.w == 2.0f || (t.w == 1.0f && occluded[s] == 0u)) result = f3(result.x + t.x, result.y + t.y, result.z + t.z);
        }
    }
    if (P.subframe_index > 0) {
        const float a = fdiv(1.0f, (float)(P.subframe_index + 1u));
        const float4 prev = P.accum_buffer[i];
        result = f3(fm(a, result.x - prev.x, prev.x), fm(a, result.y - prev.y, prev.y), fm(a, result.z - prev.z, prev.z));
    }
    P.accum_buffer[i] = make_float4(result.x, result.y, result.z, 1.0f);
    if (P.frame_buffer) P.frame_buffer[i] = make_color(result);
}

// ---- host --------------------------------------------------------------------------------------------------------------------------
int launch_whitted(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr d_params, const b200rt_shader_binding_table* sbt, unsigned width, unsigned height)
{
    B2_REQUIRE(ctx, d_params && sbt, "null argument");
    B2_REQUIRE(ctx, sbt->hitgroupRecordBase && sbt->hitgroupRecordCount > 0 && sbt->hitgroupRecordStrideInBytes >= 32 + 352,
               "hit-group records (whitted::HitGroupData, 2 per primitive group) are required");
    const uint64_t npix64 = (uint64_t)width * height;
    B2_REQUIRE(ctx, npix64 < (1ull << 28), "launch too large");
    if (npix64 == 0) return 0;
    DeviceGuard guard(ctx->device);
    // errors of earlier launches on this context surface here (no synchronisation on the steady-state path)
    WAsyncFlags* flags = (WAsyncFlags*)((char*)ctx->pinned + 1024);
    if (flags->unsupported_material) {
        flags->unsupported_material = 0;
        return set_error(ctx, B200RT_ERROR_NOT_SUPPORTED,
                         "an earlier whitted launch hit a MASK / BLEND material: the any-hit programs of whitted.cu are not built yet (it was rendered as opaque)");
    }
    if (flags->too_many_lights) {
        ctx->w_params = 0;  // re-read the light count below
        flags->too_many_lights = 0;
        return set_error(ctx, B200RT_ERROR_INVALID_OPERATION, "an earlier whitted launch found more lights than its workspace was sized for (frame incomplete); relaunch");
    }
    // The workspace is sized by the light count.  It is read back (one stream synchronisation) on the first launch with a given
    // d_params; later launches with the same d_params reuse it and the kernels check the live value on the device.
    if (ctx->w_params != d_params) {
        WParams hp;
        B2_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, (const void*)d_params, sizeof(WParams), cudaMemcpyDeviceToHost, s));
        B2_CUDA(ctx, cudaStreamSynchronize(s));
        memcpy(&hp, ctx->pinned, sizeof(WParams));
        B2_REQUIRE(ctx, hp.handle && hp.accum_buffer, "LaunchParams has null pointers");
        B2_REQUIRE(ctx, hp.lights.count <= 64 && (hp.lights.count == 0 || hp.lights.data), "bad light list");
        ctx->w_params = d_params;
        ctx->w_lights = hp.lights.count;
    }
    const uint32_t npix = (uint32_t)npix64, nl = ctx->w_lights, nlp = std::max(nl, 1u);
    size_t off = 16384;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_cnt = take(sizeof(WCounters)), o_rays = take(32ull * npix), o_hits = take(sizeof(ExtHit) * (size_t)npix), o_sbt = take(4ull * npix),
                 o_base = take(16ull * npix), o_slot = take(4ull * npix), o_probe = take(32ull * npix * nlp), o_terms = take(16ull * npix * nlp),
                 o_occ = take(4ull * npix * nlp);
    int rc = ensure_workspace(ctx, off, s);
    if (rc) return rc;
    char* W = (char*)ctx->ws.ptr;
    WCounters* cnt = (WCounters*)(W + o_cnt);
    float4* rays = (float4*)(W + o_rays);
    ExtHit* hits = (ExtHit*)(W + o_hits);
    uint32_t* hsbt = (uint32_t*)(W + o_sbt);
    float4* base = (float4*)(W + o_base);
    int* slot = (int*)(W + o_slot);
    float4* probes = (float4*)(W + o_probe);
    float4* terms = (float4*)(W + o_terms);
    uint32_t* occ = (uint32_t*)(W + o_occ);
    const WParams* dp = (const WParams*)d_params;
    const b200rt_deviceptr handle_dev = d_params + offsetof(WParams, handle);  // the kernels read the live handle
    w_raygen_kernel<<<div_up(npix, 256), 256, 0, s>>>(dp, width, height, rays, cnt);
    B2_LAUNCH_CHECK(ctx);
    // radiance rays cull back faces (whitted_cuda.h:110); DISABLE_TRIANGLE_FACE_CULLING geometry (doubleSided) is exempt in the triangle test
    rc = trace_buffer(ctx, s, 0, (b200rt_deviceptr)rays, npix, nullptr, 1, 0, B200RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES, (b200rt_deviceptr)hits, 0,
                      (b200rt_deviceptr)hsbt, handle_dev);
    if (rc) return rc;
    w_shade_kernel<<<div_up(npix, 128), 128, 0, s>>>(dp, width, height, rays, hits, hsbt, (const char*)sbt->hitgroupRecordBase,
                                                     sbt->hitgroupRecordStrideInBytes, sbt->hitgroupRecordCount, base, slot, probes, terms, cnt, nl, flags);
    B2_LAUNCH_CHECK(ctx);
    if (nl) {
        rc = trace_buffer(ctx, s, 0, (b200rt_deviceptr)probes, (uint64_t)npix * nl, &cnt->nhit, nl, 1, 0u, (b200rt_deviceptr)occ, 0, 0, handle_dev);
        if (rc) return rc;
    }
    w_resolve_kernel<<<div_up(npix, 256), 256, 0, s>>>(dp, width, height, base, slot, terms, occ, nl);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

// sutil::Scene::addImage + addSampler (SDK/sutil/Scene.cpp:576-652): 8-bit RGBA image -> CUDA array -> texture object with
// normalised coordinates, normalised-float reads, the given address modes and filter.  Returns the cudaTextureObject_t.
int texture_create(b200rt_context ctx, int width, int height, const void* rgba8, int address_s, int address_t, int linear, uint64_t* tex_out,
                   uint64_t* array_out)
{
    B2_REQUIRE(ctx, width > 0 && height > 0 && rgba8 && tex_out && array_out, "bad argument");
    DeviceGuard guard(ctx->device);
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
    cudaArray_t arr = nullptr;
    B2_CUDA(ctx, cudaMallocArray(&arr, &cd, width, height));
    B2_CUDA(ctx, cudaMemcpy2DToArray(arr, 0, 0, rgba8, (size_t)width * 4, (size_t)width * 4, height, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = (cudaTextureAddressMode)address_s;
    td.addressMode[1] = (cudaTextureAddressMode)address_t;
    td.filterMode = linear ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    td.maxAnisotropy = 1;
    td.maxMipmapLevelClamp = 99;
    td.minMipmapLevelClamp = 0;
    td.mipmapFilterMode = cudaFilterModePoint;
    td.borderColor[0] = 1.0f;
    td.sRGB = 0;
    cudaTextureObject_t tex = 0;
    B2_CUDA(ctx, cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    *tex_out = (uint64_t)tex;
    *array_out = (uint64_t)(uintptr_t)arr;
    return 0;
}

int texture_destroy(b200rt_context ctx, uint64_t tex, uint64_t array)
{
    DeviceGuard guard(ctx->device);
    if (tex) cudaDestroyTextureObject((cudaTextureObject_t)tex);
    if (array) cudaFreeArray((cudaArray_t)(uintptr_t)array);
    return 0;
}

}  // namespace b200rt
