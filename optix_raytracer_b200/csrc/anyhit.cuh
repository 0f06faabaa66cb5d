// anyhit.cuh — the reference's any-hit programs, run inside traversal (trav_coop.cuh) for triangles whose geometry
// flags leave any-hit enabled (MASK / BLEND materials, SDK/sutil/Scene.cpp:904-966).
//
//   AH_TEXTURE_MASK  __anyhit__texture_mask       SDK/optixRaycasting/optixRaycasting.cu:89-102   (1 ray type)
//   AH_WHITTED       __anyhit__radiance/occlusion SDK/cuda/whitted.cu:100-137                      (2 ray types; the occlusion
//                    program multiplies the ray's pending attenuation by 1 - alpha and lets the ray pass, whitted_cuda.h:127-159)
//
// The programs read whitted::HitGroupData from the hit-group record the hit selects (instance sbtOffset + GAS-local index x ray
// types + ray type), interpolate the texture coordinates as getLocalGeometry does (SDK/cuda/LocalGeometry.h:132-160) and fetch
// the base-colour texture with the caller's cudaTextureObject_t (sampleTexture, SDK/cuda/LocalShading.h:37-53).  Record offsets
// are pinned against the reference headers in tests/golden/kat.json ("hitgroup_layout", "whitted_layout").
//
// A program is a pure function of the hit (no state besides the returned factor), so the order in which traversal meets the
// candidates does not change which hits are accepted; the occlusion factors are multiplied up in double precision by the owning
// lane, so their order does not change the fp32 result either (a product of <= 8 fp32 factors is exact to 2^-50).
#pragma once
#include "accel.h"
#include "rt_math.cuh"

namespace b200rt {

constexpr uint32_t AH_NONE = 0, AH_TEXTURE_MASK = 1, AH_WHITTED = 2;

struct AnyHitCfg {
    const char* hg_base;   // hit-group records (header + whitted::HitGroupData)
    uint32_t hg_stride, hg_count;
    uint32_t mode;         // AH_*
};

struct AhBufView { uint64_t data; uint32_t count; uint16_t byte_stride; uint16_t elmt; };  // SDK/cuda/BufferView.h:32-38
struct AhTexture { int texcoord; int pad; cudaTextureObject_t tex; float2 offset, rotation, scale; };  // MaterialData::Texture (40 B)

// alpha of the base-colour texture at the hit: sampleTexture<float4>(material.pbr.base_color_tex, geom).w; 0 without a texture (T())
__device__ __forceinline__ float ah_base_alpha(const char* __restrict__ rec, uint32_t prim, float b1, float b2)
{
    const AhTexture t = *(const AhTexture*)(rec + 112 + 152);  // MaterialData @112, pbr.base_color_tex @152
    if (!t.tex) return 0.0f;
    const AhBufView vt = *(const AhBufView*)(rec + 64 + 16 * (t.texcoord & 1));
    float2 uv = make_float2(b1, b2);
    if (vt.data) {
        const AhBufView vi = *(const AhBufView*)(rec + 16);
        uint32_t i0, i1, i2;
        if (vi.elmt == 4) { const uint32_t* ip = (const uint32_t*)vi.data + 3 * (size_t)prim; i0 = ip[0]; i1 = ip[1]; i2 = ip[2]; }
        else if (vi.elmt == 2) { const uint16_t* ip = (const uint16_t*)vi.data + 3 * (size_t)prim; i0 = ip[0]; i1 = ip[1]; i2 = ip[2]; }
        else { i0 = 3 * prim; i1 = i0 + 1; i2 = i0 + 2; }
        const uint32_t st = vt.byte_stride ? vt.byte_stride : 8u;
        const float* u0 = (const float*)(vt.data + (uint64_t)i0 * st);
        const float* u1 = (const float*)(vt.data + (uint64_t)i1 * st);
        const float* u2 = (const float*)(vt.data + (uint64_t)i2 * st);
        const float b0 = (1.0f - b1) - b2;
        uv = make_float2(fm(b2, u2[0], fm(b1, u1[0], b0 * u0[0])), fm(b2, u2[1], fm(b1, u1[1], b0 * u0[1])));
    }
    const float ux = uv.x * t.scale.x, uy = uv.y * t.scale.y;
    const float tx = fm(uy, t.rotation.x, ux * t.rotation.y) + t.offset.x;
    const float ty = fm(uy, t.rotation.y, ux * -t.rotation.x) + t.offset.y;
    return tex2D<float4>(t.tex, tx, ty).w;
}

// Returns true when the hit stands, false for optixIgnoreIntersection().  `factor` multiplies the occlusion ray's pending
// attenuation (1 = unchanged).  `sbt_local` is the triangle's GAS-local SBT index, `inst_sbt` the instance's sbtOffset (0 for a GAS).
static __device__ __forceinline__ bool run_anyhit(AnyHitCfg c, uint32_t prim, uint32_t sbt_local, uint32_t inst_sbt, bool occlusion, float b1, float b2, float& factor)
{
    factor = 1.0f;
    const uint32_t ray_types = c.mode == AH_WHITTED ? 2u : 1u;
    uint32_t idx = inst_sbt + sbt_local * ray_types + ((c.mode == AH_WHITTED && occlusion) ? 1u : 0u);
    if (idx >= c.hg_count) idx = c.hg_count - 1;
    const char* rec = c.hg_base + (size_t)idx * c.hg_stride + 32;  // OPTIX_SBT_RECORD_HEADER_SIZE
    const int alpha_mode = *(const int*)(rec + 112 + 48);           // MaterialData::alpha_mode: 0 OPAQUE, 1 MASK, 2 BLEND
    const float alpha_cutoff = *(const float*)(rec + 112 + 52);
    if (c.mode == AH_TEXTURE_MASK) {
        if (alpha_mode == 1) return !(ah_base_alpha(rec, prim, b1, b2) < alpha_cutoff);
        return true;
    }
    // whitted: both programs do nothing without a base-colour texture
    if (*(const cudaTextureObject_t*)(rec + 112 + 152 + 8) == 0) return true;
    const float base_alpha = ah_base_alpha(rec, prim, b1, b2);
    if (!occlusion) return !(base_alpha < alpha_cutoff);  // "force mask mode, even for blend mode"
    if (alpha_mode == 0) return true;
    if (alpha_mode == 1 && base_alpha < alpha_cutoff) return false;
    // attenuation = payload * (1 - alpha); > 0 -> stored and the intersection ignored.  The payload is positive while pending, so the
    // sign of the product is the sign of 1 - alpha.
    const float f = 1.0f - base_alpha;
    if (f > 0.0f) { factor = f; return false; }
    return true;
}

}  // namespace b200rt
