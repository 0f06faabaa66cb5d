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
//   __anyhit__radiance / __anyhit__occlusion (100-137)   alpha cut-outs on radiance rays, pending attenuation on occlusion rays
//                                        (anyhit.cuh, run inside traversal)
//   ALPHA_MODE_BLEND continuation (266-286)  result = result * alpha + trace(same ray, tmin = t, depth + 1) * (1 - alpha), depth < 8
//
// Four launches per subframe for a scene without BLEND materials (one sample per pixel and launch, as the reference):
//   RAYGEN   one thread per pixel: the camera ray against the padded scene bounds; a ray that passes them by is a miss and its pixel is
//            written here (miss program + raygen tail); the rest become work items
//   PRIMARY  persistent traversal (trav_coop.cuh) of those camera rays: miss -> pixel written in the commit, hit -> a 48-byte hit slot
//   SHADE    one thread per hit slot: getLocalGeometry, the material, the per-light BRDF factors and shadow probes
//   SHADOW   persistent traversal, one item per (hit slot, light): the probe (TERMINATE_ON_FIRST_HIT, attenuation through the any-hit
//            program); the item that completes a slot sums the slot's terms in light order and writes the pixel
// BLEND materials add levels: SHADE puts the continuation of a BLEND hit on a list, the host reads the count (the only
// synchronisation, and only for scenes that have such materials) and runs PRIMARY / SHADE / SHADOW for the list; COMBINE then folds
// each pixel's chain of levels back to front exactly as the recursion of the reference returns.
// (One persistent launch for the whole program — pixel = work item — was built and measured slower: profiles/r01_whitted_launches.md.)
#include <string.h>

#include <algorithm>

#include "accel.h"
#include "internal.h"
#include "rt_math.cuh"
#include "trav_coop.cuh"
#include "loop_graph.h"
#include "anyhit.cuh"

#ifndef B200RT_WHITTED_COHERENT_AXES
#define B200RT_WHITTED_COHERENT_AXES 0   // axis-specialised triangle test in the cooperative rounds too: measured slower (opaque 0.195 vs 0.191 ms, MASK 0.293 vs 0.266: spills)
#endif

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

constexpr uint32_t W_RAY_TYPES = 2;  // whitted::RAY_TYPE_COUNT
constexpr uint32_t W_MAX_TRACE_DEPTH = 8;
#ifndef B200RT_W_MIN_CTAS
#define B200RT_W_MIN_CTAS 5
#endif
constexpr int W_MIN_CTAS = B200RT_W_MIN_CTAS;  // resident CTAs per SM the two traversal kernels are compiled for (5 -> 96 registers, no spills)
constexpr uint32_t W_BLEND_SLOT_FACTOR = 4;  // hit-slot capacity per pixel when the scene has BLEND materials (chains of up to 8 levels)

// one hit of a radiance ray (level 0: the camera ray; level l: the l-th BLEND continuation of a pixel)
struct WSlot {
    uint32_t pixel;
    int parent;        // slot whose continuation this is (-1 at level 0)
    float t;
    uint32_t prim, inst;
    float b1, b2;
    uint32_t sbt;      // GAS-local SBT index of the triangle
    int next;          // set by the next level: slot of the continuation's hit, W_NEXT_MISS, or W_NEXT_NONE (no continuation was traced)
    uint32_t flags;    // WS_*
    float one_minus_alpha;
    uint32_t level;
};
static_assert(sizeof(WSlot) == 48, "WSlot");
constexpr int W_NEXT_NONE = -1, W_NEXT_MISS = -2;
constexpr uint32_t WS_BLEND = 1u;   // ALPHA_MODE_BLEND: the level's result is scaled by alpha and the continuation added
constexpr uint32_t WS_CONT = 2u;    // a continuation ray was put on the list

struct WCounters {
    unsigned int nslots;   // hit slots allocated so far (all levels)
    unsigned int ncont;    // continuation rays SHADE put on the list for the next level
    unsigned int nchain;   // level-0 BLEND slots (pixels COMBINE has to fold)
    unsigned int overflow; // hits dropped because the slot capacity was reached
    unsigned int nprimary; // camera rays RAYGEN found heading for the scene bounds (PRIMARY's work items at level 0)
    // BLEND scenes: the levels run as a device-side loop (loop_graph.h); its state
    unsigned int level;        // level being traced
    unsigned int level_start;  // first hit slot of the level
    unsigned int n_items;      // continuation rays of the level (level > 0)
};
// asynchronous launch errors, written by the kernels into the context's pinned host block and reported by the NEXT launch
// (like CUDA's own asynchronous errors)
struct WAsyncFlags { unsigned int unexpected_blend; unsigned int too_many_lights; unsigned int slot_overflow; };

// everything the three stages share, passed by value (constant bank)
struct WK;
__device__ __forceinline__ uint32_t w_level(const WK& k);
__device__ __forceinline__ uint32_t w_level_start(const WK& k);
__device__ __forceinline__ unsigned int* w_fetch(const WK& k, uint32_t which);
struct WK {
    const WParams* params;
    uint32_t width, height, nl_cap, level, level_start, cap_slots;
    const char* hg_base;
    uint32_t hg_stride, hg_count;
    uint32_t raygen_traverses;  // 1: RAYGEN traverses the camera rays of an opaque scene itself (launches without BLEND levels)
    uint32_t blend_levels;  // 1: the BLEND levels are run (the host found such a material; level / level_start / fetch come from the
                            //    device-side loop state in WCounters); 0: one level, and SHADE flags a BLEND material as an error
    WCounters* counters;
    WSlot* slots;
    float4* base;        // per slot: emission part of the result
    float4* t0;          // per (slot, light): light colour xyz, intensity
    float4* t1;          // per (slot, light): (diff + spec) xyz, N.L
    uint32_t* kinds;     // per (slot, light): 0 nothing, 1 point light behind a shadow probe, 2 ambient term in t0.xyz
    float4* probes;      // per (slot, light): origin | tmin, direction | tmax
    float4* result;      // per slot: the level's radiance (already scaled by alpha for BLEND)
    float* att;          // per (slot, light): committed attenuation of the probe
    unsigned int* arrived;  // per slot: probe items finished (zeroed by SHADE)
    uint32_t* cont;      // parent slots of the next level's continuation rays
    uint32_t* chain;     // level-0 BLEND slots
    uint32_t* primary;   // pixels whose camera ray heads for the scene bounds (written by RAYGEN)
    unsigned int* fetch; // work-item cursor of the persistent launch
    WAsyncFlags* async_flags;
};
__device__ __forceinline__ uint32_t w_level(const WK& k) { return k.blend_levels ? k.counters->level : 0u; }
__device__ __forceinline__ uint32_t w_level_start(const WK& k) { return k.blend_levels ? k.counters->level_start : 0u; }
// work-item cursor of a persistent launch: two per level (PRIMARY, SHADOW), all zeroed at the start of the frame
__device__ __forceinline__ unsigned int* w_fetch(const WK& k, uint32_t which) { return k.fetch + 2u * w_level(k) + which; }


// __raygen__pinhole's ray for a pixel (whitted.cu:44-80); a continuation regenerates it (optixGetWorldRayOrigin / Direction)
__device__ __forceinline__ void w_camera_ray(const WParams& P, uint32_t width, uint32_t height, uint32_t pixel, float3& org, float3& dir)
{
    const uint32_t ix = pixel % width, iy = pixel / width;
    uint32_t seed = tea4(iy * width + ix, P.subframe_index);
    float jx = 0.5f, jy = 0.5f;
    if (P.subframe_index != 0) { jx = rnd(seed); jy = rnd(seed); }
    const float dx = fm(2.0f, fdiv((float)ix + jx, (float)width), -1.0f), dy = fm(2.0f, fdiv((float)iy + jy, (float)height), -1.0f);
    dir = normalize(f3(fm(dy, P.V.x, dx * P.U.x) + P.W.x, fm(dy, P.V.y, dx * P.U.y) + P.W.y, fm(dy, P.V.z, dx * P.U.z) + P.W.z));
    org = P.eye;
}

// tail of __raygen__pinhole (whitted.cu:84-97)
__device__ __forceinline__ void w_write_pixel(const WParams& P, uint32_t pixel, float3 result, float4 prev)   // prev = accum_buffer[pixel], read by the caller
{
    if (P.subframe_index > 0) {
        const float a = fdiv(1.0f, (float)(P.subframe_index + 1u));
        result = f3(fm(a, result.x - prev.x, prev.x), fm(a, result.y - prev.y, prev.y), fm(a, result.z - prev.z, prev.z));
    }
    P.accum_buffer[pixel] = make_float4(result.x, result.y, result.z, 1.0f);
    if (P.frame_buffer) P.frame_buffer[pixel] = make_color(result);
}
__device__ __forceinline__ void w_write_pixel(const WParams& P, uint32_t pixel, float3 result)
{
    float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
    if (P.subframe_index > 0) prev = P.accum_buffer[pixel];
    w_write_pixel(P, pixel, result, prev);
}

__device__ __forceinline__ uint32_t w_inst_sbt(const AccelHeader* handle, uint32_t inst)
{
    return handle->kind == ACCEL_KIND_IAS ? ((const InstanceRecord*)((const char*)handle + handle->inst_off) + inst)->sbt_offset : 0u;
}

// ---- PRIMARY: radiance rays of one level ---------------------------------------------------------------------------------------------
template <bool AH>
struct WPrimaryWork {
    static constexpr bool CONTINUES = false;
    static constexpr bool ANYHIT = AH;
    static constexpr bool COHERENT_AXES = B200RT_WHITTED_COHERENT_AXES != 0;   // camera rays
    const WK& k;
    const AccelHeader* handle;
    uint32_t level;
    uint32_t pixel;
    int parent;
    __device__ WPrimaryWork(const WK& k_, const AccelHeader* h, uint32_t level_) : k(k_), handle(h), level(level_), pixel(0), parent(-1) {}

    __device__ __forceinline__ bool anyhit_enabled() const { return handle->anyhit != 0u; }
    __device__ __forceinline__ bool anyhit(uint32_t prim, uint32_t sbt, uint32_t inst, uint32_t pack, float b1, float b2, float& factor) const
    {
        return run_anyhit(AnyHitCfg{k.hg_base, k.hg_stride, k.hg_count, AH_WHITTED}, prim, sbt, w_inst_sbt(handle, inst), false, b1, b2, factor);
    }
    __device__ __forceinline__ void attenuate(float) {}
    __device__ __forceinline__ bool stream_triangles() const { return false; }
    __device__ __forceinline__ void ray(float3& o, float3& d, float& tmin) const
    {
        const WParams& P = *k.params;
        w_camera_ray(P, k.width, k.height, pixel, o, d);
        tmin = parent >= 0 ? k.slots[parent].t : 0.0f;  // continuation: tmin = optixGetRayTmax() of the BLEND hit (whitted.cu:279)
    }
    __device__ __forceinline__ bool fetch(uint32_t item, Trav& s, float* my_ray)
    {
        if (level == 0) { pixel = k.primary[item]; parent = -1; }
        else { parent = (int)k.cont[item]; pixel = k.slots[parent].pixel; }
        float3 o, d;
        float tmin;
        ray(o, d, tmin);
        s.best.t = 1e16f;
        // radiance rays cull back faces (whitted_cuda.h:110); DISABLE_TRIANGLE_FACE_CULLING geometry (doubleSided) is exempt in the triangle test
        // most camera rays of a model viewer pass the model by: they are dropped at the bounds of the instance (trav_begin<BOUNDS>)
        if (!trav_begin_handle<true>(s, my_ray, handle, o, d, tmin, 0u, B200RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES & 0xf0u, 0u)) { commit(s, false); return false; }
        return true;
    }
    __device__ __forceinline__ bool next_instance(Trav& s, float* my_ray)
    {
        if (handle->kind == ACCEL_KIND_GAS) return false;
        float3 o, d;
        float tmin;
        ray(o, d, tmin);
        return trav_begin_handle<true>(s, my_ray, handle, o, d, tmin, s.pack & TP_FOUND_ANY, B200RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES & 0xf0u, s.inst + 1u);
    }
    __device__ __forceinline__ void commit(const Trav& s, bool found)
    {
        // hits take a slot (one warp-aggregated atomic for the lanes committing together)
        const uint32_t mask = __ballot_sync(__activemask(), found);
        if (!found) {
            // __miss__constant_radiance (whitted.cu:139-142)
            if (parent < 0) { const WParams& P = *k.params; w_write_pixel(P, pixel, P.miss_color); }
            else k.slots[parent].next = W_NEXT_MISS;
            return;
        }
        const uint32_t lane = threadIdx.x & 31u, leader = __ffs(mask) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&k.counters->nslots, (unsigned)__popc(mask));
        base = __shfl_sync(mask, base, leader);
        const uint32_t slot = base + __popc(mask & ((1u << lane) - 1u));
        if (slot >= k.cap_slots) {
            // capacity reached (only possible on BLEND levels): the continuation is dropped like one beyond the depth cap
            k.counters->overflow = 1u;
            if (parent >= 0) k.slots[parent].next = W_NEXT_NONE;
            return;
        }
        WSlot w;
        w.pixel = pixel; w.parent = parent; w.t = s.best.t; w.prim = s.best.prim; w.inst = s.best.inst; w.b1 = s.best.b1; w.b2 = s.best.b2;
        w.sbt = s.best.sbt & TRI_SBT_MASK; w.next = W_NEXT_NONE; w.flags = 0u; w.one_minus_alpha = 0.f; w.level = level;
        k.slots[slot] = w;
        if (parent >= 0) k.slots[parent].next = (int)slot;
    }
};

// Both instantiations are enqueued; each reads AccelHeader::anyhit on the device and returns at once unless the traversable is its
// kind (raycast.cu: trace_rays_kernel): opaque scenes pay nothing for the any-hit machinery, and the choice cannot go stale when
// another handle is written into the same LaunchParams block.
template <bool AH>
__global__ void __launch_bounds__(COOP_BLOCK, W_MIN_CTAS) w_primary_kernel(const __grid_constant__ WK k, uint32_t n_items)
{
    chain_enter();
    const AccelHeader* handle = (const AccelHeader*)k.params->handle;
    if ((handle->anyhit != 0u) != AH) return;
    const uint32_t level = w_level(k);
    n_items = level == 0 ? k.counters->nprimary : k.counters->n_items;  // counted by RAYGEN / by the previous level's SHADE
    WPrimaryWork<AH> work(k, handle, level);
    trace_persistent(work, n_items, w_fetch(k, 0u), nullptr);
}

// ---- RAYGEN: one thread per pixel — the camera ray against the bounds of the scene -----------------------------------------------------
// A model viewer's camera rays mostly pass the model by.  Those pixels are finished right here, by 2 M independent threads (the miss
// program and the raygen tail: whitted.cu:84-97,139-142).  What happens to the rays that reach the (padded, hence conservative:
// trav_coop.cuh trav_begin<BOUNDS>) scene bounds depends on the scene:
//   * opaque scene (WK::raygen_traverses; B200RT_WHITTED_INLINE bit 0, default on): the thread traverses its ray on the spot
//     (trav_coop.cuh: trace_one_per_thread), and the probes run one per thread too (w_shadow_simple_kernel; bit 1, default on).  Camera
//     rays in 8 x 4 pixel tiles are coherent, the probes of neighbouring hit slots start next to each other and aim at the same light.
//     History on the Duck at 1080p: this form in rows of 32 pixels 0.257 ms against 0.248 for the persistent kernels; in tiles 0.227
//     (probes persistent) / 0.226; with the axis-specialised triangle test 0.191-0.202 / 0.184-0.189 (profiles/r02_session4_experiments.md);
//   * scene with any-hit geometry, or B200RT_WHITTED_INLINE=0: the pixel becomes a work item of the persistent traversal
//     (w_primary_kernel), whose lanes take one item at a time and would otherwise spend their time on the latency of fetch -> set-up
//     -> accum read for rays that hit nothing (measured: 150 us of a 270 us frame, profiles/r01_whitted_launches.md).
template <bool AH>
struct WPixelWork : WPrimaryWork<AH> {
    using WPrimaryWork<AH>::pixel; using WPrimaryWork<AH>::parent; using WPrimaryWork<AH>::handle; using WPrimaryWork<AH>::ray; using WPrimaryWork<AH>::commit;
    __device__ WPixelWork(const WK& k_, const AccelHeader* h) : WPrimaryWork<AH>(k_, h, 0u) {}
    __device__ __forceinline__ bool fetch(uint32_t item, Trav& s, float* my_ray)
    {
        pixel = item; parent = -1;
        float3 o, d;
        float tmin;
        ray(o, d, tmin);
        s.best.t = 1e16f;
        if (!trav_begin_handle<true>(s, my_ray, handle, o, d, tmin, 0u, B200RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES & 0xf0u, 0u)) { commit(s, false); return false; }
        return true;
    }
};

#ifndef B200RT_WHITTED_TILED
#define B200RT_WHITTED_TILED 1
#endif
#ifndef B200RT_WHITTED_SCREEN_RECT
#define B200RT_WHITTED_SCREEN_RECT 1
#endif
#ifndef B200RT_WHITTED_INLINE_ANYHIT
#define B200RT_WHITTED_INLINE_ANYHIT 1
#endif
#ifndef B200RT_WHITTED_INLINE_DEFAULT
#define B200RT_WHITTED_INLINE_DEFAULT 3   // Duck 1080p: 0 0.243, 1 0.227, 2 0.255, 3 0.226 ms with tiles alone; with the axis-specialised triangle test of the one-ray-per-thread driver 1: 0.191-0.202, 3: 0.184-0.189
#endif
// A warp takes an 8 x 4 tile of pixels (a CTA 16 x 8), so the candidates a warp appends to the PRIMARY work list are neighbours in both
// directions: the 32 camera rays a PRIMARY warp picks up together walk the same nodes more often than 32 pixels of one row do.
#ifdef B200RT_WRAYGEN_MIN_CTAS
#define W_RAYGEN_BOUNDS __launch_bounds__(128, B200RT_WRAYGEN_MIN_CTAS)
#else
#define W_RAYGEN_BOUNDS __launch_bounds__(128)   // 80 registers
#endif
__global__ void W_RAYGEN_BOUNDS w_raygen_kernel(const __grid_constant__ WK k, uint32_t npix)
{
    chain_enter();
#if B200RT_WHITTED_TILED
    uint32_t tx, ty;
    tile_xy(tx, ty);
    const bool inside = tx < k.width && ty < k.height;
    const uint32_t i = inside ? ty * k.width + tx : npix;
#else
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
#endif
    const WParams& P = *k.params;
    const AccelHeader* h = (const AccelHeader*)P.handle;
    bool candidate = false;
#if B200RT_WHITTED_TILED && B200RT_WHITTED_SCREEN_RECT
    // Most pixels of a model viewer lie outside the picture of the scene bounds: they do not need a camera ray (tea<4>, two IEEE
    // divisions, a normalisation) and a slab test to learn that.  Every warp projects the eight corners of the padded bounds (lanes
    // 0-7; plain fp32 arithmetic, the rectangle is widened by two pixels) and takes their bounding rectangle in pixel coordinates: a
    // ray that reaches the box — a convex body in front of the eye — goes through a pixel inside it.  Any corner beside or behind
    // the eye, or a non-finite value: the rectangle is the whole frame.
    bool in_rect;
    {
        const uint32_t c = threadIdx.x & 7u;
        const float lx = h->bounds[0], ly = h->bounds[1], lz = h->bounds[2], hx = h->bounds[3], hy = h->bounds[4], hz = h->bounds[5];
        const float pad = fmaxf(fmaxf(hx - lx, hy - ly), hz - lz) * 2.44140625e-04f;   // four times ray_reaches_bounds' 2^-14
        const float3 q = f3(((c & 1u) ? hx + pad : lx - pad) - P.eye.x, ((c & 2u) ? hy + pad : ly - pad) - P.eye.y, ((c & 4u) ? hz + pad : lz - pad) - P.eye.z);
        // q = a U + b V + w W  ->  the pixel's (dx, dy) = (a / w, b / w): Cramer's rule, whose common denominator cancels in the ratios
        const float3 vxw = cross(P.V, P.W), qxw = cross(q, P.W), vxq = cross(P.V, q);
        const float det = dot(P.U, vxw), an = dot(q, vxw), bn = dot(P.U, qxw), wn = dot(P.U, vxq);
        const float fx = (__fdividef(an, wn) + 1.0f) * 0.5f * (float)k.width, fy = (__fdividef(bn, wn) + 1.0f) * 0.5f * (float)k.height;
        // in front of the eye by a margin: w = wn / det > 1e-6 |q| (the 1-norm stands in for the length)
        bool bad = !(wn * copysignf(1.0f, det) > 1e-6f * (fabsf(q.x) + fabsf(q.y) + fabsf(q.z)) * fabsf(det)) || !(fabsf(fx) < 1e9f) || !(fabsf(fy) < 1e9f);
        float x0 = fx, x1 = fx, y0 = fy, y1 = fy;
#pragma unroll
        for (int off = 4; off; off >>= 1) {
            x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, off)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, off));
            y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, off)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, off));
        }
        bad = __any_sync(0xffffffffu, bad);
        in_rect = bad || ((float)tx >= floorf(x0) - 2.0f && (float)tx <= floorf(x1) + 2.0f && (float)ty >= floorf(y0) - 2.0f && (float)ty <= floorf(y1) + 2.0f);
    }
#else
    const bool in_rect = true;
#endif
    if (i < npix) {
        // the miss program's read of the running mean is asked for before the camera ray is made, not after
        float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.subframe_index > 0) prev = P.accum_buffer[i];
        if (in_rect) {
            float3 o, d;
            w_camera_ray(P, k.width, k.height, i, o, d);
            candidate = ray_reaches_bounds(h, o, d, 0.0f, 1e16f);
        }
        if (!candidate) w_write_pixel(P, i, P.miss_color, prev);
    }
    if (k.raygen_traverses) {   // uniform over the launch
        if (h->anyhit == 0u) {
            WPixelWork<false> work(k, h);
            trace_one_per_thread(work, i, candidate, nullptr);
            return;
        }
#if B200RT_WHITTED_INLINE_ANYHIT
        WPixelWork<true> work(k, h);   // the candidate's any-hit program (__anyhit__radiance) runs on the traversing lane
        trace_one_per_thread(work, i, candidate, nullptr);
        return;
#endif
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, candidate);
    if (!mask) return;
    const uint32_t lane = threadIdx.x & 31u, leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&k.counters->nprimary, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (candidate) k.primary[base + __popc(mask & ((1u << lane) - 1u))] = i;
}

// ---- SHADE: __closesthit__radiance up to the shadow rays, one thread per hit slot of the level --------------------------------------
__global__ void __launch_bounds__(128) w_shade_kernel(const __grid_constant__ WK k)
{
    chain_enter();
    const WParams P = *k.params;
    const uint32_t end = min(k.counters->nslots, k.cap_slots);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (P.lights.count > k.nl_cap) k.async_flags->too_many_lights = P.lights.count;
        if (k.counters->overflow) k.async_flags->slot_overflow = 1u;
    }
    const uint32_t nl = min(P.lights.count, k.nl_cap);
    const AccelHeader* handle = (const AccelHeader*)P.handle;
    for (uint32_t si = w_level_start(k) + blockIdx.x * blockDim.x + threadIdx.x; si < end; si += gridDim.x * blockDim.x) {
        const WSlot h = k.slots[si];
        float3 org, dir;
        w_camera_ray(P, k.width, k.height, h.pixel, org, dir);
        const InstanceRecord* ir = nullptr;
        // SBT index = instance.sbtOffset + GAS-local index * RAY_TYPE_COUNT + RAY_TYPE_RADIANCE (Scene.cpp:1147-1154: sbtOffset advances by
        // primitive groups * ray types)
        uint32_t rec_idx = h.sbt * W_RAY_TYPES;
        if (handle->kind == ACCEL_KIND_IAS) {
            ir = (const InstanceRecord*)((const char*)handle + handle->inst_off) + h.inst;
            rec_idx += ir->sbt_offset;
        }
        if (rec_idx >= k.hg_count) rec_idx = k.hg_count - 1;
        const char* rec = k.hg_base + (size_t)rec_idx * k.hg_stride + B200RT_SBT_RECORD_HEADER_SIZE;
        const WGeom g = w_local_geometry(rec, h.prim, h.b1, h.b2, ir);
        const WMaterial& m = *(const WMaterial*)(rec + 112);

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
        // lights (whitted.cu:222-262): depth = payload depth + 1 = level + 1
        const uint32_t depth = h.level + 1u;
        const float3 V = neg(normalize(dir));
        for (uint32_t li = 0; li < nl; ++li) {
            const WLight L = *(const WLight*)(P.lights.data + (uint64_t)li * (P.lights.byte_stride ? P.lights.byte_stride : 36u));
            uint32_t kind = 0;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, po = a0, pd = a0;
            if (L.type == 0) {
                if (depth < W_MAX_TRACE_DEPTH) {
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
                        // result += light.color * attenuation * intensity * N_dot_L * (diff + spec): the factors are kept apart because the
                        // attenuation (known after the probe) multiplies first
                        kind = 1;
                        a0 = make_float4(L.color[0], L.color[1], L.color[2], L.intensity);
                        a1 = make_float4(diff.x + spec.x, diff.y + spec.y, diff.z + spec.z, N_dot_L);
                        po = make_float4(g.P.x, g.P.y, g.P.z, 0.001f);
                        pd = make_float4(Ld.x, Ld.y, Ld.z, L_dist - 0.001f);
                    }
                }
            } else if (L.type == 1) {
                kind = 2;
                a0 = make_float4(L.color[0] * bc.x, L.color[1] * bc.y, L.color[2] * bc.z, 0.f);
            }
            const size_t s = (size_t)si * nl + li;
            k.kinds[s] = kind;
            if (kind) k.t0[s] = a0;
            if (kind == 1) { k.t1[s] = a1; k.probes[2 * s] = po; k.probes[2 * s + 1] = pd; }
        }
        k.base[si] = make_float4(result.x, result.y, result.z, bc.w);
        k.arrived[si] = 0u;
        // ALPHA_MODE_BLEND (whitted.cu:266-286): result *= alpha; a continuation from the hit while depth < MAX_TRACE_DEPTH
        if (m.alpha_mode == 2) {
            uint32_t flags = WS_BLEND;
            if (!k.blend_levels) k.async_flags->unexpected_blend = 1u;
            else if (depth < W_MAX_TRACE_DEPTH) {
                flags |= WS_CONT;
                k.cont[atomicAdd(&k.counters->ncont, 1u)] = si;
            }
            if (h.level == 0 && k.blend_levels) k.chain[atomicAdd(&k.counters->nchain, 1u)] = si;
            k.slots[si].flags = flags;
            k.slots[si].one_minus_alpha = 1.0f - bc.w;
        }
    }
}

// ---- SHADOW: one work item per (hit slot, light) — the probe; the item that completes a slot finishes __closesthit__radiance for it ------
// The hits of a model viewer are few (tens of thousands), so the probes run one per lane for as much parallelism as there is; the lane
// that stores a slot's last attenuation sums the slot's terms in LIGHT order (not arrival order: the fp32 sum is the reference's) and,
// at level 0 without BLEND, runs the raygen tail for the pixel.
template <bool AH>
struct WShadowWork {
    static constexpr bool CONTINUES = false;
    static constexpr bool ANYHIT = AH;
    static constexpr bool COHERENT_AXES = B200RT_WHITTED_COHERENT_AXES != 0;   // probes of neighbouring hits towards the same light
    const WK& k;
    const AccelHeader* handle;
    uint32_t nl, per_slot, slot, li, level_start;
    double att;
    __device__ WShadowWork(const WK& k_, const AccelHeader* h, uint32_t nl_, uint32_t level_start_)
        : k(k_), handle(h), nl(nl_), per_slot(max(nl_, 1u)), slot(0), li(0), level_start(level_start_), att(1.0) {}

    __device__ __forceinline__ bool anyhit_enabled() const { return handle->anyhit != 0u; }
    __device__ __forceinline__ bool anyhit(uint32_t prim, uint32_t sbt, uint32_t inst, uint32_t pack, float b1, float b2, float& factor) const
    {
        return run_anyhit(AnyHitCfg{k.hg_base, k.hg_stride, k.hg_count, AH_WHITTED}, prim, sbt, w_inst_sbt(handle, inst), true, b1, b2, factor);
    }
    __device__ __forceinline__ void attenuate(float f) { att *= (double)f; }
    __device__ __forceinline__ bool stream_triangles() const { return false; }
    __device__ __forceinline__ bool fetch(uint32_t item, Trav& s, float* my_ray)
    {
        slot = level_start + item / per_slot;
        li = item % per_slot;
        att = 1.0;
        if (li < nl) {
            // kind and probe are asked for together (a term without a probe reads two lines it does not need)
            const size_t idx = (size_t)slot * nl + li;
            const uint32_t kind = k.kinds[idx];
            const float4 po = k.probes[2 * idx], pd = k.probes[2 * idx + 1];
            if (kind == 1u) {
            s.best.t = pd.w;
            // traceOcclusion (whitted_cuda.h:127-159): TERMINATE_ON_FIRST_HIT | DISABLE_CLOSESTHIT, no face culling
            if (trav_begin_handle(s, my_ray, handle, f3(po.x, po.y, po.z), f3(pd.x, pd.y, pd.z), po.w, TP_ANY, 0u, 0u)) return true;
            // nothing to traverse: the miss program commits the untouched attenuation
            }
        }
        commit(s, false);
        return false;
    }
    __device__ __forceinline__ bool next_instance(Trav& s, float* my_ray)
    {
        if (any_ray_done(s) || handle->kind != ACCEL_KIND_IAS) return false;
        const size_t idx = (size_t)slot * nl + li;
        const float4 po = k.probes[2 * idx], pd = k.probes[2 * idx + 1];
        return trav_begin_handle(s, my_ray, handle, f3(po.x, po.y, po.z), f3(pd.x, pd.y, pd.z), po.w, s.pack & (TP_ANY | TP_FOUND_ANY), 0u, s.inst + 1u);
    }
    __device__ __forceinline__ void commit(const Trav&, bool found)
    {
        // occluded -> the attenuation is never committed (0), else the pending product (whitted_cuda.h:155-158)
        const float my_att = found ? 0.0f : (float)att;
        if (li < nl) k.att[(size_t)slot * nl + li] = my_att;
        // what the item that completes the slot needs is asked for before the ticket, not behind it (with one light every item is that item)
        const float4 b = k.base[slot];
        const WSlot h = k.slots[slot];
        // release on the ticket (no __threadfence: it would invalidate the L1 the traversal lives in, common.h); the last item reads the
        // others' attenuations with volatile loads, behind the control dependency on its ticket.  One item per slot: no ticket.
        if (per_slot > 1u && atomic_add_release(&k.arrived[slot], 1u) != per_slot - 1u) return;
        // last item of the slot: result += light.color * attenuation * intensity * N_dot_L * (diff + spec) in light order (whitted.cu:249-256)
        float3 result = f3(b.x, b.y, b.z);
        for (uint32_t l = 0; l < nl; ++l) {
            const size_t idx = (size_t)slot * nl + l;
            const uint32_t kind = k.kinds[idx];
            if (kind == 2) { const float4 c = k.t0[idx]; result = f3(result.x + c.x, result.y + c.y, result.z + c.z); }
            else if (kind == 1) {
                const float a = l == li ? my_att : *(volatile const float*)&k.att[idx];
                if (a > 0.0f) {
                    const float4 c = k.t0[idx], d = k.t1[idx];
                    result = f3(result.x + (((c.x * a) * c.w) * d.w) * d.x, result.y + (((c.y * a) * c.w) * d.w) * d.y, result.z + (((c.z * a) * c.w) * d.w) * d.z);
                }
            }
        }
        if (h.flags & WS_BLEND) {
            const float alpha = b.w;  // base_color.w
            k.result[slot] = make_float4(result.x * alpha, result.y * alpha, result.z * alpha, h.one_minus_alpha);
        } else if (h.level == 0) {
            w_write_pixel(*k.params, h.pixel, result);
        } else {
            k.result[slot] = make_float4(result.x, result.y, result.z, 0.f);
        }
    }
};

template <bool AH>
__global__ void __launch_bounds__(COOP_BLOCK, W_MIN_CTAS) w_shadow_kernel(const __grid_constant__ WK k)
{
    chain_enter();
    const WParams* P = k.params;
    const AccelHeader* handle = (const AccelHeader*)P->handle;
    if ((handle->anyhit != 0u) != AH) return;
    const uint32_t end = min(k.counters->nslots, k.cap_slots);
    const uint32_t nl = min(P->lights.count, k.nl_cap);
    const uint32_t level_start = w_level_start(k);
    const uint32_t n_items = end > level_start ? (end - level_start) * max(nl, 1u) : 0u;
    WShadowWork<AH> work(k, handle, nl, level_start);
    trace_persistent(work, n_items, w_fetch(k, 1u), nullptr);
}

// The probes of an opaque scene on the one-ray-per-thread driver: the hit slots are in (roughly) pixel order, so neighbouring probes
// start next to each other and aim at the same light — coherent like the camera rays (see w_raygen_kernel).  Grid-stride over the
// items, whose count only the device knows.
__global__ void __launch_bounds__(128) w_shadow_simple_kernel(const __grid_constant__ WK k)
{
    chain_enter();
    const WParams* P = k.params;
    const AccelHeader* handle = (const AccelHeader*)P->handle;
    if (handle->anyhit != 0u) return;
    const uint32_t end = min(k.counters->nslots, k.cap_slots);
    const uint32_t nl = min(P->lights.count, k.nl_cap);
    const uint32_t level_start = w_level_start(k);
    const uint32_t n_items = end > level_start ? (end - level_start) * max(nl, 1u) : 0u;
    WShadowWork<false> work(k, handle, nl, level_start);
    // the loop is uniform over the warp: trace_one_per_thread wants all 32 lanes, the last warp's spare ones with valid = false
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_items; base += gridDim.x * blockDim.x)
        trace_one_per_thread(work, base + threadIdx.x, base + threadIdx.x < n_items, nullptr);
}

// ---- COMBINE: fold the levels of a BLEND pixel back to front, as the recursion of __closesthit__radiance returns -----------------------
__global__ void __launch_bounds__(128) w_combine_kernel(const __grid_constant__ WK k)
{
    const WParams P = *k.params;
    const uint32_t n = k.counters->nchain;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        int ids[W_MAX_TRACE_DEPTH];
        int depth = 0, tail = W_NEXT_NONE;
        for (int s = (int)k.chain[c]; depth < (int)W_MAX_TRACE_DEPTH;) {
            ids[depth++] = s;
            const int nx = k.slots[s].next;
            if (nx < 0) { tail = nx; break; }
            s = nx;
        }
        // the innermost continuation: the miss colour, or nothing when no ray was traced (depth cap / opaque hit)
        bool have = tail == W_NEXT_MISS;
        float3 R = P.miss_color;
        for (int i = depth - 1; i >= 0; --i) {
            const float4 r = k.result[ids[i]];
            const uint32_t flags = k.slots[ids[i]].flags;
            // result = result * alpha (done in SHADOW); result += payload.result * (1 - alpha) when the continuation was traced
            if ((flags & WS_BLEND) && (flags & WS_CONT) && have) R = f3(fm(R.x, r.w, r.x), fm(R.y, r.w, r.y), fm(R.z, r.w, r.z));
            else R = f3(r.x, r.y, r.z);
            have = true;
        }
        w_write_pixel(P, k.slots[ids[0]].pixel, R);
    }
}

// last kernel of a BLEND level: the next level traces the continuation rays SHADE listed, its hit slots start behind all slots so far
__global__ void w_next_level_kernel(const __grid_constant__ WK k, cudaGraphConditionalHandle cond)
{
    WCounters* c = k.counters;
    c->n_items = c->ncont;
    c->level_start = min(c->nslots, k.cap_slots);
    c->ncont = 0u;
    c->level += 1u;
    cudaGraphSetConditional(cond, (c->n_items != 0u && c->level < W_MAX_TRACE_DEPTH) ? 1u : 0u);
}

struct WLoopCache {   // the instantiated level loop of the last BLEND launch shape (one per context)
    WK key;
    unsigned grids[3] = {0, 0, 0};
    LoopGraph* loop = nullptr;
};
void whitted_release(b200rt_context ctx)
{
    if (!ctx->w_loop) return;
    delete ctx->w_loop->loop;
    delete ctx->w_loop;
    ctx->w_loop = nullptr;
}

// ---- host --------------------------------------------------------------------------------------------------------------------------
static unsigned w_persistent_grid(b200rt_context ctx, const void* kernel, uint64_t n)
{
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, COOP_BLOCK, 0);
    const uint64_t cap = (uint64_t)std::max(occ, 1) * ctx->sm_count;
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(cap, (n + COOP_BLOCK - 1) / COOP_BLOCK));
}

int launch_whitted(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr d_params, const b200rt_shader_binding_table* sbt, unsigned width, unsigned height)
{
    B2_REQUIRE(ctx, d_params && sbt, "null argument");
    B2_REQUIRE(ctx, sbt->hitgroupRecordBase && sbt->hitgroupRecordCount > 0 && sbt->hitgroupRecordStrideInBytes >= 32 + 352,
               "hit-group records (whitted::HitGroupData, 2 per primitive group) are required");
    const uint64_t npix64 = (uint64_t)width * height;
    B2_REQUIRE(ctx, npix64 < (1ull << 28), "launch too large");
    if (npix64 == 0) return 0;
    DeviceGuard guard(ctx->device);  // the C ABI entry point holds ctx->mu
    // errors of earlier launches on this context surface here (no synchronisation on the steady-state path)
    WAsyncFlags* flags = (WAsyncFlags*)((char*)ctx->pinned + 1024);
    if (flags->unexpected_blend) {
        flags->unexpected_blend = 0;
        ctx->w_params = 0;  // look at the hit-group records again
        return set_error(ctx, B200RT_ERROR_INVALID_OPERATION,
                         "an earlier whitted launch met an ALPHA_MODE_BLEND material that was not in the hit-group records when they were first seen "
                         "(its continuation was not traced); relaunch");
    }
    if (flags->too_many_lights) {
        ctx->w_params = 0;  // re-read the light count below
        flags->too_many_lights = 0;
        return set_error(ctx, B200RT_ERROR_INVALID_OPERATION, "an earlier whitted launch found more lights than its workspace was sized for (frame incomplete); relaunch");
    }
    if (flags->slot_overflow) {
        flags->slot_overflow = 0;
        log_msg(ctx, 2, "whitted", "an earlier launch ran out of hit slots for BLEND continuations (more than %u levels per pixel on average); the deepest were dropped",
                W_BLEND_SLOT_FACTOR);
    }
    // The workspace is sized by the light count and by whether BLEND materials exist.  Both are read back (one stream synchronisation)
    // on the first launch with a given (d_params, hit-group records); later launches reuse them and the kernels check the live values.
    if (ctx->w_params != d_params || ctx->w_sbt != sbt->hitgroupRecordBase || ctx->w_sbt_count != sbt->hitgroupRecordCount) {
        WParams hp;
        B2_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, (const void*)d_params, sizeof(WParams), cudaMemcpyDeviceToHost, s));
        const size_t rec_bytes = (size_t)sbt->hitgroupRecordCount * sbt->hitgroupRecordStrideInBytes;
        std::vector<char> recs(rec_bytes);
        B2_CUDA(ctx, cudaMemcpyAsync(recs.data(), (const void*)sbt->hitgroupRecordBase, rec_bytes, cudaMemcpyDeviceToHost, s));
        B2_CUDA(ctx, cudaStreamSynchronize(s));
        memcpy(&hp, ctx->pinned, sizeof(WParams));
        B2_REQUIRE(ctx, hp.handle && hp.accum_buffer, "LaunchParams has null pointers");
        B2_REQUIRE(ctx, hp.lights.count <= 64 && (hp.lights.count == 0 || hp.lights.data), "bad light list");
        AccelHeader ah;
        B2_CUDA(ctx, cudaMemcpyAsync(&ah, (const void*)hp.handle, sizeof(ah), cudaMemcpyDeviceToHost, s));
        B2_CUDA(ctx, cudaStreamSynchronize(s));
        B2_REQUIRE(ctx, ah.magic == ACCEL_MAGIC, "LaunchParams.handle is not a b200rt traversable");
        bool blend = false;
        for (unsigned r = 0; r < sbt->hitgroupRecordCount; ++r)
            blend |= ((const WMaterial*)(recs.data() + (size_t)r * sbt->hitgroupRecordStrideInBytes + 32 + 112))->alpha_mode == 2;
        ctx->w_params = d_params;
        ctx->w_sbt = sbt->hitgroupRecordBase;
        ctx->w_sbt_count = sbt->hitgroupRecordCount;
        ctx->w_lights = hp.lights.count;
        ctx->w_blend = blend;
    }
    const uint32_t npix = (uint32_t)npix64, nl = ctx->w_lights, nlp = std::max(nl, 1u);
    const bool blend = ctx->w_blend;
    const void* k_primary = (const void*)w_primary_kernel<true>;   // grid sizing: the instantiation with the smaller occupancy
    const void* k_shadow = (const void*)w_shadow_kernel<true>;
    const size_t cap = (size_t)npix * (blend ? W_BLEND_SLOT_FACTOR : 1u);
    size_t off = 16384;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_cnt = take(256), o_slots = take(sizeof(WSlot) * cap), o_base = take(16 * cap), o_t0 = take(16 * cap * nlp), o_t1 = take(16 * cap * nlp),
                 o_kinds = take(4 * cap * nlp), o_probes = take(32 * cap * nlp), o_result = take(blend ? 16 * cap : 0), o_cont = take(blend ? 4 * cap : 0),
                 o_chain = take(blend ? 4 * (size_t)npix : 0), o_primary = take(4 * (size_t)npix), o_att = take(4 * cap * nlp), o_arr = take(4 * cap);
    int rc = ensure_workspace(ctx, off, s);
    if (rc) return rc;
    ws_acquire(ctx, s);
    char* W = (char*)ctx->ws.ptr;
    WK k;
    memset(&k, 0, sizeof(k));  // compared byte-wise with the cached loop's key
    k.params = (const WParams*)d_params;
    k.width = width; k.height = height; k.nl_cap = nl; k.level = 0; k.level_start = 0; k.cap_slots = (uint32_t)cap;
    k.hg_base = (const char*)sbt->hitgroupRecordBase; k.hg_stride = sbt->hitgroupRecordStrideInBytes; k.hg_count = sbt->hitgroupRecordCount;
    k.blend_levels = blend ? 1u : 0u;
    k.counters = (WCounters*)(W + o_cnt);
    k.slots = (WSlot*)(W + o_slots); k.base = (float4*)(W + o_base); k.t0 = (float4*)(W + o_t0); k.t1 = (float4*)(W + o_t1);
    k.kinds = (uint32_t*)(W + o_kinds); k.probes = (float4*)(W + o_probes); k.result = (float4*)(W + o_result); k.cont = (uint32_t*)(W + o_cont);
    k.chain = (uint32_t*)(W + o_chain);
    k.primary = (uint32_t*)(W + o_primary);
    k.att = (float*)(W + o_att);
    k.arrived = (unsigned int*)(W + o_arr);
    k.async_flags = flags;
    unsigned int* cursors = (unsigned int*)(W + o_cnt + 64);  // two per level
    B2_CUDA(ctx, cudaMemsetAsync(W + o_cnt, 0, 256, s));      // counters, loop state and all work-item cursors of the frame
    k.fetch = cursors;
    // B200RT_WHITTED_INLINE: opaque scenes on the one-ray-per-thread kernels (see w_raygen_kernel)
    // bit 0: RAYGEN traverses its own camera ray, bit 1: the probes run one per thread
    static const int inline_mask = [] { const char* e = getenv("B200RT_WHITTED_INLINE"); return e ? atoi(e) : B200RT_WHITTED_INLINE_DEFAULT; }();
    const bool inline_primary = (inline_mask & 1) != 0, inline_shadow = (inline_mask & 2) != 0;
    k.raygen_traverses = (!blend && inline_primary) ? 1u : 0u;
    // the kernels of a frame without BLEND levels are a chain of programmatic dependent launches (common.h: launch_chain)
#if B200RT_WHITTED_TILED
    B2_CUDA(ctx, launch_chain(w_raygen_kernel, dim3(div_up(k.width, CTA_TILE_W), div_up(k.height, CTA_TILE_H)), dim3(TILE_CTA_THREADS), s, k, npix));
#else
    B2_CUDA(ctx, launch_chain(w_raygen_kernel, dim3(div_up(npix, 128)), dim3(128), s, k, npix));
#endif
    B2_LAUNCH_CHECK(ctx);
    const unsigned g_primary = w_persistent_grid(ctx, k_primary, npix), g_shadow = w_persistent_grid(ctx, k_shadow, (uint64_t)npix * nlp);
    const unsigned g_shade = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(div_up(npix, 128), (uint64_t)ctx->sm_count * 16));
    if (!blend) {
        // Both instantiations of the persistent kernels are enqueued: which kind the traversable is, is read on the device
        // (AccelHeader::anyhit), each kernel returns at once unless it is its kind, so the choice cannot go stale when another handle is
        // written into the same LaunchParams.  (B200RT_WHITTED_INLINE: opaque scenes on the one-ray-per-thread kernels instead.)
        // RAYGEN traverses the camera rays of either kind of scene itself (B200RT_WHITTED_INLINE bit 0 with INLINE_ANYHIT): nothing is left
        // on the candidate list, so neither PRIMARY instantiation is launched — a fact of the configuration, not of the scene
        if (!(inline_primary && B200RT_WHITTED_INLINE_ANYHIT)) { B2_CUDA(ctx, launch_chain(w_primary_kernel<true>, dim3(g_primary), dim3(COOP_BLOCK), s, k, npix)); B2_LAUNCH_CHECK(ctx); }
        if (!inline_primary) { B2_CUDA(ctx, launch_chain(w_primary_kernel<false>, dim3(g_primary), dim3(COOP_BLOCK), s, k, npix)); B2_LAUNCH_CHECK(ctx); }
        B2_CUDA(ctx, launch_chain(w_shade_kernel, dim3(g_shade), dim3(128), s, k));
        B2_LAUNCH_CHECK(ctx);
        if (!inline_shadow) B2_CUDA(ctx, launch_chain(w_shadow_kernel<false>, dim3(g_shadow), dim3(COOP_BLOCK), s, k));
        else B2_CUDA(ctx, launch_chain(w_shadow_simple_kernel, dim3(std::max(1u, std::min(div_up((uint64_t)npix * nlp, 128), (unsigned)ctx->sm_count * 8u))), dim3(128), s, k));
        B2_LAUNCH_CHECK(ctx);
        B2_CUDA(ctx, launch_chain(w_shadow_kernel<true>, dim3(g_shadow), dim3(COOP_BLOCK), s, k));
        B2_LAUNCH_CHECK(ctx);
    } else {
        // BLEND scenes: how many continuation rays a level starts is known on the device only, so the levels run as a device-side loop
        // (CUDA graph, conditional WHILE: loop_graph.h) — no host read-back, the launch stays asynchronous.  The loop is instantiated
        // once per launch shape and relaunched every subframe.
        if (!ctx->w_loop) ctx->w_loop = new WLoopCache();
        WLoopCache& lc = *ctx->w_loop;
        if (!lc.loop || memcmp(&lc.key, &k, sizeof(WK)) || lc.grids[0] != g_primary || lc.grids[1] != g_shade || lc.grids[2] != g_shadow) {
            if (lc.loop) { park_loop(ctx, lc.loop); lc.loop = nullptr; }   // may still be running: destroyed once it has finished
            retire_loops(ctx, false);
            LoopGraph* g = new LoopGraph(ctx);
            lc.loop = g;
            memcpy(&lc.key, &k, sizeof(WK));
            lc.grids[0] = g_primary; lc.grids[1] = g_shade; lc.grids[2] = g_shadow;
            uint32_t n0 = npix;
            if ((rc = g->begin())) return rc;
            if ((rc = g->add((const void*)w_primary_kernel<false>, g_primary, COOP_BLOCK, 0, k, n0))) return rc;
            if ((rc = g->add((const void*)w_primary_kernel<true>, g_primary, COOP_BLOCK, 0, k, n0))) return rc;
            if ((rc = g->add((const void*)w_shade_kernel, g_shade, 128, 0, k))) return rc;
            if ((rc = g->add((const void*)w_shadow_kernel<false>, g_shadow, COOP_BLOCK, 0, k))) return rc;
            if ((rc = g->add((const void*)w_shadow_kernel<true>, g_shadow, COOP_BLOCK, 0, k))) return rc;
            if ((rc = g->add((const void*)w_next_level_kernel, 1, 1, 0, k, g->cond()))) return rc;
        }
        if ((rc = lc.loop->launch(s))) return rc;
        ctx->launches += 1;
    }
    if (blend) {
        w_combine_kernel<<<(unsigned)std::max<uint64_t>(1, std::min<uint64_t>(div_up(npix, 128), (uint64_t)ctx->sm_count * 8)), 128, 0, s>>>(k);
        B2_LAUNCH_CHECK(ctx);
    }
    ws_release(ctx, s);
    return 0;
}

// sutil::Scene::addImage + addSampler (SDK/sutil/Scene.cpp:576-652): 8-bit RGBA image -> CUDA array -> texture object with
// normalised coordinates, normalised-float reads, the given address modes and filter.  Returns the cudaTextureObject_t.
// A texture object of THIS context's device over an existing CUDA array — the array may live on another device the context's device
// has peer access to: optixNVLink keeps one copy of a texture per P2P island and gives every device of the island its own sampler over
// it (defineTextureOnDevice / loadTexture, SDK/optixNVLink/optixNVLink.cpp:1446-1468,1522-1561).  Same sampler state as texture_create.
int texture_view(b200rt_context ctx, uint64_t cuda_array, int address_s, int address_t, int linear, uint64_t* tex_out)
{
    B2_REQUIRE(ctx, cuda_array && tex_out, "bad argument");
    DeviceGuard guard(ctx->device);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = (cudaArray_t)(uintptr_t)cuda_array;
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
    const cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    if (e != cudaSuccess) return set_error(ctx, B200RT_ERROR_CUDA_ERROR, "cudaCreateTextureObject: %s", cudaGetErrorString(e));
    *tex_out = (uint64_t)tex;
    return 0;
}

int texture_create(b200rt_context ctx, int width, int height, const void* rgba8, int address_s, int address_t, int linear, uint64_t* tex_out,
                   uint64_t* array_out)
{
    B2_REQUIRE(ctx, width > 0 && height > 0 && rgba8 && tex_out && array_out, "bad argument");
    DeviceGuard guard(ctx->device);
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
    cudaArray_t arr = nullptr;
    B2_CUDA(ctx, cudaMallocArray(&arr, &cd, width, height));
    {
        const cudaError_t e = cudaMemcpy2DToArray(arr, 0, 0, rgba8, (size_t)width * 4, (size_t)width * 4, height, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFreeArray(arr); return set_error(ctx, B200RT_ERROR_CUDA_ERROR, "texture upload: %s", cudaGetErrorString(e)); }
    }
    const int rc = texture_view(ctx, (uint64_t)(uintptr_t)arr, address_s, address_t, linear, tex_out);
    if (rc) { cudaFreeArray(arr); return rc; }
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
