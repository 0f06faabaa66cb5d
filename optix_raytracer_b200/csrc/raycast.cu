// raycast.cu — the optixRaycasting path: ray-buffer generation, closest-hit queries over a
// traversable, the sample's closest-hit/miss programs, and the hit shading kernel.
//
// Reference: SDK/optixRaycasting/optixRaycastingKernels.cu:42-115 (createRaysOrtho / translateRays /
// shadeHits), SDK/optixRaycasting/optixRaycasting.cu:45-86 (raygen/miss/closest-hit),
// SDK/cuda/LocalGeometry.h:59-176 (shading normal), SDK/optixRaycasting/optixRaycasting.cpp:289-317 (launch).
// Rays are 32-byte AoS records and hits 16-byte records exactly as in the reference, so a warp reads
// 1 KiB and writes 512 B contiguous per request; each ray record is fetched as two 16-byte loads.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "accel.h"
#include "internal.h"
#include "trav_coop.cuh"
#include "anyhit.cuh"

#ifndef B200RT_RAY_BOUNDS
#define B200RT_RAY_BOUNDS 1   // drop the rays of a buffer that pass the scene by at its bounds (trav_coop.cuh: trav_begin<BOUNDS>)
#endif

namespace b200rt {

struct RayRec { float ox, oy, oz, tmin, dx, dy, dz, tmax; };  // optixRaycastingKernels.h:35-41

// ---- createRaysOrthoKernel (optixRaycastingKernels.cu:42-55): origin = (x0 + ix*dx, y0 + iy*dy, z) ----
__global__ void __launch_bounds__(512) create_rays_ortho_kernel(float4* __restrict__ rays, int width, int height, float x0, float y0,
                                                                 float z, float dx, float dy)
{
    const int rayx = threadIdx.x + blockIdx.x * blockDim.x;
    const int rayy = threadIdx.y + blockIdx.y * blockDim.y;
    if (rayx >= width || rayy >= height) return;
    const size_t idx = (size_t)rayx + (size_t)rayy * width;
    rays[2 * idx + 0] = make_float4(fm((float)rayx, dx, x0), fm((float)rayy, dy, y0), z, 0.0f);
    rays[2 * idx + 1] = make_float4(0.0f, 0.0f, 1.0f, 1e34f);
}

__global__ void __launch_bounds__(512) translate_rays_kernel(float4* __restrict__ rays, int count, float3 off)
{
    const int idx = threadIdx.x + blockIdx.x * blockDim.x;
    if (idx >= count) return;
    float4 o = rays[2 * (size_t)idx];
    o.x += off.x; o.y += off.y; o.z += off.z;
    rays[2 * (size_t)idx] = o;
}

// shadeHitsKernel (optixRaycastingKernels.cu:91-106): 0.5*N + 0.5, background 0.2
__global__ void __launch_bounds__(512) shade_hits_kernel(float* __restrict__ image, int count, const float4* __restrict__ hits)
{
    const int idx = threadIdx.x + blockIdx.x * blockDim.x;
    if (idx >= count) return;
    const float4 h = hits[idx];
    float3 c;
    if (h.x < 0.0f) c = f3(0.2f, 0.2f, 0.2f);
    else c = f3(fm(0.5f, h.y, 0.5f), fm(0.5f, h.z, 0.5f), fm(0.5f, h.w, 0.5f));
    image[3 * (size_t)idx + 0] = c.x;
    image[3 * (size_t)idx + 1] = c.y;
    image[3 * (size_t)idx + 2] = c.z;
}

// ---- ray-buffer queries on the persistent traversal driver (trav_coop.cuh) -----------------------------
struct BufView { uint64_t data; uint32_t count; uint16_t byte_stride; uint16_t elmt; };  // SDK/cuda/BufferView.h:32-38

__device__ __forceinline__ ExtHit make_ext(const Trav& s, bool found)
{
    ExtHit e;
    if (found) { e.t = s.best.t; e.prim = s.best.prim; e.inst = s.best.inst; e.b1 = s.best.b1; e.b2 = s.best.b2; }
    else { e.t = -1.0f; e.prim = 0xffffffffu; e.inst = 0xffffffffu; e.b1 = 0.f; e.b2 = 0.f; }
    return e;
}

// KIND 0: closest hit -> ExtHit; 1: any hit -> u32 flag (AH: fp32 attenuation, whitted_cuda.h:127-159); 2: the optixRaycasting
// programs -> Hit (+ optional ExtHit).  AH: the launch has any-hit programs (anyhit.cuh) and the scene may hold geometry that runs them.
template <int KIND, bool AH>
struct RayWork {
    static constexpr bool CONTINUES = false;
    static constexpr bool ANYHIT = AH;
    static constexpr bool SMEM_STATE = true;  // commit wants the whole hit record: keep it in the lane's shared slot (trav_coop.cuh)
    AnyHitCfg ah;
    double att;  // AH, KIND 1: pending occlusion attenuation of this lane's ray
    const AccelHeader* handle;
    const float4* rays;
    uint32_t ray_flags;
    ExtHit* ext;
    uint32_t* occluded;
    float4* hits;
    const char* hg_base;
    uint32_t hg_stride, hg_count;
    uint32_t flag_period;  // > 0: every flag_period-th ray (the last of each group) ignores the CULL_*_ANYHIT ray flags (playground.cu)
    uint64_t item;

    __device__ __forceinline__ uint32_t cull_flags(uint32_t i) const
    {
        const uint32_t f = ray_flags & 0xff00f3u;  // CULL_* and DISABLE / ENFORCE_ANYHIT (accel.h: cull_word), visibility mask (ray_visibility)
        return (flag_period && (i % flag_period) == flag_period - 1u) ? (f & 0xff0033u) : f;
    }
    __device__ __forceinline__ bool anyhit(uint32_t prim, uint32_t sbt, uint32_t inst, uint32_t pack, float b1, float b2, float& factor) const
    {
        uint32_t inst_sbt = 0;
        if (handle->kind == ACCEL_KIND_IAS) inst_sbt = ((const InstanceRecord*)((const char*)handle + handle->inst_off) + inst)->sbt_offset;
        return run_anyhit(ah, prim, sbt, inst_sbt, (pack & TP_ANY) != 0u, b1, b2, factor);
    }
    __device__ __forceinline__ void attenuate(float factor) { att *= (double)factor; }
    __device__ __forceinline__ bool stream_triangles() const { return false; }
    __device__ __forceinline__ bool anyhit_enabled() const { return ah.mode != AH_NONE && handle->anyhit != 0u; }
    __device__ __forceinline__ bool fetch(uint32_t i, Trav& s, float* my_ray)
    {
        item = i;
        if (AH) att = 1.0;
        const float4 a = __ldg(rays + 2 * (size_t)i), b = __ldg(rays + 2 * (size_t)i + 1);
        s.best.t = b.w;
        // rays of a buffer come from anywhere: those that pass the scene (or an instance) by are dropped at its bounds
        if (!trav_begin_handle<B200RT_RAY_BOUNDS != 0, SMEM_STATE>(s, my_ray, handle, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, KIND == 1 ? TP_ANY : 0u, cull_flags(i), 0u)) {
            commit(s, false);
            return false;
        }
        return true;
    }
    __device__ __forceinline__ bool next_instance(Trav& s, float* my_ray)
    {
        if (handle->kind == ACCEL_KIND_GAS || any_ray_done(s)) return false;
        const float4 a = __ldg(rays + 2 * item), b = __ldg(rays + 2 * item + 1);
        return trav_begin_handle<B200RT_RAY_BOUNDS != 0, SMEM_STATE>(s, my_ray, handle, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, s.pack & (TP_ANY | TP_FOUND_ANY),
                                 cull_flags((uint32_t)item), s.inst + 1u);
    }
    __device__ __forceinline__ void commit(const Trav& s, bool found)
    {
        if (KIND == 1) {
            // AH: the committed attenuation (0 = occluded: the miss program never ran); otherwise a flag
            if (AH) ((float*)occluded)[item] = found ? 0.0f : (float)att;
            else occluded[item] = found ? 1u : 0u;
            return;
        }
        if (KIND == 0) {
            ext[item] = make_ext(s, found);
            if (occluded) occluded[item] = found ? (s.best.sbt & TRI_SBT_MASK) : 0u;  // optional: GAS-local SBT index of the hit (whitted.cu)
            return;
        }
        // KIND 2: __miss__buffer_miss / __closesthit__buffer_hit (optixRaycasting.cu:65-86)
        float4 out;
        if (!found) {
            out = make_float4(-1.0f, 1.0f, 0.0f, 0.0f);
        } else {
            // SBT record = instance.sbtOffset + GAS-local index (ray type 0, stride 1)
            uint32_t sbt = s.best.sbt & TRI_SBT_MASK;
            const InstanceRecord* ir = nullptr;
            if (handle->kind == ACCEL_KIND_IAS) {
                ir = (const InstanceRecord*)((const char*)handle + handle->inst_off) + s.best.inst;
                sbt += ir->sbt_offset;
            }
            if (sbt >= hg_count) sbt = hg_count - 1;
            const char* rec = hg_base + (size_t)sbt * hg_stride + B200RT_SBT_RECORD_HEADER_SIZE;
            // whitted::HitGroupData -> GeometryData{type @0, union @16 (16-byte aligned): TriangleMesh{indices @16, positions @32,
            // normals @48, texcoords[2] @64/@80, colors @96}} — offsets measured on the reference headers (oracle/ref_shim.cpp,
            // tests/golden/kat.json "hitgroup_layout")
            const BufView vi = *(const BufView*)(rec + 16), vp = *(const BufView*)(rec + 32), vn = *(const BufView*)(rec + 48);
            const uint32_t prim = s.best.prim;
            uint32_t i0, i1, i2;
            if (vi.elmt == 4) { const uint32_t* ip = (const uint32_t*)vi.data + 3 * (size_t)prim; i0 = ip[0]; i1 = ip[1]; i2 = ip[2]; }
            else if (vi.elmt == 2) { const uint16_t* ip = (const uint16_t*)vi.data + 3 * (size_t)prim; i0 = ip[0]; i1 = ip[1]; i2 = ip[2]; }
            else { i0 = 3 * prim; i1 = i0 + 1; i2 = i0 + 2; }
            float3 N;
            if (vn.data) {
                const uint32_t st = vn.byte_stride ? vn.byte_stride : 12u;
                const float* n0 = (const float*)(vn.data + (size_t)i0 * st);
                const float* n1 = (const float*)(vn.data + (size_t)i1 * st);
                const float* n2 = (const float*)(vn.data + (size_t)i2 * st);
                const float b1 = s.best.b1, b2 = s.best.b2;
                const float b0 = (1.0f - b1) - b2;
                N = f3(fm(b2, n2[0], fm(b1, n1[0], b0 * n0[0])), fm(b2, n2[1], fm(b1, n1[1], b0 * n0[1])), fm(b2, n2[2], fm(b1, n1[2], b0 * n0[2])));
            } else {
                const uint32_t st = vp.byte_stride ? vp.byte_stride : 12u;
                const float* p0 = (const float*)(vp.data + (size_t)i0 * st);
                const float* p1 = (const float*)(vp.data + (size_t)i1 * st);
                const float* p2 = (const float*)(vp.data + (size_t)i2 * st);
                const float3 P0 = f3(p0[0], p0[1], p0[2]);
                N = cross(f3(p1[0], p1[1], p1[2]) - P0, f3(p2[0], p2[1], p2[2]) - P0);
            }
            if (ir) N = xform_normal(ir->inv, N);
            N = normalize(N);
            // `const unsigned int t = optixGetRayTmax();` — the reference truncates t to an integer
            out = make_float4((float)(unsigned int)s.best.t, N.x, N.y, N.z);
        }
        hits[item] = out;
        if (ext) ext[item] = make_ext(s, found);
    }
};

struct RaycastParamsDev { uint64_t handle; const float4* rays; float4* hits; };  // optixRaycasting.h:41-46

#ifndef B200RT_RAY_MIN_CTAS
#define B200RT_RAY_MIN_CTAS 8
#endif
// AH: this instantiation carries the any-hit machinery (anyhit.cuh).  Launches that HAVE any-hit programs (KIND 2 with records long
// enough to hold material and texture coordinates; the whitted stages) enqueue BOTH instantiations with DUAL set: each reads
// AccelHeader::anyhit on the device and returns at once unless the traversable is its kind — scenes without any-hit geometry run the
// plain kernel at its full speed (a merged kernel measured 8 % slower on the Duck ray buffers: twice the shared memory, more spills), the
// other launch costs about two microseconds, and no decision rests on what the host remembers of an earlier launch (another handle may
// have been written into the same Params block).
template <int KIND, bool STATS, bool AH>
__global__ void __launch_bounds__(COOP_BLOCK, B200RT_RAY_MIN_CTAS) trace_rays_kernel(const AccelHeader* __restrict__ handle, const float4* __restrict__ rays, uint32_t n,
                                                          uint32_t ray_flags, ExtHit* __restrict__ ext, uint32_t* __restrict__ occluded,
                                                          const RaycastParamsDev* __restrict__ rc_params, const char* __restrict__ hg_base,
                                                          uint32_t hg_stride, uint32_t hg_count, unsigned int* __restrict__ counter,
                                                          unsigned long long* __restrict__ stats, const unsigned int* __restrict__ n_dev,
                                                          uint32_t n_mult, uint32_t flag_period, AnyHitCfg ah, uint32_t dual)
{
    chain_enter();
    if (n_dev) n = min(n, *n_dev * n_mult);  // ray count produced on the device by an earlier stage (playground.cu)
    if (KIND != 2 && hg_base) handle = (const AccelHeader*)*(const uint64_t*)hg_base;  // traversable handle read from device memory (whitted.cu)
    RayWork<KIND, AH> w;
    w.ah = ah;
    w.att = 1.0;
    w.flag_period = flag_period;
    w.handle = handle; w.rays = rays; w.hits = nullptr;
    if (KIND == 2) { const RaycastParamsDev P = *rc_params; w.handle = (const AccelHeader*)P.handle; w.rays = P.rays; w.hits = P.hits; }
    if (dual && (w.handle->anyhit != 0u) != AH) return;
    w.ray_flags = ray_flags; w.ext = ext; w.occluded = occluded;
    w.hg_base = hg_base; w.hg_stride = hg_stride; w.hg_count = hg_count; w.item = 0;
    TravStats st{0, 0};
    trace_persistent(w, n, counter, STATS ? &st : nullptr);
    if (STATS) {
        for (int off = 16; off; off >>= 1) {
            st.nodes += __shfl_xor_sync(0xffffffffu, st.nodes, off);
            st.tris += __shfl_xor_sync(0xffffffffu, st.tris, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&stats[0], (unsigned long long)st.nodes);
            atomicAdd(&stats[1], (unsigned long long)st.tris);
        }
    }
}

// optixRaycasting's launch on the one-ray-per-thread driver (trav_coop.cuh: trace_one_per_thread): the sample's ray buffers are coherent
// (orthographic grid in pixel order, optixRaycastingKernels.cu:42-55).  Scenes with any-hit geometry return at once, like the plain
// cooperative kernel does, and the any-hit cooperative kernel enqueued behind this one takes them.
// TILED: the launch is a width x height grid (optixLaunch's dimensions; the raygen program's ray index is y * width + x,
// optixRaycasting.cu:47-51) and a warp takes an 8 x 4 tile of launch indices instead of 32 consecutive ones (trav_coop.cuh: tile_xy) — rays of a
// tile walk the same nodes more often than rays of a 32 x 1 strip (Duck ray buffers 0.247 -> 0.210 ms).
#ifdef B200RT_RAYCAST_MIN_CTAS
#define RAYCAST_SIMPLE_BOUNDS __launch_bounds__(128, B200RT_RAYCAST_MIN_CTAS)
#else
#define RAYCAST_SIMPLE_BOUNDS __launch_bounds__(128)   // 72 registers
#endif
template <bool TILED, bool AH>
__global__ void RAYCAST_SIMPLE_BOUNDS raycast_simple_kernel(const RaycastParamsDev* __restrict__ rc_params, uint32_t n, uint32_t width, uint32_t height,
                                                              ExtHit* __restrict__ ext, const char* __restrict__ hg_base, uint32_t hg_stride, uint32_t hg_count,
                                                              uint32_t dual)
{
    chain_enter();
    const RaycastParamsDev P = *rc_params;
    RayWork<2, AH> w;
    w.ah = AH ? AnyHitCfg{hg_base, hg_stride, hg_count, AH_TEXTURE_MASK} : AnyHitCfg{nullptr, 0u, 0u, AH_NONE};
    w.att = 1.0;
    w.flag_period = 0;
    w.handle = (const AccelHeader*)P.handle; w.rays = P.rays; w.hits = P.hits;
    if (dual && (w.handle->anyhit != 0u) != AH) return;
    w.ray_flags = 0u; w.ext = ext; w.occluded = nullptr;
    w.hg_base = hg_base; w.hg_stride = hg_stride; w.hg_count = hg_count; w.item = 0;
    if (AH) {
        // The any-hit instantiation is enqueued behind the plain one for every launch whose records could hold an any-hit program and
        // returns above unless the scene has such geometry: it is launched as a small 1-D grid that walks the CTA tiles (rows of 128
        // indices), so that the launch that does nothing is a thousand CTAs, not one per tile.
        if (TILED) {
            const uint32_t tiles_x = (width + CTA_TILE_W - 1u) / CTA_TILE_W, tiles = tiles_x * ((height + CTA_TILE_H - 1u) / CTA_TILE_H);
            for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                uint32_t x, y;
                tile_xy(t % tiles_x, t / tiles_x, x, y);
                trace_one_per_thread(w, y * width + x, x < width && y < height, nullptr);
            }
        } else {
            for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
                trace_one_per_thread(w, base + threadIdx.x, base + threadIdx.x < n, nullptr);
        }
    } else if (TILED) {
        uint32_t x, y;
        tile_xy(x, y);
        trace_one_per_thread(w, y * width + x, x < width && y < height, nullptr);
    } else {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        trace_one_per_thread(w, i, i < n, nullptr);
    }
}

#ifndef B200RT_RAYCAST_TILED
#define B200RT_RAYCAST_TILED 1
#endif
// a zeroed fetch counter for one persistent launch: slots rotate so launches on different streams do not share one
static int next_counter(b200rt_context ctx, cudaStream_t s, unsigned int** out)
{
    int rc = ensure_workspace(ctx, 1 << 20, s);
    if (rc) return rc;
    unsigned int* base = (unsigned int*)((char*)ctx->ws.ptr + 4096);
    unsigned int* c = base + (ctx->counter_slot++ % 512u) * 4u;
    B2_CUDA(ctx, cudaMemsetAsync(c, 0, sizeof(unsigned int), s));
    *out = c;
    return 0;
}

template <int KIND, bool STATS, bool AH = false>
static unsigned persistent_grid_rays(b200rt_context ctx, uint64_t n)
{
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_rays_kernel<KIND, STATS, AH>, COOP_BLOCK, 0);
    const uint64_t cap = (uint64_t)std::max(occ, 1) * ctx->sm_count;
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(cap, (n + COOP_BLOCK - 1) / COOP_BLOCK));
}

int create_rays_ortho(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr rays, int width, int height, const float* bbmin,
                      const float* bbmax, float padding)
{
    B2_REQUIRE(ctx, rays && width > 0 && height > 0 && bbmin && bbmax, "bad argument");
    DeviceGuard guard(ctx->device);
    // host scalars exactly as createRaysOrthoOnDevice (optixRaycastingKernels.cu:59-66); host code, no contraction
    const float sx = bbmax[0] - bbmin[0], sy = bbmax[1] - bbmin[1], sz = bbmax[2] - bbmin[2];
    const float dx = sx * (1 + 2 * padding) / width;
    const float dy = sy * (1 + 2 * padding) / height;
    const float x0 = bbmin[0] - sx * padding + dx / 2;
    const float y0 = bbmin[1] - sy * padding + dy / 2;
    const float z = bbmin[2] - fmaxf(sz, 1.0f) * .001f;
    dim3 block(32, 16), grid(div_up(width, 32), div_up(height, 16));
    create_rays_ortho_kernel<<<grid, block, 0, s>>>((float4*)rays, width, height, x0, y0, z, dx, dy);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int translate_rays(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr rays, int count, const float* off)
{
    B2_REQUIRE(ctx, rays && count >= 0 && off, "bad argument");
    DeviceGuard guard(ctx->device);
    if (count == 0) return 0;
    translate_rays_kernel<<<div_up(count, 512), 512, 0, s>>>((float4*)rays, count, make_float3(off[0], off[1], off[2]));
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int shade_hits(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr image, int count, b200rt_deviceptr hits)
{
    B2_REQUIRE(ctx, image && hits && count >= 0, "bad argument");
    DeviceGuard guard(ctx->device);
    if (count == 0) return 0;
    shade_hits_kernel<<<div_up(count, 512), 512, 0, s>>>((float*)image, count, (const float4*)hits);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int trace_closest(b200rt_context ctx, cudaStream_t s, b200rt_traversable handle, b200rt_deviceptr rays, uint64_t n, unsigned ray_flags,
                  b200rt_deviceptr ext)
{
    B2_REQUIRE(ctx, handle && (n == 0 || (rays && ext)) && n < (1ull << 32), "bad argument");
    DeviceGuard guard(ctx->device);
    if (n == 0) return 0;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    unsigned int* counter = nullptr;
    int rc = next_counter(ctx, s, &counter);
    if (rc) return rc;
    trace_rays_kernel<0, false, false><<<persistent_grid_rays<0, false>(ctx, n), COOP_BLOCK, 0, s>>>(
        (const AccelHeader*)handle, (const float4*)rays, (uint32_t)n, ray_flags, (ExtHit*)ext, nullptr, nullptr, nullptr, 0, 0, counter, nullptr, nullptr, 1u, 0u,
        AnyHitCfg{nullptr, 0u, 0u, AH_NONE}, 0u);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int trace_any(b200rt_context ctx, cudaStream_t s, b200rt_traversable handle, b200rt_deviceptr rays, uint64_t n, unsigned ray_flags,
              b200rt_deviceptr occ)
{
    B2_REQUIRE(ctx, handle && (n == 0 || (rays && occ)) && n < (1ull << 32), "bad argument");
    DeviceGuard guard(ctx->device);
    if (n == 0) return 0;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    unsigned int* counter = nullptr;
    int rc = next_counter(ctx, s, &counter);
    if (rc) return rc;
    trace_rays_kernel<1, false, false><<<persistent_grid_rays<1, false>(ctx, n), COOP_BLOCK, 0, s>>>(
        (const AccelHeader*)handle, (const float4*)rays, (uint32_t)n, ray_flags, nullptr, (uint32_t*)occ, nullptr, nullptr, 0, 0, counter, nullptr, nullptr, 1u, 0u,
        AnyHitCfg{nullptr, 0u, 0u, AH_NONE}, 0u);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int trace_stats(b200rt_context ctx, cudaStream_t s, b200rt_traversable handle, b200rt_deviceptr rays, uint64_t n, uint64_t* nodes,
                uint64_t* tris)
{
    B2_REQUIRE(ctx, handle && rays && nodes && tris && n < (1ull << 32), "bad argument");
    DeviceGuard guard(ctx->device);
    unsigned int* counter = nullptr;
    int rc = next_counter(ctx, s, &counter);
    if (rc) return rc;
    unsigned long long* d_stats = (unsigned long long*)ctx->ws.ptr;
    ExtHit* scratch = nullptr;
    B2_CUDA(ctx, cudaMalloc(&scratch, sizeof(ExtHit) * std::max<uint64_t>(n, 1)));
    {
        const cudaError_t e0 = cudaMemsetAsync(d_stats, 0, 16, s);
        if (e0 != cudaSuccess) { cudaFree(scratch); return set_error(ctx, B200RT_ERROR_CUDA_ERROR, "trace_stats: %s", cudaGetErrorString(e0)); }
    }
    if (n) {
        trace_rays_kernel<0, true, false><<<persistent_grid_rays<0, true>(ctx, n), COOP_BLOCK, 0, s>>>(
            (const AccelHeader*)handle, (const float4*)rays, (uint32_t)n, 0u, scratch, nullptr, nullptr, nullptr, 0, 0, counter, d_stats, nullptr, 1u, 0u, AnyHitCfg{nullptr, 0u, 0u, AH_NONE}, 0u);
        ctx->launches++;
    }
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(h, d_stats, 16, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(scratch);
    if (e != cudaSuccess) return set_error(ctx, B200RT_ERROR_CUDA_ERROR, "trace_stats: %s", cudaGetErrorString(e));
    *nodes = h[0];
    *tris = h[1];
    return 0;
}

int launch_raycast(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr d_params, const b200rt_shader_binding_table* sbt, unsigned width,
                   unsigned height, b200rt_deviceptr ext)
{
    B2_REQUIRE(ctx, d_params && sbt, "null argument");
    B2_REQUIRE(ctx, sbt->hitgroupRecordBase && sbt->hitgroupRecordCount > 0 && sbt->hitgroupRecordStrideInBytes >= 32 + 64,
               "hit-group records (whitted::HitGroupData) are required");
    DeviceGuard guard(ctx->device);
    const uint64_t n = (uint64_t)width * height;
    B2_REQUIRE(ctx, n < (1ull << 32), "launch too large");
    if (n == 0) return 0;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    // __anyhit__texture_mask (optixRaycasting.cu:89-102) needs the material and the texture coordinates of whitted::HitGroupData (a
    // record too short to hold them runs without any-hit programs); whether the traversable holds geometry that runs any-hit programs
    // at all is decided on the device (trace_rays_kernel), so the launch never waits for anything
    const bool full_records = sbt->hitgroupRecordStrideInBytes >= 32 + 352;
    const uint32_t dual = full_records ? 1u : 0u;
    // B200RT_RAYCAST_DRIVER=coop puts the launch back on the persistent cooperative driver (A/B runs; buffers of incoherent rays);
    // B200RT_RAYCAST_DRIVER=coop_ah only its any-hit half
    static const int coop = [] { const char* e = getenv("B200RT_RAYCAST_DRIVER"); return !e ? 0 : !strcmp(e, "coop") ? 3 : !strcmp(e, "coop_ah") ? 2 : 0; }();
    unsigned int* counter = nullptr;   // work-item cursor: the persistent kernels only
    if (coop) {
        const int rc = next_counter(ctx, s, &counter);
        if (rc) return rc;
    }
    const bool tiled = height >= TILE_H && div_up(height, CTA_TILE_H) <= 65535u && B200RT_RAYCAST_TILED;
    const dim3 g_tiled(div_up(width, CTA_TILE_W), div_up(height, CTA_TILE_H));
    const char* hg = (const char*)sbt->hitgroupRecordBase;
    const uint32_t hg_stride = sbt->hitgroupRecordStrideInBytes, hg_count = sbt->hitgroupRecordCount;
    const AnyHitCfg ah{hg, hg_stride, hg_count, AH_TEXTURE_MASK};
    // the plain instantiation, then (records long enough for the any-hit program) the any-hit one: each returns at once unless the
    // traversable is its kind (trace_rays_kernel)
    if (coop & 1)
        trace_rays_kernel<2, false, false><<<persistent_grid_rays<2, false, false>(ctx, n), COOP_BLOCK, 0, s>>>(
            nullptr, nullptr, (uint32_t)n, 0u, (ExtHit*)ext, nullptr, (const RaycastParamsDev*)d_params, hg, hg_stride, hg_count, counter, nullptr, nullptr, 1u, 0u,
            AnyHitCfg{nullptr, 0u, 0u, AH_NONE}, dual);
    else if (tiled)
        B2_CUDA(ctx, launch_chain(raycast_simple_kernel<true, false>, g_tiled, dim3(TILE_CTA_THREADS), s, (const RaycastParamsDev*)d_params, (uint32_t)n, width, height,
                                  (ExtHit*)ext, hg, hg_stride, hg_count, dual));
    else
        B2_CUDA(ctx, launch_chain(raycast_simple_kernel<false, false>, dim3(div_up(n, 128)), dim3(128), s, (const RaycastParamsDev*)d_params, (uint32_t)n, width, height,
                                  (ExtHit*)ext, hg, hg_stride, hg_count, dual));
    if (full_records) {
        B2_LAUNCH_CHECK(ctx);
        // CTAs of the any-hit instantiation: B200RT_RAYCAST_AH_CTAS per SM, each walking its share of the tiles.  Measured on the Duck
        // (7 865 tiles; opaque scene, where this launch does nothing / MASK scene, where it does everything): one CTA per tile 0.203 /
        // 0.263 ms, 32 per SM 0.200 / 0.253, 24 per SM 0.200 / 0.254, 16 per SM 0.199 / 0.257, 8 per SM 0.197 / 0.277, 6 per SM (one
        // resident wave) 0.197 / 0.287.
        static const unsigned ah_per_sm = [] { const char* e = getenv("B200RT_RAYCAST_AH_CTAS"); return e && atoi(e) > 0 ? (unsigned)atoi(e) : 24u; }();
        const unsigned ah_resident = (unsigned)ctx->sm_count * ah_per_sm;
        if (coop & 2)
            trace_rays_kernel<2, false, true><<<persistent_grid_rays<2, false, true>(ctx, n), COOP_BLOCK, 0, s>>>(
                nullptr, nullptr, (uint32_t)n, 0u, (ExtHit*)ext, nullptr, (const RaycastParamsDev*)d_params, hg, hg_stride, hg_count, counter, nullptr, nullptr, 1u, 0u,
                ah, dual);
        else if (tiled)
            B2_CUDA(ctx, launch_chain(raycast_simple_kernel<true, true>, dim3(std::min(g_tiled.x * g_tiled.y, ah_resident)), dim3(TILE_CTA_THREADS), s,
                                      (const RaycastParamsDev*)d_params, (uint32_t)n, width, height, (ExtHit*)ext, hg, hg_stride, hg_count, dual));
        else
            B2_CUDA(ctx, launch_chain(raycast_simple_kernel<false, true>, dim3(std::min(div_up(n, 128), ah_resident)), dim3(128), s,
                                      (const RaycastParamsDev*)d_params, (uint32_t)n, width, height, (ExtHit*)ext, hg, hg_stride, hg_count, dual));
    }
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

// Internal ray-buffer query used by the multi-stage launches (playground.cu): kind 0 = closest hit -> ExtHit records, 1 = any hit ->
// u32 flags.  The ray count may live on the device (n_dev * n_mult, capped by n_max).  Caller holds ctx->mu.
int trace_buffer(b200rt_context ctx, cudaStream_t s, b200rt_traversable handle, b200rt_deviceptr rays, uint64_t n_max, const unsigned int* n_dev,
                 unsigned n_mult, int kind, unsigned ray_flags, b200rt_deviceptr out, unsigned flag_period, b200rt_deviceptr sbt_out,
                 b200rt_deviceptr handle_dev, const b200rt_shader_binding_table* ah_sbt)
{
    B2_REQUIRE(ctx, (handle || handle_dev) && rays && out && n_max < (1ull << 32), "bad argument");
    if (n_max == 0) return 0;
    unsigned int* counter = nullptr;
    int rc = next_counter(ctx, s, &counter);
    if (rc) return rc;
    const AnyHitCfg noah{nullptr, 0u, 0u, AH_NONE};
    if (ah_sbt) {
        // the whitted programs (two ray types): kind 0 = radiance rays -> ExtHit, kind 1 = occlusion rays -> fp32 attenuation
        const AnyHitCfg ah{(const char*)ah_sbt->hitgroupRecordBase, ah_sbt->hitgroupRecordStrideInBytes, ah_sbt->hitgroupRecordCount, AH_WHITTED};
        if (kind == 0)
            trace_rays_kernel<0, false, true><<<persistent_grid_rays<0, false, true>(ctx, n_max), COOP_BLOCK, 0, s>>>(
                (const AccelHeader*)handle, (const float4*)rays, (uint32_t)n_max, ray_flags, (ExtHit*)out, (uint32_t*)sbt_out, nullptr, (const char*)handle_dev, 0, 0,
                counter, nullptr, n_dev, n_mult, flag_period, ah, 0u);
        else
            trace_rays_kernel<1, false, true><<<persistent_grid_rays<1, false, true>(ctx, n_max), COOP_BLOCK, 0, s>>>(
                (const AccelHeader*)handle, (const float4*)rays, (uint32_t)n_max, ray_flags, nullptr, (uint32_t*)out, nullptr, (const char*)handle_dev, 0, 0, counter,
                nullptr, n_dev, n_mult, flag_period, ah, 0u);
    } else if (kind == 0)
        trace_rays_kernel<0, false, false><<<persistent_grid_rays<0, false>(ctx, n_max), COOP_BLOCK, 0, s>>>(
            (const AccelHeader*)handle, (const float4*)rays, (uint32_t)n_max, ray_flags, (ExtHit*)out, (uint32_t*)sbt_out, nullptr, (const char*)handle_dev, 0, 0,
            counter, nullptr, n_dev, n_mult, flag_period, noah, 0u);
    else
        trace_rays_kernel<1, false, false><<<persistent_grid_rays<1, false>(ctx, n_max), COOP_BLOCK, 0, s>>>(
            (const AccelHeader*)handle, (const float4*)rays, (uint32_t)n_max, ray_flags, nullptr, (uint32_t*)out, nullptr, (const char*)handle_dev, 0, 0, counter,
            nullptr, n_dev, n_mult, flag_period, noah, 0u);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b200rt
