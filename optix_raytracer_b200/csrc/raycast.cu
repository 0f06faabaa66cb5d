// raycast.cu — the optixRaycasting path: ray-buffer generation, closest-hit queries over a
// traversable, the sample's closest-hit/miss programs, and the hit shading kernel.
//
// Reference: SDK/optixRaycasting/optixRaycastingKernels.cu:42-115 (createRaysOrtho / translateRays /
// shadeHits), SDK/optixRaycasting/optixRaycasting.cu:45-86 (raygen/miss/closest-hit),
// SDK/cuda/LocalGeometry.h:59-176 (shading normal), SDK/optixRaycasting/optixRaycasting.cpp:289-317 (launch).
// Rays are 32-byte AoS records and hits 16-byte records exactly as in the reference, so a warp reads
// 1 KiB and writes 512 B contiguous per request; each ray record is fetched as two 16-byte loads.
#include "accel.h"
#include "internal.h"
#include "traverse.cuh"

namespace b200rt {

struct RayRec { float ox, oy, oz, tmin, dx, dy, dz, tmax; };  // optixRaycastingKernels.h:35-41
struct RaycastParams { uint64_t handle; const RayRec* rays; float4* hits; };  // optixRaycasting.h:41-46

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

// ---- generic queries ------------------------------------------------------------------------------
template <bool ANY, bool STATS>
__global__ void __launch_bounds__(256) trace_rays_kernel(const AccelHeader* __restrict__ handle, const float4* __restrict__ rays, uint64_t n,
                                                          uint32_t ray_flags, ExtHit* __restrict__ ext, uint32_t* __restrict__ occluded,
                                                          unsigned long long* __restrict__ stats)
{
    TravStats st{0, 0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
        RayHit hit;
        const bool found = trace_handle<ANY, STATS>(handle, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, ray_flags, hit, &st);
        if (ANY) occluded[i] = found ? 1u : 0u;
        else {
            ExtHit e;
            if (found) { e.t = hit.t; e.prim = hit.prim; e.inst = hit.inst; e.b1 = hit.b1; e.b2 = hit.b2; }
            else { e.t = -1.0f; e.prim = 0xffffffffu; e.inst = 0xffffffffu; e.b1 = 0.f; e.b2 = 0.f; }
            ext[i] = e;
        }
    }
    if (STATS) {
        atomicAdd(&stats[0], (unsigned long long)st.nodes);
        atomicAdd(&stats[1], (unsigned long long)st.tris);
    }
}

// ---- the optixRaycasting launch: raygen + closest-hit + miss fused per ray -------------------------
struct BufView { uint64_t data; uint32_t count; uint16_t byte_stride; uint16_t elmt; };  // SDK/cuda/BufferView.h:32-38

__global__ void __launch_bounds__(256) raycast_launch_kernel(const RaycastParams* __restrict__ params, const char* __restrict__ hg_base,
                                                              uint32_t hg_stride, uint32_t hg_count, uint64_t n, ExtHit* __restrict__ ext)
{
    const RaycastParams P = *params;
    const AccelHeader* handle = (const AccelHeader*)P.handle;
    const float4* rays = (const float4*)P.rays;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
        RayHit hit;
        const bool found = trace_handle<false, false>(handle, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, 0u, hit, nullptr);
        float4 out;
        if (!found) {
            out = make_float4(-1.0f, 1.0f, 0.0f, 0.0f);  // __miss__buffer_miss
        } else {
            // __closesthit__buffer_hit: SBT record = instance.sbtOffset + GAS-local index (ray type 0, stride 1)
            uint32_t sbt = hit.sbt & TRI_SBT_MASK;
            const InstanceRecord* ir = nullptr;
            if (handle->kind == ACCEL_KIND_IAS) {
                ir = (const InstanceRecord*)((const char*)handle + handle->inst_off) + hit.inst;
                sbt += ir->sbt_offset;
            }
            if (sbt >= hg_count) sbt = hg_count - 1;
            const char* rec = hg_base + (size_t)sbt * hg_stride + B200RT_SBT_RECORD_HEADER_SIZE;
            // whitted::HitGroupData -> GeometryData{type@0, TriangleMesh@8{indices, positions, normals, ...}}
            const BufView vi = *(const BufView*)(rec + 8), vp = *(const BufView*)(rec + 24), vn = *(const BufView*)(rec + 40);
            uint32_t i0, i1, i2;
            if (vi.elmt == 4) { const uint32_t* ip = (const uint32_t*)vi.data + 3 * (size_t)hit.prim; i0 = ip[0]; i1 = ip[1]; i2 = ip[2]; }
            else if (vi.elmt == 2) { const uint16_t* ip = (const uint16_t*)vi.data + 3 * (size_t)hit.prim; i0 = ip[0]; i1 = ip[1]; i2 = ip[2]; }
            else { i0 = 3 * hit.prim; i1 = i0 + 1; i2 = i0 + 2; }
            float3 N;
            if (vn.data) {
                const uint32_t st = vn.byte_stride ? vn.byte_stride : 12u;
                const float* n0 = (const float*)(vn.data + (size_t)i0 * st);
                const float* n1 = (const float*)(vn.data + (size_t)i1 * st);
                const float* n2 = (const float*)(vn.data + (size_t)i2 * st);
                const float b0 = (1.0f - hit.b1) - hit.b2;
                N = f3(fm(hit.b2, n2[0], fm(hit.b1, n1[0], b0 * n0[0])), fm(hit.b2, n2[1], fm(hit.b1, n1[1], b0 * n0[1])),
                       fm(hit.b2, n2[2], fm(hit.b1, n1[2], b0 * n0[2])));
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
            out = make_float4((float)(unsigned int)hit.t, N.x, N.y, N.z);
        }
        P.hits[i] = out;
        if (ext) {
            ExtHit e;
            if (found) { e.t = hit.t; e.prim = hit.prim; e.inst = hit.inst; e.b1 = hit.b1; e.b2 = hit.b2; }
            else { e.t = -1.0f; e.prim = 0xffffffffu; e.inst = 0xffffffffu; e.b1 = 0.f; e.b2 = 0.f; }
            ext[i] = e;
        }
    }
}

static unsigned grid_for(b200rt_context ctx, uint64_t n, int block, int ctas_per_sm)
{
    const uint64_t need = (n + block - 1) / block;
    const uint64_t cap = (uint64_t)ctx->sm_count * ctas_per_sm;
    return (unsigned)std::max<uint64_t>(1, std::min(need, cap));
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
    B2_REQUIRE(ctx, handle && (n == 0 || (rays && ext)), "null argument");
    DeviceGuard guard(ctx->device);
    if (n == 0) return 0;
    trace_rays_kernel<false, false><<<grid_for(ctx, n, 256, 8), 256, 0, s>>>((const AccelHeader*)handle, (const float4*)rays, n, ray_flags,
                                                                             (ExtHit*)ext, nullptr, nullptr);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int trace_any(b200rt_context ctx, cudaStream_t s, b200rt_traversable handle, b200rt_deviceptr rays, uint64_t n, unsigned ray_flags,
              b200rt_deviceptr occ)
{
    B2_REQUIRE(ctx, handle && (n == 0 || (rays && occ)), "null argument");
    DeviceGuard guard(ctx->device);
    if (n == 0) return 0;
    trace_rays_kernel<true, false><<<grid_for(ctx, n, 256, 8), 256, 0, s>>>((const AccelHeader*)handle, (const float4*)rays, n, ray_flags,
                                                                            nullptr, (uint32_t*)occ, nullptr);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int trace_stats(b200rt_context ctx, cudaStream_t s, b200rt_traversable handle, b200rt_deviceptr rays, uint64_t n, uint64_t* nodes,
                uint64_t* tris)
{
    B2_REQUIRE(ctx, handle && rays && nodes && tris, "null argument");
    DeviceGuard guard(ctx->device);
    int rc = ensure_workspace(ctx, 1 << 20, s);
    if (rc) return rc;
    unsigned long long* d_stats = (unsigned long long*)ctx->ws.ptr;
    ExtHit* scratch = nullptr;
    B2_CUDA(ctx, cudaMalloc(&scratch, sizeof(ExtHit) * std::max<uint64_t>(n, 1)));
    B2_CUDA(ctx, cudaMemsetAsync(d_stats, 0, 16, s));
    if (n) {
        trace_rays_kernel<false, true><<<grid_for(ctx, n, 256, 8), 256, 0, s>>>((const AccelHeader*)handle, (const float4*)rays, n, 0u,
                                                                                scratch, nullptr, d_stats);
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
    B2_REQUIRE(ctx, sbt->hitgroupRecordBase && sbt->hitgroupRecordCount > 0 && sbt->hitgroupRecordStrideInBytes >= 32 + 56,
               "hit-group records (whitted::HitGroupData) are required");
    DeviceGuard guard(ctx->device);
    const uint64_t n = (uint64_t)width * height;
    if (n == 0) return 0;
    raycast_launch_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, s>>>((const RaycastParams*)d_params, (const char*)sbt->hitgroupRecordBase,
                                                                   sbt->hitgroupRecordStrideInBytes, sbt->hitgroupRecordCount, n, (ExtHit*)ext);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b200rt
