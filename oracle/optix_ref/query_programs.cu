// query_programs.cu — TEST INFRASTRUCTURE: OptiX device programs for ray-buffer queries.
//
// The reference's __closesthit__buffer_hit (SDK/optixRaycasting/optixRaycasting.cu:74-86) stores only the
// truncated t and the shading normal; the north star's bit-exactness check needs primitive index, instance
// and barycentrics too ("a 10-line patched closest-hit").  These programs read the same
// 32-byte Ray records (optixRaycastingKernels.h:35-41) and write the 20-byte extended hit record that
// b200rt_trace_closest writes, or a u32 occlusion flag with the ray flags of traceOcclusion
// (SDK/optixPathTracer/optixPathTracer.cu:218-240).  SBT stride 0: every geometry of a GAS uses the hit record at the
// instance's sbtOffset, so one program serves scenes built for any number of SBT records.
// Compiled to PTX with the reference's OptiX headers.
#include <optix.h>

struct QueryParams {
    OptixTraversableHandle handle;
    const float4* rays;     // 2 x float4 per ray: origin|tmin, direction|tmax
    unsigned int* out;      // closest: 5 words per ray {t, prim, inst, b1, b2}; any: 1 word per ray
    unsigned int ray_flags;
    unsigned int any_hit;   // 1: terminate on first hit, write occlusion flag
    unsigned long long n;   // number of rays (the 2-D launch may be padded)
};

extern "C" {
__constant__ QueryParams params;
}

extern "C" __global__ void __raygen__query()
{
    const uint3 idx = optixGetLaunchIndex();
    const uint3 dim = optixGetLaunchDimensions();
    const unsigned long long i = (unsigned long long)idx.y * dim.x + idx.x;
    if (i >= params.n) return;
    const float4 a = params.rays[2 * i], b = params.rays[2 * i + 1];
    unsigned int p0 = 0xbf800000u /* t = -1 */, p1 = 0xffffffffu, p2 = 0xffffffffu, p3 = 0u, p4 = 0u;
    if (params.any_hit) {
        p0 = 0u;
        optixTrace(params.handle, make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), a.w, b.w, 0.0f, OptixVisibilityMask(((params.ray_flags >> 16) ^ 1u) & 0xffu),
                   (params.ray_flags & 0xffffu) | OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT | OPTIX_RAY_FLAG_DISABLE_ANYHIT, 0, 0, 0, p0, p1, p2, p3, p4);
        params.out[i] = p0;
    } else {
        optixTrace(params.handle, make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), a.w, b.w, 0.0f, OptixVisibilityMask(((params.ray_flags >> 16) ^ 1u) & 0xffu),
                   (params.ray_flags & 0xffffu) | OPTIX_RAY_FLAG_DISABLE_ANYHIT, 0, 0, 0, p0, p1, p2, p3, p4);
        unsigned int* o = params.out + 5 * i;
        o[0] = p0; o[1] = p1; o[2] = p2; o[3] = p3; o[4] = p4;
    }
}

extern "C" __global__ void __miss__query() {}

extern "C" __global__ void __closesthit__query()
{
    if (params.any_hit) {
        optixSetPayload_0(1u);
        return;
    }
    const float2 bc = optixGetTriangleBarycentrics();
    optixSetPayload_0(__float_as_uint(optixGetRayTmax()));
    optixSetPayload_1(optixGetPrimitiveIndex());
    optixSetPayload_2(optixGetInstanceIndex());
    optixSetPayload_3(__float_as_uint(bc.x));
    optixSetPayload_4(__float_as_uint(bc.y));
}
