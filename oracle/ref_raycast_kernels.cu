// TEST INFRASTRUCTURE — the reference's own plain-CUDA helper kernels of optixRaycasting (createRaysOrthoOnDevice, translateRaysOnDevice,
// shadeHitsOnDevice: SDK/optixRaycasting/optixRaycastingKernels.cu:42-115), compiled where they lie with the reference's nvcc flags
// (--use_fast_math, SDK/CMakeLists.txt:267) behind three extern "C" entry points, as the checker of b200rt_create_rays_ortho /
// b200rt_translate_rays / b200rt_shade_hits.  Built by oracle/Makefile into oracle/_ref/librefraycast.so only
// where /root/reference exists; nothing is copied.  The reference launches on the default stream; the wrappers synchronise.
#include <optixRaycasting/optixRaycastingKernels.cu>

extern "C" int ref_create_rays_ortho(void* rays, int width, int height, const float* bbmin, const float* bbmax, float padding)
{
    createRaysOrthoOnDevice((Ray*)rays, width, height, make_float3(bbmin[0], bbmin[1], bbmin[2]), make_float3(bbmax[0], bbmax[1], bbmax[2]), padding);
    return (int)cudaDeviceSynchronize();
}
extern "C" int ref_translate_rays(void* rays, int count, const float* offset)
{
    translateRaysOnDevice((Ray*)rays, count, make_float3(offset[0], offset[1], offset[2]));
    return (int)cudaDeviceSynchronize();
}
extern "C" int ref_shade_hits(void* image_float3, int count, const void* hits)
{
    shadeHitsOnDevice((float3*)image_float3, count, (const Hit*)hits);
    return (int)cudaDeviceSynchronize();
}
