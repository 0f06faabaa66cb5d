// TEST INFRASTRUCTURE — not product code.
//
// Thin C-ABI wrapper that compiles the *reference's own* host-compilable headers IN PLACE
// (-I/root/reference/SDK, nothing is copied into this repo) so the oracle restatement in
// oracle/oracle.cpp can be pinned against what the reference itself computes:
//   SDK/cuda/random.h:30-67            tea<4>, lcg, rnd
//   SDK/sutil/WorkDistribution.h:50-81 StaticWorkDistribution::{numSamples,getSamplePixel}
//   SDK/sutil/Camera.cpp:34-46         Camera::UVWFrame
//   SDK/sutil/vec_math.h               normalize / cross / length (used by the above)
// Built only where /root/reference exists (this container) into oracle/_ref/libref_shim.so by
// oracle/Makefile; tools/make_golden.py calls it to write tests/golden/kat.json, which is what
// the tests on the GPU box read.
#include <cuda_runtime.h>
#include <sutil/vec_math.h>
#include <cuda/random.h>
#include <sutil/WorkDistribution.h>
#include <sutil/Camera.h>

extern "C" {

unsigned ref_tea4(unsigned v0, unsigned v1) { return tea<4>(v0, v1); }

// advances *state, returns the 24-bit lcg value
unsigned ref_lcg(unsigned* state) { return lcg(*state); }

float ref_rnd(unsigned* state) { return rnd(*state); }

int ref_wd_num_samples(int w, int h, int ngpu, int gpu)
{
    StaticWorkDistribution wd;
    wd.setRasterSize(w, h);
    wd.setNumGPUs(ngpu);
    return wd.numSamples(gpu);
}

void ref_wd_sample_pixel(int w, int h, int ngpu, int gpu, int sample, int* xy)
{
    StaticWorkDistribution wd;
    wd.setRasterSize(w, h);
    wd.setNumGPUs(ngpu);
    int2 p = wd.getSamplePixel(gpu, sample);
    xy[0] = p.x;
    xy[1] = p.y;
}

void ref_camera_uvw(const float* eye, const float* lookat, const float* up, float fovy, float aspect, float* uvw)
{
    sutil::Camera cam(make_float3(eye[0], eye[1], eye[2]), make_float3(lookat[0], lookat[1], lookat[2]),
                      make_float3(up[0], up[1], up[2]), fovy, aspect);
    float3 U, V, W;
    cam.UVWFrame(U, V, W);
    uvw[0] = U.x; uvw[1] = U.y; uvw[2] = U.z;
    uvw[3] = V.x; uvw[4] = V.y; uvw[5] = V.z;
    uvw[6] = W.x; uvw[7] = W.y; uvw[8] = W.z;
}

// light normal as the reference computes it: normalize(cross(v1, v2))  (optixPathTracer.cpp:439)
void ref_light_normal(const float* v1, const float* v2, float* n)
{
    float3 r = normalize(cross(make_float3(v1[0], v1[1], v1[2]), make_float3(v2[0], v2[1], v2[2])));
    n[0] = r.x; n[1] = r.y; n[2] = r.z;
}

}  // extern "C"

// ---- struct layouts of the launch ABI, measured on the reference's own headers ------------------------------------
// (SDK/cuda/whitted.h:44-48, SDK/cuda/GeometryData.h:73-80,248-262, SDK/cuda/BufferView.h:32-38, SDK/optixPathTracer/optixPathTracer.h,
//  SDK/optixMultiGPU/optixMultiGPU.h, SDK/optixRaycasting/optixRaycasting.h + optixRaycastingKernels.h)
#include <optix_types.h>
#include <optix_function_table.h>
#include <cuda/whitted.h>
#include <cstring>
namespace pt {
#include <optixPathTracer/optixPathTracer.h>
}
namespace mg {
#include <optixMultiGPU/optixMultiGPU.h>
}
namespace rc {
#include <optixRaycasting/optixRaycastingKernels.h>
#include <optixRaycasting/optixRaycasting.h>
}

extern "C" {

// out[0..] = sizeof(whitted::HitGroupData), sizeof(GeometryData), sizeof(MaterialData), then the byte offsets inside HitGroupData of
// TriangleMesh::indices / positions / normals / texcoords[0] / texcoords[1] / colors (found by writing sentinels through
// GeometryData::setTriangleMesh and scanning), then sizeof(BufferView<float3>), offsetof data/count/byte_stride/elmt_byte_size,
// then offset of material_data inside HitGroupData.
int ref_hitgroup_layout(int* out)
{
    whitted::HitGroupData hg;  // default-constructed: geometry_data.type == UNKNOWN_TYPE, which setTriangleMesh asserts
    GeometryData::TriangleMesh tm = {};
    tm.indices.data = 0x1111111111111111ull;
    tm.positions.data = 0x2222222222222222ull;
    tm.normals.data = 0x3333333333333333ull;
    tm.texcoords[0].data = 0x4444444444444444ull;
    tm.texcoords[1].data = 0x5555555555555555ull;
    tm.colors.data = 0x6666666666666666ull;
    hg.geometry_data.setTriangleMesh(tm);
    int n = 0;
    out[n++] = (int)sizeof(whitted::HitGroupData);
    out[n++] = (int)sizeof(GeometryData);
    out[n++] = (int)sizeof(MaterialData);
    const unsigned char* p = (const unsigned char*)&hg;
    for (unsigned long long s = 1; s <= 6; ++s) {
        const unsigned long long pat = s * 0x1111111111111111ull;
        int found = -1;
        for (size_t o = 0; o + 8 <= sizeof hg; ++o)
            if (!memcmp(p + o, &pat, 8)) { found = (int)o; break; }
        out[n++] = found;
    }
    out[n++] = (int)sizeof(BufferView<float3>);
    out[n++] = (int)offsetof(BufferView<float3>, data);
    out[n++] = (int)offsetof(BufferView<float3>, count);
    out[n++] = (int)offsetof(BufferView<float3>, byte_stride);
    out[n++] = (int)offsetof(BufferView<float3>, elmt_byte_size);
    out[n++] = (int)offsetof(whitted::HitGroupData, material_data);
    out[n++] = (int)hg.geometry_data.type;  // TRIANGLE_MESH
    return n;
}

// sizes / offsets of the Params structs and SBT payloads the launches consume
int ref_params_layout(int* out)
{
    int n = 0;
    out[n++] = (int)sizeof(pt::Params); out[n++] = (int)offsetof(pt::Params, eye); out[n++] = (int)offsetof(pt::Params, light); out[n++] = (int)offsetof(pt::Params, handle);
    out[n++] = (int)sizeof(pt::HitGroupData); out[n++] = (int)offsetof(pt::HitGroupData, diffuse_color); out[n++] = (int)offsetof(pt::HitGroupData, vertices);
    out[n++] = (int)sizeof(pt::MissData);
    out[n++] = (int)sizeof(mg::Params); out[n++] = (int)offsetof(mg::Params, eye); out[n++] = (int)offsetof(mg::Params, light); out[n++] = (int)offsetof(mg::Params, handle);
    out[n++] = (int)offsetof(mg::Params, sample_index_buffer); out[n++] = (int)offsetof(mg::Params, device_idx);
    out[n++] = (int)sizeof(rc::Params); out[n++] = (int)sizeof(rc::Ray); out[n++] = (int)sizeof(rc::Hit);
    out[n++] = (int)sizeof(OptixBuildInput); out[n++] = (int)sizeof(OptixInstance); out[n++] = (int)sizeof(OptixShaderBindingTable);
    out[n++] = (int)sizeof(OptixAccelBuildOptions); out[n++] = (int)sizeof(OptixBuildInputTriangleArray);
    return n;
}


// MaterialData / Texture / Light / whitted::LaunchParams (SDK/cuda/MaterialData.h:34-140, SDK/cuda/Light.h:31-71, SDK/cuda/whitted.h:59-77)
int ref_whitted_layout(int* out)
{
    int n = 0;
    out[n++] = (int)sizeof(MaterialData); out[n++] = (int)offsetof(MaterialData, type); out[n++] = (int)offsetof(MaterialData, normal_tex);
    out[n++] = (int)offsetof(MaterialData, alpha_mode); out[n++] = (int)offsetof(MaterialData, alpha_cutoff);
    out[n++] = (int)offsetof(MaterialData, emissive_factor); out[n++] = (int)offsetof(MaterialData, emissive_tex);
    out[n++] = (int)offsetof(MaterialData, doubleSided); out[n++] = (int)offsetof(MaterialData, pbr);
    out[n++] = (int)offsetof(MaterialData::Pbr, base_color); out[n++] = (int)offsetof(MaterialData::Pbr, metallic);
    out[n++] = (int)offsetof(MaterialData::Pbr, roughness); out[n++] = (int)offsetof(MaterialData::Pbr, base_color_tex);
    out[n++] = (int)offsetof(MaterialData::Pbr, metallic_roughness_tex);
    out[n++] = (int)sizeof(MaterialData::Texture); out[n++] = (int)offsetof(MaterialData::Texture, texcoord); out[n++] = (int)offsetof(MaterialData::Texture, tex);
    out[n++] = (int)offsetof(MaterialData::Texture, texcoord_offset); out[n++] = (int)offsetof(MaterialData::Texture, texcoord_rotation);
    out[n++] = (int)offsetof(MaterialData::Texture, texcoord_scale);
    out[n++] = (int)sizeof(Light); out[n++] = (int)offsetof(Light, type); out[n++] = (int)offsetof(Light, point);
    out[n++] = (int)offsetof(Light::Point, color); out[n++] = (int)offsetof(Light::Point, intensity); out[n++] = (int)offsetof(Light::Point, position);
    out[n++] = (int)offsetof(Light::Point, falloff);
    out[n++] = (int)sizeof(whitted::LaunchParams); out[n++] = (int)offsetof(whitted::LaunchParams, subframe_index);
    out[n++] = (int)offsetof(whitted::LaunchParams, accum_buffer); out[n++] = (int)offsetof(whitted::LaunchParams, frame_buffer);
    out[n++] = (int)offsetof(whitted::LaunchParams, eye); out[n++] = (int)offsetof(whitted::LaunchParams, U); out[n++] = (int)offsetof(whitted::LaunchParams, lights);
    out[n++] = (int)offsetof(whitted::LaunchParams, miss_color); out[n++] = (int)offsetof(whitted::LaunchParams, handle);
    return n;
}

// OptiX host API layouts the function-table shim mirrors (optix_raytracer_b200/csrc/optix_shim.cu: b200rt_optix_shim_layout returns the
// same list): ABI version, sizeof(OptixFunctionTable), OptixDeviceContextOptions, OptixProgramGroupDesc and the offsets of its entry
// names, OptixPipelineCompileOptions, OptixStackSizes, program-group kinds, the RTCORE_VERSION property id.
int ref_optix_api_layout(int* out)
{
    int k = 0;
    out[k++] = OPTIX_ABI_VERSION;
    out[k++] = (int)sizeof(OptixFunctionTable);
    out[k++] = (int)sizeof(OptixDeviceContextOptions);
    out[k++] = (int)sizeof(OptixProgramGroupDesc);
    out[k++] = (int)(offsetof(OptixProgramGroupDesc, raygen) + offsetof(OptixProgramGroupSingleModule, entryFunctionName));
    out[k++] = (int)(offsetof(OptixProgramGroupDesc, hitgroup) + offsetof(OptixProgramGroupHitgroup, entryFunctionNameCH));
    out[k++] = (int)(offsetof(OptixProgramGroupDesc, hitgroup) + offsetof(OptixProgramGroupHitgroup, entryFunctionNameAH));
    out[k++] = (int)(offsetof(OptixProgramGroupDesc, hitgroup) + offsetof(OptixProgramGroupHitgroup, entryFunctionNameIS));
    out[k++] = (int)sizeof(OptixPipelineCompileOptions);
    out[k++] = (int)sizeof(OptixStackSizes);
    out[k++] = (int)OPTIX_PROGRAM_GROUP_KIND_RAYGEN;
    out[k++] = (int)OPTIX_PROGRAM_GROUP_KIND_MISS;
    out[k++] = (int)OPTIX_PROGRAM_GROUP_KIND_HITGROUP;
    out[k++] = (int)OPTIX_DEVICE_PROPERTY_RTCORE_VERSION;
    return k;
}

}  // extern "C"
