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
