// playground.cu — wavefront restatement of the author's playground sample `imgui_test`.
//
// Replaces   optixLaunch(pipeline, stream, d_param, sizeof(Params), &sbt, buf_width, buf_height, 1)
// (SDK/imgui_test/tracer_window.cpp:96-105) and the device programs of SDK/imgui_test/optixTriangle.cu:
//   __raygen__rg (103-150)      samples_per_frame primary rays per pixel from Camera::compute_ray (camera.h:127-144),
//                               payload summed, film set / accumulated, image = make_color(film / dt)
//   __miss__ms (153-171)        payload = direction * 0.5 + 0.5
//   __closesthit__ch (174-268)  interpolated (un-normalised) vertex normal, P = hit + normal * 1e-4, one shadow probe per light
//                               (LightVariant::wi / lumi: light.h:19-40, volumetric_light.h:22-27, directional_light.h:19-24,
//                               point_light.h), DiffuseMaterial::f (diffuse.h:8-12), one cosine-hemisphere occlusion probe that
//                               only asks hit / no hit, constant 0.01 ambient
// The launch consumes the sample's own 128-byte Params (optixTriangle.h:42-108) and the objects it points to (Camera 92 B,
// LightVariant 44 B, DiffuseMaterial 12 B, float3 vertex / normal arrays, int material indices) byte for byte; layouts are
// pinned against the reference headers in tests/golden/kat.json ("playground_layout").
//
// Stages per BATCH of samples of the frame (lane = sample-in-batch * pixels + pixel: every sample of the batch, all pixels at once — a
// frame of 8 samples is 5 launches of 16 M lanes instead of 40 of 2 M; the samples are independent except for the order in which the
// reference adds their payloads into `result`, which ACCUMULATE keeps):
//   RAYGEN   pixel seed tea<4>(pixel, dt) advanced by the 2 lens draws of every earlier sample -> primary Ray buffer
//   TRACE    closest hit over the ray buffer (trav_coop.cuh persistent driver, as b200rt_trace_closest)
//   SHADE    miss -> payload into the frame sum; hit -> normal, P, light directions (consuming the closest-hit seed in the
//            reference's order), the nlights + 1 probe rays appended to a dense probe buffer
//   TRACE    any-hit over the probe buffer (hit / no hit is all optixHitObjectIsHit() is asked for, also for the bounce probe)
//   RESOLVE  per hit lane: sum of the unoccluded light terms + ambient term -> the lane's payload
//   ACCUMULATE  per pixel: frame sum += payloads of the batch in sample order
// then FINISH writes film and image.  Arithmetic follows the contract of rt_math.cuh (named IEEE operations, fma where nvcc
// contracts the reference source), so the film is bit-identical to the scalar oracle (oracle/oracle.cpp: playground_*).
#include <string.h>

#include <algorithm>

#include "accel.h"
#include "internal.h"
#include "rt_math.cuh"
#include "trav_coop.cuh"

namespace b200rt {

struct PGCamera {  // SDK/imgui_test/camera.h:154-171 (DeviceObject<> is an empty base)
    float3 eye, lookat, up;
    bool ortho;
    float fov, fd, aperture, speed;
    float3 u, v, w;
};
static_assert(sizeof(PGCamera) == 92 && offsetof(PGCamera, fov) == 40 && offsetof(PGCamera, aperture) == 48 && offsetof(PGCamera, u) == 56,
              "imgui_test Camera layout");

struct PGLight {  // LightVariant (light.h:42-50): 40-byte union + tag {Point 0, Directional 1, Volumetric 2}
    float a[3];   // position (point, volumetric) / direction (directional)
    float b[3];   // point: lumi xyz | directional: lumi xyz | volumetric: radius, lumi.x, lumi.y
    float c[4];   // point: dark xyz, - | directional: jitter, dark xyz | volumetric: lumi.z, dark xyz
    int tag;
};
static_assert(sizeof(PGLight) == 44, "imgui_test LightVariant layout");

struct PGParams {  // SDK/imgui_test/optixTriangle.h:42-108
    unsigned int image_width, image_height, samples_per_frame;
    const PGCamera* camera;
    unsigned int dt;
    bool dirty;
    uchar4* image;
    float* film;  // float3[]
    float tfactor;
    uint64_t handle;
    const float* normals;   // float3[3 * ntri]
    const float* vertices;  // float3[3 * ntri]
    const int* mat_indices;
    int nmat_indices;
    const PGLight* lights;
    int nlights;
    const float* materials;  // DiffuseMaterial = float3 color
    int nmaterials;
};
static_assert(sizeof(PGParams) == 128 && offsetof(PGParams, camera) == 16 && offsetof(PGParams, dt) == 24 && offsetof(PGParams, dirty) == 28 &&
                  offsetof(PGParams, image) == 32 && offsetof(PGParams, film) == 40 && offsetof(PGParams, handle) == 56 &&
                  offsetof(PGParams, normals) == 64 && offsetof(PGParams, mat_indices) == 80 && offsetof(PGParams, lights) == 96 &&
                  offsetof(PGParams, nlights) == 104 && offsetof(PGParams, materials) == 112 && offsetof(PGParams, nmaterials) == 120,
              "imgui_test Params layout");

struct PGCounters { unsigned int nhit; unsigned int ncand; unsigned int pad[2]; };

// What the host needs to shape a launch (samples per frame, light count) is read from Params once per Params address and kept
// (ctx->pg_*); every launch's RAYGEN compares the live values with them on the device and reports a difference here, in the context's
// pinned block — the launch after that returns an error and reads Params again.  Kernels take the light count from the HOST's value
// (buffer strides) and never read more lights than Params holds now, so a stale value cannot run anything out of bounds.
struct PGAsyncFlags { unsigned int params_changed; };
constexpr size_t PG_ASYNC_FLAGS_OFFSET = 1536;  // inside ctx->pinned (whitted.cu: 1024, pathtracer.cu: 2048)

// lane index within a sample -> pixel: tile order (trav_coop.cuh), so that the candidate rays a traversal warp picks up together are
// 2-D neighbours; every kernel of the frame uses this one mapping
#ifndef B200RT_PG_TILED
#define B200RT_PG_TILED 0   // measured neutral (imgui_test close camera 15.78 vs 15.81 ms): left off
#endif
__device__ __forceinline__ void pg_pixel(uint32_t i, uint32_t width, uint32_t height, uint32_t& ix, uint32_t& iy)
{
#if B200RT_PG_TILED
    tile_order_xy(i, width, height, ix, iy);
#else
    ix = i % width; iy = i / width;
#endif
}

// ---- RAYGEN: Camera::compute_ray (camera.h:127-144) --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pg_raygen_kernel(const PGParams* __restrict__ params, uint32_t width, uint32_t height, uint32_t sample0,
                                                         uint32_t nlanes, float4* __restrict__ rays, uint32_t* __restrict__ cand,
                                                         float4* __restrict__ payload, PGCounters* __restrict__ counters, uint32_t host_spf,
                                                         int host_nl, PGAsyncFlags* __restrict__ flags)
{
    const uint32_t lane = blockIdx.x * blockDim.x + threadIdx.x;
    bool candidate = false;
    float3 org = f3(0.f, 0.f, 0.f), dir = org;
    if (lane == 0 && (params->samples_per_frame != host_spf || params->nlights != host_nl)) flags->params_changed = 1u;
    if (lane < nlanes) {
    const PGParams P = *params;
    const PGCamera cam = *P.camera;
    const uint32_t npix = width * height;
    const uint32_t i = lane % npix, sample = sample0 + lane / npix;
    uint32_t ix, iy;
    pg_pixel(i, width, height, ix, iy);
    uint32_t seed = tea4(ix + width * iy, P.dt);
    float dx = fm(2.0f, fdiv((float)ix, (float)width), -1.0f), dy = fm(2.0f, fdiv((float)iy, (float)height), -1.0f);
    if (cam.ortho) {
        dir = normalize(f3(fm(dy, cam.v.x, dx * cam.u.x) + cam.w.x, fm(dy, cam.v.y, dx * cam.u.y) + cam.w.y, fm(dy, cam.v.z, dx * cam.u.z) + cam.w.z));
        org = f3(fm(dy, cam.v.x, fm(dx, cam.u.x, cam.eye.x)), fm(dy, cam.v.y, fm(dx, cam.u.y, cam.eye.y)), fm(dy, cam.v.z, fm(dx, cam.u.z, cam.eye.z)));
    } else {
        // the raygen seed runs through the samples of the frame: every earlier sample drew its two lens numbers from it
        for (uint32_t k = 0; k < 2u * sample; ++k) lcg(seed);
        const float lx = (rnd(seed) - 0.5f) * cam.aperture, ly = (rnd(seed) - 0.5f) * cam.aperture;
        dx = dx - lx; dy = dy - ly;
        dir = normalize(f3(fm(dy, cam.v.x, dx * cam.u.x) + cam.w.x, fm(dy, cam.v.y, dx * cam.u.y) + cam.w.y, fm(dy, cam.v.z, dx * cam.u.z) + cam.w.z));
        org = f3(fm(ly, cam.v.x, fm(lx, cam.u.x, cam.eye.x)), fm(ly, cam.v.y, fm(lx, cam.u.y, cam.eye.y)), fm(ly, cam.v.z, fm(lx, cam.u.z, cam.eye.z)));
    }
    // a ray that passes the scene bounds by is a miss: __miss__ms on the spot (payload = direction * 0.5 + 0.5); the others are
    // compacted into the ray buffer of the persistent traversal
    candidate = ray_reaches_bounds((const AccelHeader*)P.handle, org, dir, 0.0f, 1e16f);
    if (!candidate) payload[lane] = make_float4(fm(dir.x, 0.5f, 0.5f), fm(dir.y, 0.5f, 0.5f), fm(dir.z, 0.5f, 0.5f), 0.f);
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, candidate);
    if (!mask) return;
    const uint32_t l32 = threadIdx.x & 31u, leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (l32 == leader) base = atomicAdd(&counters->ncand, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!candidate) return;
    const uint32_t j = base + __popc(mask & ((1u << l32) - 1u));
    cand[j] = lane;
    rays[2 * (size_t)j] = make_float4(org.x, org.y, org.z, 0.0f);
    rays[2 * (size_t)j + 1] = make_float4(dir.x, dir.y, dir.z, 1e16f);
}

// LightVariant::wi / lumi
__device__ __forceinline__ float3 pg_light_wi(const PGLight& l, float3 p, uint32_t& seed)
{
    if (l.tag == 0) return f3(l.a[0] - p.x, l.a[1] - p.y, l.a[2] - p.z);
    if (l.tag == 1) {
        const float j = l.c[0];
        const float r0 = rnd(seed), r1 = rnd(seed), r2 = rnd(seed);
        return f3(fm(j, r0, l.a[0]), fm(j, r1, l.a[1]), fm(j, r2, l.a[2]));
    }
    const float r = l.b[0];
    const float r0 = rnd(seed), r1 = rnd(seed), r2 = rnd(seed);
    return f3(fm(r, r0, l.a[0]) - p.x, fm(r, r1, l.a[1]) - p.y, fm(r, r2, l.a[2]) - p.z);
}
__device__ __forceinline__ float3 pg_light_lumi(const PGLight& l)
{
    return l.tag == 2 ? f3(l.b[1], l.b[2], l.c[0]) : f3(l.b[0], l.b[1], l.b[2]);
}

// ---- SHADE: __miss__ms and the first half of __closesthit__ch -------------------------------------------------------------------
__global__ void __launch_bounds__(256) pg_shade_kernel(const PGParams* __restrict__ params, uint32_t width, uint32_t height, uint32_t nlanes,
                                                        const float4* __restrict__ rays, const uint32_t* __restrict__ cand,
                                                        const ExtHit* __restrict__ hits, float4* __restrict__ payload,
                                                        float4* __restrict__ probes, float* __restrict__ ndw, uint2* __restrict__ hitinfo,
                                                        PGCounters* __restrict__ counters, int nl)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;   // candidate ray
    const uint32_t npix = width * height;
    const PGParams P = *params;
    const int nl_live = min(nl, P.nlights);
    bool is_hit = false;
    ExtHit h;
    float4 ro, rd;
    uint32_t lane_idx = 0;
    if (j < min(counters->ncand, nlanes)) {
        lane_idx = cand[j];
        h = hits[j];
        ro = rays[2 * (size_t)j]; rd = rays[2 * (size_t)j + 1];
        is_hit = h.t >= 0.0f;
        // __miss__ms: payload = direction * 0.5 + 0.5
        if (!is_hit) payload[lane_idx] = make_float4(fm(rd.x, 0.5f, 0.5f), fm(rd.y, 0.5f, 0.5f), fm(rd.z, 0.5f, 0.5f), 0.f);
    }
    const uint32_t i = lane_idx % npix;
    // dense slot for the probes of this pixel
    const uint32_t mask = __ballot_sync(0xffffffffu, is_hit);
    uint32_t base = 0;
    const uint32_t lane = threadIdx.x & 31u;
    if (mask && lane == (uint32_t)(__ffs(mask) - 1)) base = atomicAdd(&counters->nhit, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, mask ? __ffs(mask) - 1 : 0);
    if (!is_hit) return;
    const uint32_t k = base + __popc(mask & ((1u << lane) - 1u));
    const uint32_t vo = h.prim * 3u;
    const float* N = P.normals;
    const float b0 = (1.0f - h.b1) - h.b2;
    // normal = b.x * n1 + b.y * n2 + (1 - b.x - b.y) * n0   (not normalised, optixTriangle.cu:191-193)
    const float3 n = f3(fm(b0, N[3 * vo + 0], fm(h.b2, N[3 * (vo + 2) + 0], h.b1 * N[3 * (vo + 1) + 0])),
                        fm(b0, N[3 * vo + 1], fm(h.b2, N[3 * (vo + 2) + 1], h.b1 * N[3 * (vo + 1) + 1])),
                        fm(b0, N[3 * vo + 2], fm(h.b2, N[3 * (vo + 2) + 2], h.b1 * N[3 * (vo + 1) + 2])));
    const float3 Pp = f3(fm(n.x, 0.0001f, fm(h.t, rd.x, ro.x)), fm(n.y, 0.0001f, fm(h.t, rd.y, ro.y)), fm(n.z, 0.0001f, fm(h.t, rd.z, ro.z)));
    uint32_t ix, iy;
    pg_pixel(i, width, height, ix, iy);
    uint32_t seed = tea4(ix + width * iy, P.dt);
    const size_t pb = (size_t)k * (size_t)(nl + 1);
    for (int li = 0; li < nl; ++li) {
        PGLight l;
        if (li < nl_live) l = P.lights[li];
        else { l = PGLight{}; l.a[1] = 1.0f; l.tag = 1; }  // Params changed under the launch (reported by RAYGEN): a harmless stand-in
        const float3 wi = pg_light_wi(l, Pp, seed);
        ndw[pb + li] = dot(n, wi);
        probes[2 * (pb + li)] = make_float4(Pp.x, Pp.y, Pp.z, 0.01f);
        probes[2 * (pb + li) + 1] = make_float4(wi.x, wi.y, wi.z, 1.0f);
    }
    // Onb(normal) + cosine_sample_hemisphere (optixTriangle.cu:45-85,232-241)
    float3 bn;
    if (fabsf(n.x) > fabsf(n.z)) bn = f3(-n.y, n.x, 0.0f);
    else bn = f3(0.0f, -n.z, n.y);
    bn = normalize(bn);
    const float3 tg = cross(bn, n);
    const float u1 = rnd(seed), u2 = rnd(seed);
    float sn, cs;
    det_sincos(6.2831855f * u2, sn, cs);
    const float r = fsqrt(u1);
    float3 w_in = f3(r * cs, r * sn, 0.0f);
    w_in.z = fsqrt(fmaxf(0.0f, fm(-w_in.y, w_in.y, fm(-w_in.x, w_in.x, 1.0f))));
    const float3 out = f3(fm(w_in.z, n.x, fm(w_in.y, bn.x, w_in.x * tg.x)), fm(w_in.z, n.y, fm(w_in.y, bn.y, w_in.x * tg.y)),
                          fm(w_in.z, n.z, fm(w_in.y, bn.z, w_in.x * tg.z)));
    probes[2 * (pb + nl)] = make_float4(Pp.x, Pp.y, Pp.z, 0.01f);
    probes[2 * (pb + nl) + 1] = make_float4(out.x, out.y, out.z, 1e16f);
    hitinfo[k] = make_uint2(lane_idx, h.prim);
}

// ---- RESOLVE: second half of __closesthit__ch ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pg_resolve_kernel(const PGParams* __restrict__ params, const uint32_t* __restrict__ occluded,
                                                          const float* __restrict__ ndw, const uint2* __restrict__ hitinfo,
                                                          float4* __restrict__ payload, const PGCounters* __restrict__ counters, int nl)
{
    const PGParams P = *params;
    const uint32_t nhit = counters->nhit;
    const int nl_live = min(nl, P.nlights);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nhit; k += gridDim.x * blockDim.x) {
        const uint2 hi = hitinfo[k];
        int m = P.mat_indices[hi.y];
        const float3 color = f3(P.materials[3 * m], P.materials[3 * m + 1], P.materials[3 * m + 2]);
        const size_t pb = (size_t)k * (size_t)(nl + 1);
        float3 result = f3(0.f, 0.f, 0.f);
        for (int li = 0; li < nl; ++li) {
            const float nd = ndw[pb + li];
            const bool dark = occluded[pb + li] != 0u || nd < 0.0f;
            const float3 lumi = li < nl_live ? pg_light_lumi(P.lights[li]) : f3(0.f, 0.f, 0.f);
            // mat.f(...) * lights[li].lumi() * ndotwi
            const float3 term = f3((color.x * lumi.x) * nd, (color.y * lumi.y) * nd, (color.z * lumi.z) * nd);
            result = result + (dark ? f3(0.f, 0.f, 0.f) : term);
        }
        const bool bounce_hit = occluded[pb + nl] != 0u;
        const float3 amb = f3(0.01f, 0.01f, 0.01f);
        result = result + (bounce_hit ? amb : amb * color);
        payload[hi.x] = make_float4(result.x, result.y, result.z, 0.f);
    }
}

// ---- ACCUMULATE: result += payload, sample after sample (optixTriangle.cu:121-141) -----------------------------------------------------
__global__ void __launch_bounds__(256) pg_accumulate_kernel(uint32_t npix, uint32_t nbatch, int first_batch, const float4* __restrict__ payload,
                                                             float4* __restrict__ sum)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 s = first_batch ? make_float4(0.f, 0.f, 0.f, 0.f) : sum[i];
    for (uint32_t b = 0; b < nbatch; ++b) {
        const float4 v = payload[(size_t)b * npix + i];
        s.x += v.x; s.y += v.y; s.z += v.z;
    }
    sum[i] = s;
}

// ---- FINISH: tail of __raygen__rg (optixTriangle.cu:143-149) ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pg_finish_kernel(const PGParams* __restrict__ params, uint32_t width, uint32_t height,
                                                         const float4* __restrict__ sum)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width * height) return;
    const PGParams P = *params;
    uint32_t ix, iy;
    pg_pixel(i, width, height, ix, iy);
    const size_t index = (size_t)iy * P.image_width + ix;
    const float4 s = sum[i];
    float3 f = f3(s.x, s.y, s.z);
    if (!P.dirty) f = f3(P.film[3 * index] + f.x, P.film[3 * index + 1] + f.y, P.film[3 * index + 2] + f.z);
    P.film[3 * index] = f.x; P.film[3 * index + 1] = f.y; P.film[3 * index + 2] = f.z;
    const float dt = (float)P.dt;
    if (P.image) P.image[index] = make_color(f3(fdiv(f.x, dt), fdiv(f.y, dt), fdiv(f.z, dt)));
}

// ---- stand-in scene of BASELINE.json configs[3] (the cover.png model is not in the reference repo) ---------------
// 5 x 5 displaced lat-long blobs (pitch 0.25, material gy * 5 + gx) over the reference's own 20 x 20 x 2 floor tessellation
// (triangle_gas.cpp:147-166, material 26).  Unindexed float3 vertices and per-vertex normals (unit: normalize(v - centre)).
struct PGLayout { uint32_t rows, cols; uint64_t blob_each, total; };
__host__ __device__ inline PGLayout pg_layout(uint32_t rows)
{
    PGLayout l;
    l.rows = rows; l.cols = 2 * rows;
    l.blob_each = 2ull * rows * l.cols;
    l.total = 25ull * l.blob_each + 800ull;
    return l;
}
__device__ __forceinline__ float3 pg_blob_vertex(int b, uint32_t i, uint32_t j, uint32_t rows, uint32_t cols, uint32_t seed, float3& nrm)
{
    const int gx = b % 5, gy = b / 5;
    const float3 ctr = f3(0.25f * (float)(gx - 2), 0.1f, 0.25f * (float)(gy - 2));
    const float rad = 0.08f;
    if (i == 0) { nrm = f3(0.f, 1.f, 0.f); return f3(ctr.x, ctr.y + rad, ctr.z); }
    if (i == rows) { nrm = f3(0.f, -1.f, 0.f); return f3(ctr.x, ctr.y - rad, ctr.z); }
    j = j % cols;
    const float theta = fdiv(3.14159265358979f * (float)i, (float)rows);
    const float phi = fdiv(6.28318530717959f * (float)j, (float)cols);
    float st, ct, sp, cp, s1, s2, unused;
    det_sincos(theta, st, ct);
    det_sincos(phi, sp, cp);
    const float ph = 0.37f * (float)((seed + 7u * (uint32_t)b) % 17u);
    det_sincos(fm(3.0f, theta, ph), s1, unused);
    det_sincos(fm(4.0f, phi, ph), s2, unused);
    const float disp = fm(st, (0.15f * s1) * s2, 1.0f);
    const float rr = rad * disp;
    const float3 d = f3((rr * st) * cp, rr * ct, (rr * st) * sp);
    nrm = normalize(d);
    return f3(d.x + ctr.x, d.y + ctr.y, d.z + ctr.z);
}
__global__ void __launch_bounds__(256) pg_scene_kernel(PGLayout L, uint32_t seed, float* __restrict__ verts, float* __restrict__ normals,
                                                        int* __restrict__ mats)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.total) return;
    float3 a, b, c, na, nb, nc;
    int mat;
    if (t < 25 * L.blob_each) {
        const int bi = (int)(t / L.blob_each);
        const uint64_t r = t - (uint64_t)bi * L.blob_each;
        const uint64_t quad = r >> 1;
        const uint32_t i = (uint32_t)(quad / L.cols), j = (uint32_t)(quad % L.cols);
        float3 n00, n01, n10, n11;
        const float3 p00 = pg_blob_vertex(bi, i, j, L.rows, L.cols, seed, n00), p01 = pg_blob_vertex(bi, i, j + 1, L.rows, L.cols, seed, n01);
        const float3 p10 = pg_blob_vertex(bi, i + 1, j, L.rows, L.cols, seed, n10), p11 = pg_blob_vertex(bi, i + 1, j + 1, L.rows, L.cols, seed, n11);
        if (r & 1) { a = p00; b = p11; c = p01; na = n00; nb = n11; nc = n01; }
        else { a = p00; b = p10; c = p11; na = n00; nb = n10; nc = n11; }
        mat = bi;
    } else {
        // triangle_gas.cpp:147-166: i, j in [-10, 10), two triangles per cell, y = floor (0 here)
        const uint32_t r = (uint32_t)(t - 25 * L.blob_each);
        const int cell = (int)(r >> 1), i = cell / 20 - 10, j = cell % 20 - 10;
        const float x0 = (float)i * 0.1f, x1 = (float)(i + 1) * 0.1f, z0 = (float)j * 0.1f, z1 = (float)(j + 1) * 0.1f;
        if (r & 1) { a = f3(x1, 0.f, z0); b = f3(x0, 0.f, z1); c = f3(x1, 0.f, z1); }
        else { a = f3(x0, 0.f, z0); b = f3(x0, 0.f, z1); c = f3(x1, 0.f, z0); }
        na = nb = nc = f3(0.f, 1.f, 0.f);
        mat = 26;
    }
    float* v = verts + 9 * t;
    float* n = normals + 9 * t;
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = b.x; v[4] = b.y; v[5] = b.z; v[6] = c.x; v[7] = c.y; v[8] = c.z;
    n[0] = na.x; n[1] = na.y; n[2] = na.z; n[3] = nb.x; n[4] = nb.y; n[5] = nb.z; n[6] = nc.x; n[7] = nc.y; n[8] = nc.z;
    mats[t] = mat;
}

// ---- host ------------------------------------------------------------------------------------------------------------------------
int launch_playground(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr d_params, unsigned width, unsigned height, const b200rt_pt_options* opt)
{
    B2_REQUIRE(ctx, d_params, "null params");
    const uint64_t npix64 = (uint64_t)width * height;
    B2_REQUIRE(ctx, npix64 < (1ull << 28), "launch too large");
    if (npix64 == 0) return 0;
    DeviceGuard guard(ctx->device);
    // the launch needs nlights / samples_per_frame on the host to size its buffers and count its batches: one 128-byte read-back on
    // the first launch with a given Params address (the reference allocates, copies and frees its device Params every frame,
    // tracer_window.cpp:96-105); later launches are asynchronous and check the live values on the device (PGAsyncFlags)
    PGAsyncFlags* flags = (PGAsyncFlags*)((char*)ctx->pinned + PG_ASYNC_FLAGS_OFFSET);
    if (flags->params_changed) {
        flags->params_changed = 0;
        ctx->pg_params = 0;
        return set_error(ctx, B200RT_ERROR_INVALID_OPERATION,
                         "samples_per_frame or nlights of Params changed since they were first read for this Params address (the last frame was rendered "
                         "with the old values); relaunch");
    }
    if (ctx->pg_params != d_params) {
        PGParams hp0;
        B2_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, (const void*)d_params, sizeof(PGParams), cudaMemcpyDeviceToHost, s));
        B2_CUDA(ctx, cudaStreamSynchronize(s));
        memcpy(&hp0, ctx->pinned, sizeof(PGParams));
        B2_REQUIRE(ctx, hp0.camera && hp0.film && hp0.handle && hp0.normals && hp0.mat_indices && hp0.materials, "Params has null pointers");
        B2_REQUIRE(ctx, hp0.nlights >= 0 && hp0.nlights <= 64 && (hp0.nlights == 0 || hp0.lights), "bad light list");
        B2_REQUIRE(ctx, hp0.dt > 0, "Params::dt must be > 0 (frame_step() before the launch, tracer_window.cpp:93-94)");
        ctx->pg_params = d_params;
        ctx->pg_spf = hp0.samples_per_frame;
        ctx->pg_nlights = hp0.nlights;
    }
    struct { unsigned int samples_per_frame; int nlights; } hp = {ctx->pg_spf, ctx->pg_nlights};
    const b200rt_deviceptr handle_dev = d_params + offsetof(PGParams, handle);  // the traversal reads the live handle from Params
    const uint32_t npix = (uint32_t)npix64, nl = (uint32_t)hp.nlights, np1 = nl + 1;
    // samples per batch: all of the frame's unless that takes more than ~8 GB of lane buffers (32 + 20 + 16 + 8 + 40 per probe bytes a lane)
    const size_t lane_bytes = 32 + sizeof(ExtHit) + 16 + 8 + (size_t)np1 * 40;
    const uint32_t max_batch = (uint32_t)std::max<size_t>(1, std::min<size_t>((8ull << 30) / (lane_bytes * npix), (1ull << 31) / ((size_t)npix * np1)));
    const uint32_t batch = std::max(1u, std::min(hp.samples_per_frame, max_batch));
    const size_t L = (size_t)npix * batch;
    size_t off = 16384;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_cnt = take(sizeof(PGCounters)), o_rays = take(32 * L), o_hits = take(sizeof(ExtHit) * L), o_pay = take(16 * L),
                 o_sum = take(16ull * npix), o_probe = take(32 * L * np1), o_ndw = take(4 * L * np1), o_occ = take(4 * L * np1),
                 o_info = take(8 * L), o_cand = take(4 * L);
    int rc = ensure_workspace(ctx, off, s);
    if (rc) return rc;
    ws_acquire(ctx, s);
    char* W = (char*)ctx->ws.ptr;
    PGCounters* cnt = (PGCounters*)(W + o_cnt);
    float4* rays = (float4*)(W + o_rays);
    ExtHit* hits = (ExtHit*)(W + o_hits);
    float4* payload = (float4*)(W + o_pay);
    float4* sum = (float4*)(W + o_sum);
    float4* probes = (float4*)(W + o_probe);
    float* ndw = (float*)(W + o_ndw);
    uint32_t* occ = (uint32_t*)(W + o_occ);
    uint2* info = (uint2*)(W + o_info);
    uint32_t* cand = (uint32_t*)(W + o_cand);
    const PGParams* dp = (const PGParams*)d_params;
    const unsigned grid = div_up(npix, 256);
    const uint64_t launches0 = ctx->launches;
    uint64_t primary = 0;
    if (hp.samples_per_frame == 0) B2_CUDA(ctx, cudaMemsetAsync(sum, 0, 16ull * npix, s));
    for (uint32_t smp = 0; smp < hp.samples_per_frame; smp += batch) {
        const uint32_t nb = std::min(batch, hp.samples_per_frame - smp);
        const uint32_t nlanes = npix * nb;
        const unsigned lgrid = div_up(nlanes, 256);
        B2_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(PGCounters), s));
        pg_raygen_kernel<<<lgrid, 256, 0, s>>>(dp, width, height, smp, nlanes, rays, cand, payload, cnt, hp.samples_per_frame, hp.nlights, flags);
        B2_LAUNCH_CHECK(ctx);
        rc = trace_buffer(ctx, s, 0, (b200rt_deviceptr)rays, nlanes, &cnt->ncand, 1, 0, B200RT_RAY_VISIBILITY_MASK(255) /* optixTriangle.cu:130 */, (b200rt_deviceptr)hits, 0, 0, handle_dev);
        if (rc) return rc;
        pg_shade_kernel<<<lgrid, 256, 0, s>>>(dp, width, height, nlanes, rays, cand, hits, payload, probes, ndw, info, cnt, hp.nlights);
        B2_LAUNCH_CHECK(ctx);
        // shadow probes: TERMINATE_ON_FIRST_HIT | CULL_DISABLED_ANYHIT (optixTriangle.cu:213-223); the bounce probe carries no flags
        // but is only asked hit / no hit, so both kinds go through one any-hit batch; the last ray of every group of nl + 1 (the
        // bounce probe) ignores CULL_DISABLED_ANYHIT.  With the sample's OPTIX_GEOMETRY_FLAG_NONE build input nothing is culled;
        // a DISABLE_ANYHIT geometry is invisible to the light probes, exactly as in OptiX.
        rc = trace_buffer(ctx, s, 0, (b200rt_deviceptr)probes, (uint64_t)nlanes * np1, &cnt->nhit, np1, 1, 64u /* CULL_DISABLED_ANYHIT */,
                          (b200rt_deviceptr)occ, np1, 0, handle_dev);
        if (rc) return rc;
        pg_resolve_kernel<<<std::min<unsigned>(lgrid, (unsigned)ctx->sm_count * 8u), 256, 0, s>>>(dp, occ, ndw, info, payload, cnt, hp.nlights);
        B2_LAUNCH_CHECK(ctx);
        pg_accumulate_kernel<<<grid, 256, 0, s>>>(npix, nb, smp == 0 ? 1 : 0, payload, sum);
        B2_LAUNCH_CHECK(ctx);
        primary += nlanes;
    }
    pg_finish_kernel<<<grid, 256, 0, s>>>(dp, width, height, sum);
    B2_LAUNCH_CHECK(ctx);
    ws_release(ctx, s);
    if (opt && opt->stats && opt->collect_stats) {
        b200rt_pt_stats* st = opt->stats;
        memset(st, 0, sizeof(*st));
        st->radiance_segments = primary;
        st->iterations = hp.samples_per_frame;
        st->kernel_launches = (uint32_t)(ctx->launches - launches0);
    }
    return 0;
}

int generate_playground_scene(b200rt_context ctx, cudaStream_t s, uint32_t rows, uint32_t seed, b200rt_deviceptr verts, b200rt_deviceptr normals,
                              b200rt_deviceptr mats, uint64_t* num_triangles)
{
    B2_REQUIRE(ctx, rows >= 2 && rows <= 4096 && num_triangles, "bad argument");
    const PGLayout L = pg_layout(rows);
    *num_triangles = L.total;
    if (!verts) return 0;  // size query
    B2_REQUIRE(ctx, normals && mats, "null buffer");
    DeviceGuard guard(ctx->device);
    pg_scene_kernel<<<div_up(L.total, 256), 256, 0, s>>>(L, seed, (float*)verts, (float*)normals, (int*)mats);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b200rt
