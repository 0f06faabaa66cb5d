// pathtracer.cu — wavefront restatement of the Cornell path tracers.
//
// Replaces optixLaunch + the device programs of
//   SDK/optixPathTracer/optixPathTracer.cu:249-413   (mode 0: Russian roulette)
//   SDK/optixMultiGPU/optixMultiGPU.cu:214-385       (mode 1: depth cap 3, sticky emitted/radiance)
// with persistent stage kernels over SoA lane state:
//
//   INIT    one lane per launch index (pixel, or sample index in mode 1): pixel seed tea<4>, first
//           camera ray, lane id appended to the active queue
//   loop:   TRACE  per active lane: resolve the pending shadow ray of the previous bounce (any-hit,
//                  adds radiance*attenuation into the running pixel sum), then closest hit of the
//                  extension ray -> (t, prim, sbt)
//           SHADE  per active lane: miss / closest-hit program, NEE sample -> pending shadow ray,
//                  Russian roulette or depth cap, next bounce ray or *path regeneration* (the lane
//                  starts its next sample), finished lanes write accum + sRGB frame; survivors are
//                  appended to the other queue (warp-aggregated atomics)
//
// A lane runs its samples_per_launch samples back to back, so the fp32 summation order of the
// reference's raygen loop (result += emitted; result += radiance*attenuation, sample after
// sample) and its RNG stream order (2 jitter draws from the pixel seed; per bounce: 2 bounce, 2
// light, 1 roulette draw from the path seed) are preserved exactly; with the arithmetic contract of
// rt_math.cuh the accumulated radiance is bit-identical to the scalar oracle.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "accel.h"
#include "internal.h"
#include "trav_coop.cuh"

namespace b200rt {

struct ParallelogramLight { float3 corner, v1, v2, normal, emission; };

struct PTParams {  // SDK/optixPathTracer/optixPathTracer.h:91-107
    unsigned int subframe_index;
    float4* accum_buffer;
    uchar4* frame_buffer;
    unsigned int width, height, samples_per_launch;
    float3 eye, U, V, W;
    ParallelogramLight light;
    uint64_t handle;
};
static_assert(sizeof(PTParams) == 152 && offsetof(PTParams, eye) == 36 && offsetof(PTParams, light) == 84 && offsetof(PTParams, handle) == 144,
              "optixPathTracer Params layout");

struct MGParams {  // SDK/optixMultiGPU/optixMultiGPU.h:46-64
    unsigned int subframe_index;
    int2* sample_index_buffer;
    float4* sample_accum_buffer;
    uchar4* result_buffer;
    unsigned int width, height, samples_per_launch, device_idx;
    float3 eye, U, V, W;
    ParallelogramLight light;
    uint64_t handle;
};
static_assert(sizeof(MGParams) == 168 && offsetof(MGParams, eye) == 48 && offsetof(MGParams, light) == 96 && offsetof(MGParams, handle) == 160,
              "optixMultiGPU Params layout");

struct HitGroupData { float3 emission_color; float3 diffuse_color; const float4* vertices; };  // optixPathTracer.h:121-126

// unified view of the two Params structs, built on the device at the top of every stage kernel
struct Frame {
    unsigned int subframe, width, height, spl, device_idx;
    unsigned int groups, nlaunch;  // sample groups per launch index (1 = the reference's flat summation order) and launch indices
    float3 eye, U, V, W;
    ParallelogramLight light;
    const AccelHeader* handle;
    float4* accum;
    uchar4* frame;
    const int2* sample_index;
};
template <int MODE>
__device__ __forceinline__ Frame load_frame(const void* p, uint32_t groups, uint32_t nlaunch)
{
    Frame f;
    f.groups = groups; f.nlaunch = nlaunch;
    if (MODE == 0) {
        const PTParams* q = (const PTParams*)p;
        f.subframe = q->subframe_index; f.width = q->width; f.height = q->height; f.spl = q->samples_per_launch; f.device_idx = 0;
        f.eye = q->eye; f.U = q->U; f.V = q->V; f.W = q->W; f.light = q->light;
        f.handle = (const AccelHeader*)q->handle; f.accum = q->accum_buffer; f.frame = q->frame_buffer; f.sample_index = nullptr;
    } else {
        const MGParams* q = (const MGParams*)p;
        f.subframe = q->subframe_index; f.width = q->width; f.height = q->height; f.spl = q->samples_per_launch; f.device_idx = q->device_idx;
        f.eye = q->eye; f.U = q->U; f.V = q->V; f.W = q->W; f.light = q->light;
        f.handle = (const AccelHeader*)q->handle; f.accum = q->sample_accum_buffer; f.frame = q->result_buffer; f.sample_index = q->sample_index_buffer;
    }
    return f;
}

// lane flag word (ray_d.w)
constexpr uint32_t LF_DEPTH_MASK = 0xffu;
constexpr uint32_t LF_SAMPLES_SHIFT = 8, LF_SAMPLES_MASK = 0xfffffu;
constexpr uint32_t LF_SHADOW = 1u << 28;      // a shadow ray is pending
constexpr uint32_t LF_NO_EXT = 1u << 29;      // no extension ray: lane only waits for its last shadow ray
constexpr uint32_t LF_COUNT_EMITTED = 1u << 30;

struct Counters {
    unsigned int qcount[2];  // active lanes (SHADE's work list), double buffered like the two below
    unsigned int n_ext[2];   // lanes with an extension ray to trace (dense list ext_list)
    unsigned int n_shd[2];   // lanes with a pending shadow ray (dense list shd_list)
    unsigned int fetch;      // work-item cursor of the persistent trace kernel (zeroed by INIT / SHADE)
    unsigned int iterations; // wavefront iterations run so far (written by the loop-condition kernel of the launch graph)
    unsigned long long radiance_segments, shadow_segments;
    unsigned long long nodes_fetched, tris_tested;
};

struct Lanes {
    float4* ray_o;   // origin, w = path seed
    float4* ray_d;   // direction, w = flag word
    float4* att;     // attenuation, w = pixel seed
    float4* res;     // running pixel sum, w = hit t (-1 = miss)
    float4* shd_o;   // pending shadow ray origin, w = tmax
    float4* shd_d;   // pending shadow ray direction, w = light weight
    float4* pend;    // attenuation the pending radiance is multiplied with (pre-roulette)
    float4* rad;     // mode 1: sticky radiance (xyz)
    float4* emi;     // mode 1: sticky emitted (xyz)
    uint2* hitp;     // prim, sbt
    unsigned int* queue[2];
    float4* partial;         // groups > 1: per-lane sum of the lane's samples, combined in group order by pt_resolve_kernel
    uint32_t groups, nlaunch;
    unsigned int* ext_list;  // TRACE work items: item i < n_ext is the extension ray of lane ext_list[i],
    unsigned int* shd_list;  //                   item n_ext + j the shadow ray of lane shd_list[j]
    unsigned int* key_ext;   // ray_sort: sort key of ext_list[i] / shd_list[j], written with the list entry
    unsigned int* key_shd;
    const unsigned int* items;  // ray_sort: this iteration's rays in sorted order, lane | shadow << 31 (nullptr: the two lists as they are)
    Counters* counters;
};

__device__ __forceinline__ void camera_ray(const Frame& f, int px, int py, uint32_t& pixel_seed, float3& org, float3& dir)
{
    // optixPathTracer.cu:267-274
    const float jx = rnd(pixel_seed), jy = rnd(pixel_seed);
    const float dx = fm(2.0f, fdiv((float)px + jx, (float)f.width), -1.0f);
    const float dy = fm(2.0f, fdiv((float)py + jy, (float)f.height), -1.0f);
    dir = normalize(f3(fm(dy, f.V.x, dx * f.U.x) + f.W.x, fm(dy, f.V.y, dx * f.U.y) + f.W.y, fm(dy, f.V.z, dx * f.U.z) + f.W.z));
    org = f.eye;
}

// Lanes are group-major: lane = group * nlaunch + launch index, so neighbouring lanes are neighbouring pixels.
template <int MODE>
__device__ __forceinline__ bool lane_pixel(const Frame& f, uint32_t lane, int& px, int& py)
{
    const uint32_t idx = f.groups > 1 ? lane % f.nlaunch : lane;
    if (MODE == 0) {
        px = (int)(idx % f.width);
        py = (int)(idx / f.width);
        return true;
    }
    const int2 p = f.sample_index[idx];
    px = p.x; py = p.y;
    return !(px > (int)f.width - 1 || py > (int)f.height - 1);  // optixMultiGPU.cu:221-223
}

// Append `lane` to up to three queues at once, the slots aggregated over the whole CTA (256 threads): one atomic per CTA and queue instead
// of one per warp and queue.  With 33-66 M lanes per iteration the per-warp form sends millions of atomics to three words of one L2 slice, which
// serialises them (B300_MICROARCH.md: ~0.85 cycles per contended atomic): a floor of several milliseconds per launch.  Must be called
// by every thread of the CTA the same number of times (`iter` = the caller's loop counter; the scratch is double-buffered on its parity).
template <int NQ>
__device__ __forceinline__ void queue_push_block(unsigned int* const (&queue)[NQ], unsigned int* const (&count)[NQ], const bool (&active)[NQ],
                                                 uint32_t lane, uint32_t iter, uint32_t (&idx)[NQ])
{
    __shared__ uint32_t sh_base[2][NQ][8];
    const uint32_t lane_id = threadIdx.x & 31u, wid = threadIdx.x >> 5, lt = (1u << lane_id) - 1u, par = iter & 1u;
    uint32_t mask[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        mask[q] = __ballot_sync(0xffffffffu, active[q]);
        if (lane_id == 0) sh_base[par][q][wid] = (uint32_t)__popc(mask[q]);
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        const int q = (int)threadIdx.x;
        uint32_t pre[8], tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { pre[w] = tot; tot += sh_base[par][q][w]; }
        unsigned int* c = count[0];
#pragma unroll
        for (int j = 1; j < NQ; ++j) c = q == j ? count[j] : c;  // no dynamically indexed pointer array (it would live in local memory)
        const uint32_t base = tot ? atomicAdd(c, tot) : 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w) sh_base[par][q][w] = base + pre[w];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        idx[q] = sh_base[par][q][wid] + (uint32_t)__popc(mask[q] & lt);
        if (active[q]) queue[q][idx[q]] = lane;
    }
}

// ---- ray reordering (b200rt_pt_options.ray_sort) ----------------------------------------------------------------------------------
// Bounce rays of a big scene start all over it and go everywhere: a warp's 24 node fetches are 24 cache misses, and the trace kernel is
// bound by their latency (profiles/r01_trace_kernel.md: camera rays trace at 5.4 Grays/s, the average ray at 2.4).  Sorting the rays
// of an iteration by the cell of their origin, then by direction, puts rays that walk the same part of the tree into the same warp.
// Key: Morton code of the origin on a 2^SORT_OBITS grid over the scene bounds (high bits), direction on 2^SORT_DBITS steps per axis (low).
#ifndef B200RT_SORT_OBITS
#define B200RT_SORT_OBITS 6
#endif
#ifndef B200RT_SORT_DBITS
#define B200RT_SORT_DBITS 2
#endif
constexpr int SORT_OBITS = B200RT_SORT_OBITS, SORT_DBITS = B200RT_SORT_DBITS;
constexpr int SORT_PASSES = (3 * (SORT_OBITS + SORT_DBITS) + 7) / 8;
static_assert(3 * (SORT_OBITS + SORT_DBITS) <= 32, "sort key is 32 bits");

__device__ __forceinline__ uint32_t spread3(uint32_t v)  // ...b2 b1 b0 -> ...b2 0 0 b1 0 0 b0 (up to 10 bits)
{
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t ray_sort_key(const AccelHeader* __restrict__ h, float3 o, float3 d)
{
    const float on = (float)(1 << SORT_OBITS), dn = (float)(1 << SORT_DBITS);
    const float lx = h->bounds[0], ly = h->bounds[1], lz = h->bounds[2];
    const float ex = fmaxf(h->bounds[3] - lx, 1e-30f), ey = fmaxf(h->bounds[4] - ly, 1e-30f), ez = fmaxf(h->bounds[5] - lz, 1e-30f);
    const uint32_t qx = (uint32_t)fminf(fmaxf(__fdividef(o.x - lx, ex) * on, 0.0f), on - 1.0f);
    const uint32_t qy = (uint32_t)fminf(fmaxf(__fdividef(o.y - ly, ey) * on, 0.0f), on - 1.0f);
    const uint32_t qz = (uint32_t)fminf(fmaxf(__fdividef(o.z - lz, ez) * on, 0.0f), on - 1.0f);
    const uint32_t dx = (uint32_t)fminf(fmaxf((d.x * 0.5f + 0.5f) * dn, 0.0f), dn - 1.0f);
    const uint32_t dy = (uint32_t)fminf(fmaxf((d.y * 0.5f + 0.5f) * dn, 0.0f), dn - 1.0f);
    const uint32_t dz = (uint32_t)fminf(fmaxf((d.z * 0.5f + 0.5f) * dn, 0.0f), dn - 1.0f);
    const uint32_t okey = spread3(qx) | (spread3(qy) << 1) | (spread3(qz) << 2);
    const uint32_t dkey = spread3(dx) | (spread3(dy) << 1) | (spread3(dz) << 2);
    return (okey << (3 * SORT_DBITS)) | dkey;
}

// (key, lane | shadow << 31) pairs of one iteration's rays, extension rays first
__global__ void __launch_bounds__(256) pt_build_items_kernel(Lanes L, int cur, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t n_ext = L.counters->n_ext[cur], n_shd = L.counters->n_shd[cur];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ext + n_shd; i += gridDim.x * blockDim.x) {
        if (i < n_ext) { keys[i] = L.key_ext[i]; vals[i] = L.ext_list[i]; }
        else { keys[i] = L.key_shd[i - n_ext]; vals[i] = L.shd_list[i - n_ext] | 0x80000000u; }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) pt_init_kernel(const void* __restrict__ params, Lanes L, uint32_t nlanes)
{
    const Frame f = load_frame<MODE>(params, L.groups, L.nlaunch);
    uint32_t iter = 0;
    for (uint32_t base = blockIdx.x * blockDim.x; base < nlanes; base += gridDim.x * blockDim.x, ++iter) {
        const uint32_t lane = base + threadIdx.x;
        bool active = false;
        if (lane < nlanes) {
            int px, py;
            // samples [s0, s1) of the launch index belong to this lane's group
            const uint32_t g = f.groups > 1 ? lane / f.nlaunch : 0u;
            const uint32_t s0 = (uint32_t)(((uint64_t)g * f.spl) / f.groups), s1 = (uint32_t)(((uint64_t)(g + 1u) * f.spl) / f.groups);
            if (f.groups > 1) L.partial[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane_pixel<MODE>(f, lane, px, py) && s1 > s0) {
                uint32_t pixel_seed = tea4((uint32_t)(py * (int)f.width + px), f.subframe);
                for (uint32_t k = 0; k < 2u * s0; ++k) lcg(pixel_seed);  // every earlier sample drew its two jitter numbers from the pixel seed
                float3 org, dir;
                camera_ray(f, px, py, pixel_seed, org, dir);
                const uint32_t flags = (((s1 - s0) - 1u) & LF_SAMPLES_MASK) << LF_SAMPLES_SHIFT | LF_COUNT_EMITTED;
                L.ray_o[lane] = make_float4(org.x, org.y, org.z, __uint_as_float(pixel_seed));
                L.ray_d[lane] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(flags));
                L.att[lane] = make_float4(1.f, 1.f, 1.f, __uint_as_float(pixel_seed));
                L.res[lane] = make_float4(0.f, 0.f, 0.f, -1.f);
                if (MODE == 1) {
                    L.rad[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                    L.emi[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                active = true;
            }
        }
        unsigned int* const qs[2] = {L.queue[0], L.ext_list};
        unsigned int* const cs[2] = {&L.counters->qcount[0], &L.counters->n_ext[0]};
        const bool as[2] = {active, active};
        uint32_t idx[2];
        queue_push_block<2>(qs, cs, as, lane, iter, idx);
    }
}

// ---- TRACE ---------------------------------------------------------------------------------------
// Work items of one iteration come from the two dense lists SHADE (or INIT) wrote: extension rays (closest hit) first,
// then pending shadow rays (any-hit).  Every item is a ray, so a refill never comes back empty-handed.
// The launches take a GAS handle (the reference pipelines are built with OPTIX_TRAVERSABLE_GRAPH_FLAG_ALLOW_SINGLE_GAS,
// optixPathTracer.cpp:702 / optixMultiGPU.cpp:794); an IAS handle is traversed instance by instance all the same.
#ifndef B200RT_PT_SMEM_STACK
#define B200RT_PT_SMEM_STACK 0   // bottom entries of the traversal stack kept in shared memory (trav_coop.cuh: TStack)
#endif
template <int MODE>
struct PTWork {
    static constexpr bool CONTINUES = false;
    static constexpr bool NODE_POLICY = B200RT_NODE_KEEP_MB > 0;
    __device__ __forceinline__ uint64_t node_policy() const { return node_policy_for(f.handle); }
    static constexpr bool ANYHIT = false;  // the Cornell programs have no any-hit (optixPathTracer.cpp:748-767)
    static constexpr int SMEM_STACK = B200RT_PT_SMEM_STACK;
    const Frame& f;
    const Lanes& L;
    uint32_t n_ext;
    uint32_t lane;
    float weight;
    __device__ PTWork(const Frame& f_, const Lanes& L_, uint32_t n_ext_) : f(f_), L(L_), n_ext(n_ext_), lane(0), weight(0.f) {}

    __device__ __forceinline__ bool stream_triangles() const { return f.handle->kind == ACCEL_KIND_GAS && f.handle->node_bytes == NODE8_BYTES; }
    __device__ __forceinline__ bool fetch(uint32_t item, Trav& s, float* my_ray)
    {
        bool ext;
        if (L.items) { const uint32_t e = L.items[item]; lane = e & 0x7fffffffu; ext = (e >> 31) == 0u; }
        else if (item < n_ext) { lane = L.ext_list[item]; ext = true; }
        else { lane = L.shd_list[item - n_ext]; ext = false; }
        if (ext) {
            const float4 ro = L.ray_o[lane], rd = L.ray_d[lane];
            s.best.t = 1e16f;
            if (!trav_begin_handle(s, my_ray, f.handle, f3(ro.x, ro.y, ro.z), f3(rd.x, rd.y, rd.z), 0.01f, 0u, 0u, 0u)) { commit(s, false); return false; }
        } else {
            const float4 so = L.shd_o[lane], sd = L.shd_d[lane];
            weight = sd.w;
            s.best.t = so.w;
            if (!trav_begin_handle(s, my_ray, f.handle, f3(so.x, so.y, so.z), f3(sd.x, sd.y, sd.z), 0.01f, TP_ANY, 0u, 0u)) { commit(s, false); return false; }
        }
        return true;
    }
    __device__ __forceinline__ bool next_instance(Trav& s, float* my_ray)
    {
        if (f.handle->kind == ACCEL_KIND_GAS || any_ray_done(s)) return false;
        float3 o, d;
        if (s.pack & TP_ANY) { const float4 so = L.shd_o[lane], sd = L.shd_d[lane]; o = f3(so.x, so.y, so.z); d = f3(sd.x, sd.y, sd.z); }
        else { const float4 ro = L.ray_o[lane], rd = L.ray_d[lane]; o = f3(ro.x, ro.y, ro.z); d = f3(rd.x, rd.y, rd.z); }
        return trav_begin_handle(s, my_ray, f.handle, o, d, 0.01f, s.pack & (TP_ANY | TP_FOUND_ANY), 0u, s.inst + 1u);
    }
    __device__ __forceinline__ void commit(const Trav& s, bool found)
    {
        float* res = (float*)&L.res[lane];
        if (s.pack & TP_ANY) {
            // the shadow item owns res.xyz (and rad in mode 1); the extension item owns res.w and hitp
            const float4 pa = L.pend[lane];
            const float w = found ? 0.0f : weight;
            if (MODE == 0) {
                // prd.radiance = light.emission * weight; result += prd.radiance * prd.attenuation
                res[0] = fm(f.light.emission.x * w, pa.x, res[0]);
                res[1] = fm(f.light.emission.y * w, pa.y, res[1]);
                res[2] = fm(f.light.emission.z * w, pa.z, res[2]);
            } else {
                // prd->radiance += light.emission * weight (sticky); result += prd.radiance * prd.attenuation
                float4 rad = L.rad[lane];
                rad.x = fm(f.light.emission.x, w, rad.x);
                rad.y = fm(f.light.emission.y, w, rad.y);
                rad.z = fm(f.light.emission.z, w, rad.z);
                res[0] = fm(rad.x, pa.x, res[0]);
                res[1] = fm(rad.y, pa.y, res[1]);
                res[2] = fm(rad.z, pa.z, res[2]);
                // LF_COUNT_EMITTED marks a freshly regenerated camera ray: the shadow ray just resolved belonged
                // to the previous sample's last bounce, and the new sample starts with prd.radiance = 0
                const uint32_t flags = __float_as_uint(L.ray_d[lane].w);
                L.rad[lane] = (flags & LF_COUNT_EMITTED) ? make_float4(0.f, 0.f, 0.f, 0.f) : rad;
            }
        } else {
            res[3] = found ? s.best.t : -1.0f;
            if (found) L.hitp[lane] = make_uint2(s.best.prim, s.best.sbt);
        }
    }
};

template <int MODE, bool STATS>
__global__ void __launch_bounds__(COOP_BLOCK, COOP_MIN_CTAS) pt_trace_kernel(const void* __restrict__ params, Lanes L, int cur)
{
    const Frame f = load_frame<MODE>(params, L.groups, L.nlaunch);
    const uint32_t n_ext = L.counters->n_ext[cur], n_shd = L.counters->n_shd[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // SHADE of this iteration appends to the other buffers
        L.counters->qcount[cur ^ 1] = 0;
        L.counters->n_ext[cur ^ 1] = 0;
        L.counters->n_shd[cur ^ 1] = 0;
        atomicAdd(&L.counters->radiance_segments, (unsigned long long)n_ext);
        atomicAdd(&L.counters->shadow_segments, (unsigned long long)n_shd);
    }
    TravStats st{0, 0};
    PTWork<MODE> work(f, L, n_ext);
    trace_persistent(work, n_ext + n_shd, &L.counters->fetch, STATS ? &st : nullptr);
    if (STATS) {
        for (int off = 16; off; off >>= 1) {
            st.nodes += __shfl_xor_sync(0xffffffffu, st.nodes, off);
            st.tris += __shfl_xor_sync(0xffffffffu, st.tris, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&L.counters->nodes_fetched, (unsigned long long)st.nodes);
            atomicAdd(&L.counters->tris_tested, (unsigned long long)st.tris);
        }
    }
}

// ---- SHADE ---------------------------------------------------------------------------------------
__device__ __forceinline__ float3 device_color(unsigned int idx)  // optixMultiGPU.cu:199-206
{
    return f3(idx == 0 ? 0.05f : 0.0f, idx == 1 ? 0.05f : 0.0f, idx == 2 ? 0.05f : 0.0f);
}

template <int MODE>
__device__ __forceinline__ void finalize_index(const Frame& f, uint32_t idx, int px, int py, float3 result)
{
    // optixPathTracer.cu:308-319 / optixMultiGPU.cu:281-292
    const float spl = (float)f.spl;
    float3 c = f3(fdiv(result.x, spl), fdiv(result.y, spl), fdiv(result.z, spl));
    const uint32_t image_index = (uint32_t)py * f.width + (uint32_t)px;
    const uint32_t accum_index = MODE == 0 ? image_index : idx;
    if (f.subframe > 0) {
        const float a = fdiv(1.0f, (float)(f.subframe + 1u));
        const float4 prev = f.accum[accum_index];
        c = f3(fm(a, c.x - prev.x, prev.x), fm(a, c.y - prev.y, prev.y), fm(a, c.z - prev.z, prev.z));
    }
    f.accum[accum_index] = make_float4(c.x, c.y, c.z, 1.0f);
    if (f.frame) f.frame[image_index] = make_color(MODE == 0 ? c : c + device_color(f.device_idx));
}

template <int MODE>
__device__ __forceinline__ void finalize_lane(const Frame& f, const Lanes& L, uint32_t lane, int px, int py, float3 result)
{
    if (f.groups > 1) L.partial[lane] = make_float4(result.x, result.y, result.z, 0.f);
    else finalize_index<MODE>(f, lane, px, py, result);
}

// groups > 1: the pixel value is ((g0 + g1) + g2) + ... of the per-group sums, then the reference's tail
template <int MODE>
__global__ void __launch_bounds__(256) pt_resolve_kernel(const void* __restrict__ params, Lanes L)
{
    const Frame f = load_frame<MODE>(params, L.groups, L.nlaunch);
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= f.nlaunch) return;
    int px, py;
    if (!lane_pixel<MODE>(f, idx, px, py) || f.spl == 0) return;
    const float4 p0 = L.partial[idx];
    float3 r = f3(p0.x, p0.y, p0.z);
    for (uint32_t g = 1; g < f.groups; ++g) {
        const float4 p = L.partial[(size_t)g * f.nlaunch + idx];
        r = f3(r.x + p.x, r.y + p.y, r.z + p.z);
    }
    finalize_index<MODE>(f, idx, px, py, r);
}

#ifndef B200RT_SHADE_PREFETCH
#define B200RT_SHADE_PREFETCH 1
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#ifndef B200RT_SHADE_HOIST
#define B200RT_SHADE_HOIST 1
#endif
#ifndef B200RT_SHADE_MIN_CTAS
#define B200RT_SHADE_MIN_CTAS 3   // 79 registers, 12 bytes spilled; bench 2 / 3 / 4 CTAs per SM: 2173 / 2186 / 2171 Mrays/s (with the hoisted loads; 2158 before them)
#endif
template <int MODE>
__global__ void __launch_bounds__(256, B200RT_SHADE_MIN_CTAS) pt_shade_kernel(const void* __restrict__ params, Lanes L, int cur, const char* __restrict__ hg_base,
                                                        uint32_t hg_stride, uint32_t hg_count, const char* __restrict__ miss_base)
{
    const Frame f = load_frame<MODE>(params, L.groups, L.nlaunch);
    const uint32_t n = L.counters->qcount[cur];
    const unsigned int* __restrict__ queue = L.queue[cur];
    unsigned int* next_queue = L.queue[cur ^ 1];
    unsigned int* next_count = &L.counters->qcount[cur ^ 1];
    unsigned int* next_ext = &L.counters->n_ext[cur ^ 1];
    unsigned int* next_shd = &L.counters->n_shd[cur ^ 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) L.counters->fetch = 0;  // the next TRACE starts its item cursor at 0
    const float3 bg = MODE == 0 ? xyz(*(const float4*)(miss_base + B200RT_SBT_RECORD_HEADER_SIZE))
                                : f3(((const float*)(miss_base + B200RT_SBT_RECORD_HEADER_SIZE))[0],
                                     ((const float*)(miss_base + B200RT_SBT_RECORD_HEADER_SIZE))[1],
                                     ((const float*)(miss_base + B200RT_SBT_RECORD_HEADER_SIZE))[2]);
    constexpr uint32_t RAY_TYPES = MODE == 0 ? 1u : 2u;  // SBT stride of the radiance trace call
    uint32_t iter = 0;
    // SHADE is a chain of dependent gathers (queue -> lane state -> hit -> vertices) run by 16 warps per SM: ncu shows it waiting on
    // the long scoreboard with 13 % of the issue slots busy.  The queue entry of the NEXT grid-stride iteration is read one iteration
    // ahead and that lane's state lines are pulled into L2 while the current lane is shaded (B200RT_SHADE_PREFETCH >= 1); with >= 2 the
    // hit of the next lane is read ahead too and the vertices of its triangle are prefetched.
    const uint32_t stride = gridDim.x * blockDim.x;
#if B200RT_SHADE_PREFETCH >= 2
    const uint32_t nprims = f.handle->kind == ACCEL_KIND_GAS ? f.handle->num_tris : 0u;  // prefetch guard only
#endif
    uint32_t lane_next = 0;
    if (B200RT_SHADE_PREFETCH) { const uint32_t q0 = blockIdx.x * blockDim.x + threadIdx.x; if (q0 < n) lane_next = queue[q0]; }
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride, ++iter) {
        const uint32_t qi = base + threadIdx.x;
        bool keep = false, want_ext = false, want_shd = false;
        float3 key_eo = f3(0.f, 0.f, 0.f), key_ed = key_eo, key_so = key_eo, key_sd = key_eo;  // rays pushed this iteration (ray_sort keys)
        uint32_t lane = 0;
        if (B200RT_SHADE_PREFETCH) {
            lane = lane_next;
            const uint32_t nqi = qi + stride;
            if (nqi < n) {
                lane_next = queue[nqi];
                prefetch_l2(&L.ray_d[lane_next]); prefetch_l2(&L.res[lane_next]); prefetch_l2(&L.ray_o[lane_next]); prefetch_l2(&L.att[lane_next]);
                prefetch_l2(&L.hitp[lane_next]);
                if (MODE == 1) prefetch_l2(&L.emi[lane_next]);
#if B200RT_SHADE_PREFETCH >= 2
                {
                    // the hit of the next lane (valid when its res.w >= 0; a stale or never-written record is clamped into the buffer)
                    const uint2 hpn = L.hitp[lane_next];
                    uint32_t rec = (hpn.y & TRI_SBT_MASK) * RAY_TYPES;
                    if (rec >= hg_count) rec = hg_count - 1;
                    const HitGroupData* rtn = (const HitGroupData*)(hg_base + (size_t)rec * hg_stride + B200RT_SBT_RECORD_HEADER_SIZE);
                    if (hpn.x < nprims) {
                        prefetch_l2(rtn->vertices + 3 * (size_t)hpn.x);
                        prefetch_l2(rtn->vertices + 3 * (size_t)hpn.x + 2);
                    }
                }
#endif
            }
        }
        if (qi < n) {
            if (!B200RT_SHADE_PREFETCH) lane = queue[qi];
            float4 rd = L.ray_d[lane];
            uint32_t flags = __float_as_uint(rd.w);
            float4 resv = L.res[lane];
#if B200RT_SHADE_HOIST
            // every line of the lane's state is asked for at once (a lane that only waits for its last shadow ray, or missed, reads a few
            // bytes it does not need): one round trip to memory less in the chain queue -> state -> hit -> vertices
            const float4 ro_early = L.ray_o[lane], att_early = L.att[lane];
            const uint2 hp_early = L.hitp[lane];
#endif
            float3 result = f3(resv.x, resv.y, resv.z);
            int px, py;
            lane_pixel<MODE>(f, lane, px, py);
            if (flags & LF_NO_EXT) {
                // the last shadow ray of the lane has been resolved by TRACE: write the pixel
                finalize_lane<MODE>(f, L, lane, px, py, result);
            } else {
#if B200RT_SHADE_HOIST
                const float4 ro = ro_early;
                float4 attv = att_early;
#else
                const float4 ro = L.ray_o[lane];
                float4 attv = L.att[lane];
#endif
                float3 att = f3(attv.x, attv.y, attv.z);
                uint32_t seed = __float_as_uint(ro.w);
                uint32_t pixel_seed = __float_as_uint(attv.w);
                const float3 org = f3(ro.x, ro.y, ro.z), dir = f3(rd.x, rd.y, rd.z);
                uint32_t depth = flags & LF_DEPTH_MASK;
                uint32_t samples_left = (flags >> LF_SAMPLES_SHIFT) & LF_SAMPLES_MASK;
                bool count_emitted = (flags & LF_COUNT_EMITTED) != 0;
                bool has_shadow = false;
                bool done;
                float3 nrg = org, ndir = dir;
                if (resv.w < 0.0f) {
                    // __miss__radiance: radiance = bg, done.  Mode 0 clears emitted; mode 1 leaves the sticky values.
                    if (MODE == 0) {
                        result = f3(fm(bg.x, att.x, result.x), fm(bg.y, att.y, result.y), fm(bg.z, att.z, result.z));
                    } else {
                        const float4 em = L.emi[lane];
                        result = result + f3(em.x, em.y, em.z);
                        L.rad[lane] = make_float4(bg.x, bg.y, bg.z, 0.f);
                        result = f3(fm(bg.x, att.x, result.x), fm(bg.y, att.y, result.y), fm(bg.z, att.z, result.z));
                    }
                    done = true;
                } else {
                    // __closesthit__radiance (optixPathTracer.cu:338-413 / optixMultiGPU.cu:312-385)
#if B200RT_SHADE_HOIST
                    const uint2 hp = hp_early;
#else
                    const uint2 hp = L.hitp[lane];
#endif
                    uint32_t rec_idx = (hp.y & TRI_SBT_MASK) * RAY_TYPES;
                    if (rec_idx >= hg_count) rec_idx = hg_count - 1;
                    const HitGroupData* rt = (const HitGroupData*)(hg_base + (size_t)rec_idx * hg_stride + B200RT_SBT_RECORD_HEADER_SIZE);
                    const float4* vb = rt->vertices + 3 * (size_t)hp.x;
                    const float3 v0 = xyz(__ldg(vb)), v1 = xyz(__ldg(vb + 1)), v2 = xyz(__ldg(vb + 2));
                    const float3 N0 = normalize(cross(v1 - v0, v2 - v0));
                    const float3 N = N0 * copysignf(1.0f, dot(neg(dir), N0));  // faceforward(N0, -dir, N0)
                    const float t = resv.w;
                    const float3 P = f3(fm(t, dir.x, org.x), fm(t, dir.y, org.y), fm(t, dir.z, org.z));
                    const bool emit_now = MODE == 0 ? (depth == 0) : count_emitted;
                    const float3 emitted = emit_now ? rt->emission_color : f3(0.f, 0.f, 0.f);
                    const float z1 = rnd(seed), z2 = rnd(seed);
                    float s, c;
                    det_sincos(6.2831855f * z2, s, c);
                    const float r = fsqrt(z1);
                    float3 w_in = f3(r * c, r * s, 0.0f);
                    w_in.z = fsqrt(fmaxf(0.0f, fm(-w_in.y, w_in.y, fm(-w_in.x, w_in.x, 1.0f))));
                    // Onb (optixPathTracer.cu:47-78)
                    float3 bn;
                    if (fabsf(N.x) > fabsf(N.z)) bn = f3(-N.y, N.x, 0.0f);
                    else bn = f3(0.0f, -N.z, N.y);
                    bn = normalize(bn);
                    const float3 tg = cross(bn, N);
                    ndir = f3(fm(w_in.z, N.x, fm(w_in.y, bn.x, w_in.x * tg.x)), fm(w_in.z, N.y, fm(w_in.y, bn.y, w_in.x * tg.y)),
                              fm(w_in.z, N.z, fm(w_in.y, bn.z, w_in.x * tg.z)));
                    nrg = P;
                    att = att * rt->diffuse_color;
                    count_emitted = false;
                    const float l1 = rnd(seed), l2 = rnd(seed);
                    const ParallelogramLight& lt = f.light;
                    const float3 lp = f3(fm(lt.v2.x, l2, fm(lt.v1.x, l1, lt.corner.x)), fm(lt.v2.y, l2, fm(lt.v1.y, l1, lt.corner.y)),
                                         fm(lt.v2.z, l2, fm(lt.v1.z, l1, lt.corner.z)));
                    const float3 Ld = lp - P;
                    const float Ldist = length(Ld);
                    const float3 Lv = normalize(Ld);
                    const float nDl = dot(N, Lv);
                    const float LnDl = -dot(lt.normal, Lv);
                    result = result + emitted;
                    if (MODE == 1) L.emi[lane] = make_float4(emitted.x, emitted.y, emitted.z, 0.f);
                    if (nDl > 0.0f && LnDl > 0.0f) {
                        const float A = length(cross(lt.v1, lt.v2));
                        const float weight = fdiv((nDl * LnDl) * A, (3.14159265358979323846f * Ldist) * Ldist);
                        key_so = P; key_sd = Lv;
                        L.shd_o[lane] = make_float4(P.x, P.y, P.z, Ldist - 0.01f);
                        L.shd_d[lane] = make_float4(Lv.x, Lv.y, Lv.z, weight);
                        L.pend[lane] = make_float4(att.x, att.y, att.z, 0.f);
                        has_shadow = true;
                    } else if (MODE == 1) {
                        // weight 0: radiance unchanged, but result += radiance * attenuation still happens
                        const float4 rad = L.rad[lane];
                        result = f3(fm(rad.x, att.x, result.x), fm(rad.y, att.y, result.y), fm(rad.z, att.z, result.z));
                    }
                    done = false;
                }
                // raygen loop tail (optixPathTracer.cu:294-303 / optixMultiGPU.cu:271-277)
                bool path_ends;
                if (MODE == 0) {
                    const float p = dot(att, f3(0.30f, 0.59f, 0.11f));
                    path_ends = done || rnd(seed) > p;
                    if (!path_ends) att = f3(fdiv(att.x, p), fdiv(att.y, p), fdiv(att.z, p));
                } else {
                    path_ends = done || depth >= 3u;
                }
                uint32_t nflags;
                bool lane_done = false;
                if (!path_ends) {
                    depth = min(depth + 1u, 255u);
                    nflags = depth | (samples_left << LF_SAMPLES_SHIFT) | (has_shadow ? LF_SHADOW : 0u);
                } else if (samples_left > 0) {
                    // path regeneration: next sample of this launch index
                    camera_ray(f, px, py, pixel_seed, nrg, ndir);
                    seed = pixel_seed;
                    att = f3(1.f, 1.f, 1.f);
                    nflags = ((samples_left - 1u) << LF_SAMPLES_SHIFT) | LF_COUNT_EMITTED | (has_shadow ? LF_SHADOW : 0u);
                    if (MODE == 1) {
                        // prd.emitted / prd.radiance are re-initialised per sample (optixMultiGPU.cu:244-250) — but a pending
                        // shadow ray of the finished path still needs the old radiance: defer the reset through a marker
                        if (!has_shadow) { L.rad[lane] = make_float4(0.f, 0.f, 0.f, 0.f); }
                        L.emi[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                } else if (has_shadow) {
                    nflags = LF_SHADOW | LF_NO_EXT;
                } else {
                    nflags = 0;
                    lane_done = true;
                }
                if (lane_done) {
                    finalize_lane<MODE>(f, L, lane, px, py, result);
                } else {
                    keep = true;
                    want_ext = (nflags & LF_NO_EXT) == 0u;
                    want_shd = (nflags & LF_SHADOW) != 0u;
                    key_eo = nrg; key_ed = ndir;
                    L.ray_o[lane] = make_float4(nrg.x, nrg.y, nrg.z, __uint_as_float(seed));
                    L.ray_d[lane] = make_float4(ndir.x, ndir.y, ndir.z, __uint_as_float(nflags));
                    L.att[lane] = make_float4(att.x, att.y, att.z, __uint_as_float(pixel_seed));
                    L.res[lane] = make_float4(result.x, result.y, result.z, -1.0f);
                }
            }
        }
        unsigned int* const qs[3] = {next_queue, L.ext_list, L.shd_list};
        unsigned int* const cs[3] = {next_count, next_ext, next_shd};
        const bool as[3] = {keep, want_ext, want_shd};
        uint32_t idx[3];
        queue_push_block<3>(qs, cs, as, lane, iter, idx);
        const uint32_t ie = idx[1], is = idx[2];
        if (L.key_ext) {
            if (want_ext) L.key_ext[ie] = ray_sort_key(f.handle, key_eo, key_ed);
            if (want_shd) L.key_shd[is] = ray_sort_key(f.handle, key_so, key_sd);
        }
    }
}

// ---- fillSamples / de-interleave -------------------------------------------------------------------
__device__ __forceinline__ int2 wd_sample_pixel(int width, int num_gpus, int gpu_idx, int sample_idx)
{
    // StaticWorkDistribution::getSamplePixel (SDK/sutil/WorkDistribution.h:60-81), integer arithmetic
    const int strip_w = 8 * num_gpus;
    const int cols = (width + strip_w - 1) / strip_w;
    const int tile = sample_idx >> 5, in_tile = sample_idx & 31;
    const int row = tile / cols, col = tile - row * cols;
    const int rot = (gpu_idx + row % num_gpus) % num_gpus;
    return make_int2(col * strip_w + rot * 8 + (in_tile & 7), row * 4 + (in_tile >> 3));
}

__global__ void __launch_bounds__(256) fill_samples_kernel(int gpu_idx, int num_gpus, int width, int2* __restrict__ out, int num_samples)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < num_samples) out[i] = wd_sample_pixel(width, num_gpus, gpu_idx, i);
}

__global__ void __launch_bounds__(256) deinterleave_kernel(const float4* __restrict__ gathered, int num_gpus, int num_samples, int width,
                                                            int height, float4* __restrict__ accum, uchar4* __restrict__ frame)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)num_gpus * num_samples) return;
    const int gpu = (int)(i / num_samples), s = (int)(i - (long long)gpu * num_samples);
    const int2 p = wd_sample_pixel(width, num_gpus, gpu, s);
    if (p.x >= width || p.y >= height) return;
    const float4 v = gathered[i];
    const size_t idx = (size_t)p.y * width + p.x;
    if (accum) accum[idx] = v;
    if (frame) frame[idx] = make_color(f3(v.x, v.y, v.z));
}

// ---- synthetic tessellated scene (BASELINE.json configs[4]) -----------------------------------------
// Cornell-sized room (5 tessellated walls, open front), a 3x3x3 grid of displaced lat-long spheres and the
// Cornell light quad; vertex positions are closed-form functions of integer grid coordinates so shared
// vertices are bit-identical (watertight).  Materials: 0 white, 1 green, 2 red, 3 light (the Cornell table).
struct SynthLayout {
    uint64_t total, blob_tris_each, wall_tris_each, light_tris;
    uint32_t rows, cols, grid;
};
__host__ __device__ inline SynthLayout synth_layout(uint64_t total)
{
    SynthLayout s;
    s.total = total;
    const uint64_t t = total > 64 ? total - 2 : 0;  // keep >= 2 for the light
    double per_blob = (double)t * 0.9 / 27.0 / 4.0;
    uint32_t rows = (uint32_t)floor(sqrt(per_blob > 0 ? per_blob : 0));
    if (rows < 2) rows = t >= 27 * 16 ? 2 : 0;
    s.rows = rows;
    s.cols = 2 * rows;
    s.blob_tris_each = 2ull * rows * s.cols;
    const uint64_t used = 27 * s.blob_tris_each;
    uint32_t grid = (uint32_t)floor(sqrt((double)(t > used ? t - used : 0) / 10.0));
    s.grid = grid;
    s.wall_tris_each = 2ull * grid * grid;
    s.light_tris = total - used - 5 * s.wall_tris_each;
    return s;
}

// All arithmetic below is in the contract's named operations (fm / fdiv / det_sincos), so the oracle's
// restatement (oracle.cpp: synth_*) produces bit-identical vertices.
__device__ __forceinline__ float3 blob_vertex(int b, uint32_t i, uint32_t j, uint32_t rows, uint32_t cols, uint32_t seed)
{
    const int bx = b % 3, by = (b / 3) % 3, bz = b / 9;
    const float3 ctr = f3(139.0f + 139.0f * (float)bx, 110.0f + 160.0f * (float)by, 140.0f + 140.0f * (float)bz);
    const float rad = 48.0f;
    if (i == 0) return f3(ctr.x, ctr.y + rad, ctr.z);
    if (i == rows) return f3(ctr.x, ctr.y - rad, ctr.z);
    j = j % cols;
    const float theta = fdiv(3.14159265358979f * (float)i, (float)rows);
    const float phi = fdiv(6.28318530717959f * (float)j, (float)cols);
    float st, ct, sp, cp, s1, s2, s3, s4, unused;
    det_sincos(theta, st, ct);
    det_sincos(phi, sp, cp);
    const float ph = 0.37f * (float)((seed + 7u * (uint32_t)b) % 17u);
    det_sincos(fm(5.0f, theta, ph), s1, unused);
    det_sincos(fm(4.0f, phi, ph), s2, unused);
    det_sincos(23.0f * theta, s3, unused);
    det_sincos(17.0f * phi, s4, unused);
    const float disp = fm(st, fm(0.14f * s1, s2, (0.04f * s3) * s4), 1.0f);
    const float rr = rad * disp;
    return f3(fm(rr * st, cp, ctr.x), fm(rr, ct, ctr.y), fm(rr * st, sp, ctr.z));
}

__device__ __forceinline__ float3 wall_vertex(int wall, uint32_t i, uint32_t j, uint32_t grid)
{
    const float X = 556.0f, Y = 548.8f, Z = 559.2f;
    const float u = fdiv((float)i, (float)grid), v = fdiv((float)j, (float)grid);
    switch (wall) {
        case 0: return f3(X * u, 0.0f, Z * v);   // floor
        case 1: return f3(X * u, Y, Z * v);      // ceiling
        case 2: return f3(X * u, Y * v, Z);      // back wall
        case 3: return f3(0.0f, Y * u, Z * v);   // right wall (green)
        default: return f3(X, Y * u, Z * v);     // left wall (red)
    }
}

__global__ void __launch_bounds__(256) synth_mesh_kernel(SynthLayout s, uint32_t seed, float4* __restrict__ verts, uint32_t* __restrict__ mats)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s.total) return;
    float3 a, b, c;
    uint32_t mat = 0;
    const uint64_t blob_total = 27 * s.blob_tris_each;
    const uint64_t wall_total = 5 * s.wall_tris_each;
    if (t < blob_total) {
        const int bi = (int)(t / s.blob_tris_each);
        const uint64_t r = t - (uint64_t)bi * s.blob_tris_each;
        const uint64_t quad = r >> 1;
        const uint32_t i = (uint32_t)(quad / s.cols), j = (uint32_t)(quad % s.cols);
        const float3 p00 = blob_vertex(bi, i, j, s.rows, s.cols, seed), p01 = blob_vertex(bi, i, j + 1, s.rows, s.cols, seed);
        const float3 p10 = blob_vertex(bi, i + 1, j, s.rows, s.cols, seed), p11 = blob_vertex(bi, i + 1, j + 1, s.rows, s.cols, seed);
        if (r & 1) { a = p00; b = p11; c = p01; } else { a = p00; b = p10; c = p11; }
        mat = (bi % 5 == 1) ? 1u : ((bi % 5 == 3) ? 2u : 0u);
    } else if (t < blob_total + wall_total) {
        const uint64_t r0 = t - blob_total;
        const int w = (int)(r0 / s.wall_tris_each);
        const uint64_t r = r0 - (uint64_t)w * s.wall_tris_each;
        const uint64_t quad = r >> 1;
        const uint32_t i = (uint32_t)(quad / s.grid), j = (uint32_t)(quad % s.grid);
        const float3 p00 = wall_vertex(w, i, j, s.grid), p01 = wall_vertex(w, i, j + 1, s.grid);
        const float3 p10 = wall_vertex(w, i + 1, j, s.grid), p11 = wall_vertex(w, i + 1, j + 1, s.grid);
        if (r & 1) { a = p00; b = p11; c = p01; } else { a = p00; b = p10; c = p11; }
        mat = w == 3 ? 1u : (w == 4 ? 2u : 0u);
    } else {
        // light quad (343..213, 548.6, 227..332) as a strip of K quads along x
        const uint64_t r = t - blob_total - wall_total;
        const uint64_t K = s.light_tris / 2;
        mat = 3u;
        if (K == 0 || (r >> 1) >= K) {
            a = b = c = f3(343.0f, 548.6f, 227.0f);  // odd leftover: zero-area triangle (never hit)
        } else {
            const uint64_t q = r >> 1;
            const float x0 = fm(-130.0f, fdiv((float)q, (float)K), 343.0f);
            const float x1 = (q + 1 == K) ? 213.0f : fm(-130.0f, fdiv((float)(q + 1), (float)K), 343.0f);
            const float3 p00 = f3(x0, 548.6f, 227.0f), p01 = f3(x0, 548.6f, 332.0f), p10 = f3(x1, 548.6f, 227.0f), p11 = f3(x1, 548.6f, 332.0f);
            if (r & 1) { a = p00; b = p11; c = p01; } else { a = p00; b = p10; c = p11; }
        }
    }
    verts[3 * t + 0] = make_float4(a.x, a.y, a.z, 0.f);
    verts[3 * t + 1] = make_float4(b.x, b.y, b.z, 0.f);
    verts[3 * t + 2] = make_float4(c.x, c.y, c.z, 0.f);
    mats[t] = mat;
}

// ---- the wavefront loop as a CUDA graph ---------------------------------------------------------------------------------------------
// while (active lanes) { TRACE(0); SHADE(0); TRACE(1); SHADE(1); }  — a conditional WHILE node whose condition the last kernel of the
// body sets from the queue count on the device.  The body holds two iterations so that every kernel node has a fixed `cur`.  The graph
// is instantiated once per (mode, Params address, workspace, launch size, SBT) and relaunched for every subframe.
struct PTAsyncFlags {  // in the context's pinned block: written by the device, read by the host without synchronisation
    unsigned int runaway;                // the loop hit its iteration cap (reported by the next launch)
    unsigned int pad;
    unsigned long long graph_kernels;    // kernels run inside launch graphs so far (b200rt_context_kernel_launches adds them)
};
constexpr size_t PT_ASYNC_FLAGS_OFFSET = 2048;  // inside ctx->pinned (4096 bytes; whitted.cu keeps its flags at 1024)

__global__ void pt_loop_cond_kernel(Counters* __restrict__ c, cudaGraphConditionalHandle h, PTAsyncFlags* __restrict__ flags)
{
    const unsigned int it = c->iterations + 2u;
    c->iterations = it;
    unsigned int more = c->qcount[0] != 0u ? 1u : 0u;  // SHADE(cur = 1) filled queue 0
    if (more && it > 200000u) { flags->runaway = 1u; more = 0u; }
    atomicAdd_system(&flags->graph_kernels, 5ull);
    cudaGraphSetConditional(h, more);
}

struct PTLoopArgs {
    int mode;
    const void* params;
    Lanes ln;
    bool travstats;
    unsigned trace_grid, shade_grid;
    const char* hg_base;
    uint32_t hg_stride, hg_count;
    const char* miss_base;
};
struct PTGraphEntry {
    bool valid = false;
    PTLoopArgs key;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};
struct PTGraphCache { PTGraphEntry e[4]; unsigned next = 0; };

static bool same_loop(const PTLoopArgs& a, const PTLoopArgs& b)
{
    return a.mode == b.mode && a.params == b.params && !memcmp(&a.ln, &b.ln, sizeof(Lanes)) && a.travstats == b.travstats && a.trace_grid == b.trace_grid &&
           a.shade_grid == b.shade_grid && a.hg_base == b.hg_base && a.hg_stride == b.hg_stride && a.hg_count == b.hg_count && a.miss_base == b.miss_base;
}

uint64_t pathtracer_graph_kernels(b200rt_context ctx) { return ((const volatile PTAsyncFlags*)((char*)ctx->pinned + PT_ASYNC_FLAGS_OFFSET))->graph_kernels; }

void pathtracer_release(b200rt_context ctx)
{
    if (!ctx->pt_graphs) return;
    for (PTGraphEntry& e : ctx->pt_graphs->e) {
        if (e.exec) cudaGraphExecDestroy(e.exec);
        if (e.graph) cudaGraphDestroy(e.graph);
    }
    delete ctx->pt_graphs;
    ctx->pt_graphs = nullptr;
}

template <int MODE>
static int launch_pt_loop_graph(b200rt_context ctx, cudaStream_t s, const PTLoopArgs& la)
{
    PTAsyncFlags* flags = (PTAsyncFlags*)((char*)ctx->pinned + PT_ASYNC_FLAGS_OFFSET);
    if (flags->runaway) {
        flags->runaway = 0;
        return set_error(ctx, B200RT_ERROR_LAUNCH_FAILURE, "an earlier path-tracer launch did not terminate (its loop was cut off after 200000 iterations)");
    }
    if (!ctx->pt_graphs) ctx->pt_graphs = new PTGraphCache();
    PTGraphCache& cache = *ctx->pt_graphs;
    PTGraphEntry* ent = nullptr;
    for (PTGraphEntry& e : cache.e)
        if (e.valid && same_loop(e.key, la)) ent = &e;
    if (!ent) {
        ent = &cache.e[cache.next++ % 4u];
        if (ent->exec) { cudaGraphExecDestroy(ent->exec); ent->exec = nullptr; }
        if (ent->graph) { cudaGraphDestroy(ent->graph); ent->graph = nullptr; }
        ent->valid = false;
        // a zeroed Lanes has padding bytes: the key is compared with memcmp, so it is copied whole
        memcpy(&ent->key, &la, sizeof(PTLoopArgs));
        B2_CUDA(ctx, cudaGraphCreate(&ent->graph, 0));
        cudaGraphConditionalHandle cond;
        B2_CUDA(ctx, cudaGraphConditionalHandleCreate(&cond, ent->graph, 1u, cudaGraphCondAssignDefault));
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = cond;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t loop;
        B2_CUDA(ctx, cudaGraphAddNode(&loop, ent->graph, nullptr, 0, &np));
        cudaGraph_t body = np.conditional.phGraph_out[0];
        cudaGraphNode_t prev = nullptr;
        Lanes ln = la.ln;
        const void* params = la.params;
        const char* hg_base = la.hg_base;
        uint32_t hg_stride = la.hg_stride, hg_count = la.hg_count;
        const char* miss_base = la.miss_base;
        for (int cur = 0; cur < 2; ++cur) {
            int c = cur;
            void* targs[3] = {(void*)&params, (void*)&ln, (void*)&c};
            cudaKernelNodeParams kp;
            memset(&kp, 0, sizeof(kp));
            kp.func = la.travstats ? (void*)pt_trace_kernel<MODE, true> : (void*)pt_trace_kernel<MODE, false>;
            kp.gridDim = dim3(la.trace_grid); kp.blockDim = dim3(COOP_BLOCK); kp.kernelParams = targs;
            cudaGraphNode_t nt;
            B2_CUDA(ctx, cudaGraphAddKernelNode(&nt, body, prev ? &prev : nullptr, prev ? 1 : 0, &kp));
            void* sargs[7] = {(void*)&params, (void*)&ln, (void*)&c, (void*)&hg_base, (void*)&hg_stride, (void*)&hg_count, (void*)&miss_base};
            memset(&kp, 0, sizeof(kp));
            kp.func = (void*)pt_shade_kernel<MODE>;
            kp.gridDim = dim3(la.shade_grid); kp.blockDim = dim3(256); kp.kernelParams = sargs;
            cudaGraphNode_t ns;
            B2_CUDA(ctx, cudaGraphAddKernelNode(&ns, body, &nt, 1, &kp));
            prev = ns;
        }
        {
            Counters* cnt = ln.counters;
            void* cargs[3] = {(void*)&cnt, (void*)&cond, (void*)&flags};
            cudaKernelNodeParams kp;
            memset(&kp, 0, sizeof(kp));
            kp.func = (void*)pt_loop_cond_kernel;
            kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = cargs;
            cudaGraphNode_t nc;
            B2_CUDA(ctx, cudaGraphAddKernelNode(&nc, body, &prev, 1, &kp));
        }
        B2_CUDA(ctx, cudaGraphInstantiate(&ent->exec, ent->graph, 0));
        ent->valid = true;
        log_msg(ctx, 4, "pathtracer", "wavefront loop graph instantiated (mode %d, %u lanes, trace grid %u)", MODE, ln.nlaunch * ln.groups, la.trace_grid);
    }
    B2_CUDA(ctx, cudaGraphLaunch(ent->exec, s));
    return 0;
}

// ---- host ------------------------------------------------------------------------------------------
template <int MODE>
static int run_pathtracer(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr d_params, const b200rt_shader_binding_table* sbt, uint32_t nlaunch,
                          const b200rt_pt_options* opt)
{
    // sample groups: every launch index is served by `groups` lanes that each run a contiguous share of its samples (0 / 1 = one
    // lane per launch index, the reference's flat summation order)
    const uint32_t groups = (opt && opt->sample_groups > 1) ? opt->sample_groups : 1u;
    B2_REQUIRE(ctx, (uint64_t)nlaunch * groups < (1ull << 31), "launch too large (%u launch indices x %u sample groups)", nlaunch, groups);
    const uint32_t nlanes = nlaunch * groups;
    // workspace layout
    const size_t L = nlanes;
    size_t off = 16384;  // the first 16 KiB of the workspace hold the fetch counters / stats of the ray-buffer launches (raycast.cu)
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_cnt = take(sizeof(Counters));
    const size_t o_ro = take(16 * L), o_rd = take(16 * L), o_att = take(16 * L), o_res = take(16 * L), o_so = take(16 * L), o_sd = take(16 * L),
                 o_pend = take(16 * L);
    const size_t o_rad = MODE == 1 ? take(16 * L) : 0, o_emi = MODE == 1 ? take(16 * L) : 0;
    const size_t o_hit = take(8 * L), o_q0 = take(4 * L), o_q1 = take(4 * L), o_ext = take(4 * L), o_shd = take(4 * L);
    const size_t o_part = groups > 1 ? take(16 * L) : 0;
    // ray_sort: keys written with the two lists, and the (key, item) ping-pong of the radix sort over both (2 L items at most)
    const bool sorting = opt && opt->ray_sort == 1u;
    const size_t o_kext = sorting ? take(4 * L) : 0, o_kshd = sorting ? take(4 * L) : 0, o_sk0 = sorting ? take(8 * L) : 0, o_sk1 = sorting ? take(8 * L) : 0,
                 o_sv0 = sorting ? take(8 * L) : 0, o_sv1 = sorting ? take(8 * L) : 0, o_hist = sorting ? take(4 * radix_sort_hist_words(2 * L)) : 0,
                 o_scan = sorting ? take(4 * radix_sort_scan_words(2 * L)) : 0;
    int rc = ensure_workspace(ctx, off, s);
    if (rc) return rc;
    char* W = (char*)ctx->ws.ptr;
    Lanes ln;
    ln.counters = (Counters*)(W + o_cnt);
    ln.ray_o = (float4*)(W + o_ro); ln.ray_d = (float4*)(W + o_rd); ln.att = (float4*)(W + o_att); ln.res = (float4*)(W + o_res);
    ln.shd_o = (float4*)(W + o_so); ln.shd_d = (float4*)(W + o_sd); ln.pend = (float4*)(W + o_pend);
    ln.rad = MODE == 1 ? (float4*)(W + o_rad) : nullptr;
    ln.emi = MODE == 1 ? (float4*)(W + o_emi) : nullptr;
    ln.hitp = (uint2*)(W + o_hit);
    ln.queue[0] = (unsigned int*)(W + o_q0);
    ln.queue[1] = (unsigned int*)(W + o_q1);
    ln.ext_list = (unsigned int*)(W + o_ext);
    ln.shd_list = (unsigned int*)(W + o_shd);
    ln.partial = groups > 1 ? (float4*)(W + o_part) : nullptr;
    ln.groups = groups;
    ln.nlaunch = nlaunch;
    ln.key_ext = sorting ? (unsigned int*)(W + o_kext) : nullptr;
    ln.key_shd = sorting ? (unsigned int*)(W + o_kshd) : nullptr;
    ln.items = nullptr;
    uint32_t* sort_keys[2] = {(uint32_t*)(W + o_sk0), (uint32_t*)(W + o_sk1)};
    uint32_t* sort_vals[2] = {(uint32_t*)(W + o_sv0), (uint32_t*)(W + o_sv1)};

    const uint64_t launches0 = ctx->launches;
    ws_acquire(ctx, s);
    B2_CUDA(ctx, cudaMemsetAsync(ln.counters, 0, sizeof(Counters), s));
    const unsigned grid = persistent_grid(ctx, nlanes, 256, 8);
    const void* params = (const void*)d_params;
    pt_init_kernel<MODE><<<grid, 256, 0, s>>>(params, ln, nlanes);
    B2_LAUNCH_CHECK(ctx);
    Counters* h_cnt = (Counters*)((char*)ctx->pinned + 256);
    // persistent trace kernel: as many CTAs as fit on the device (occupancy query), never more than the work needs
    {
        // L1 and shared memory share 256 KB per SM.  The trace kernel needs 8 x 7.7 KB of shared memory; left to itself the driver carved
        // out 135 KB (ncu: launch__shared_mem_config_size), i.e. took ~70 KB of L1 away from the BVH nodes.
        static const int carveout = [] { const char* e = getenv("B200RT_TRACE_CARVEOUT"); return e ? atoi(e) : 28; }();
        if (carveout >= 0) {
            cudaFuncSetAttribute(pt_trace_kernel<MODE, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
            cudaFuncSetAttribute(pt_trace_kernel<MODE, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
        }
    }
    int occ = 0, occ_stats = 0;
    B2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pt_trace_kernel<MODE, false>, COOP_BLOCK, 0));
    B2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_stats, pt_trace_kernel<MODE, true>, COOP_BLOCK, 0));
    const unsigned work_cap = (unsigned)std::max<uint64_t>(1, (2ull * nlanes + COOP_BLOCK - 1) / COOP_BLOCK);
    const unsigned trace_grid = std::min(work_cap, (unsigned)(std::max(occ, 1) * ctx->sm_count));
    const unsigned trace_grid_stats = std::min(work_cap, (unsigned)(std::max(occ_stats, 1) * ctx->sm_count));
    const uint32_t want = (opt && opt->stats) ? opt->collect_stats : 0u;
    const bool timing = (want & B200RT_PT_STATS_TIMING) != 0, travstats = (want & B200RT_PT_STATS_TRAVERSAL) != 0;
    static const bool force_host_loop = [] { const char* e = getenv("B200RT_PT_HOST_LOOP"); return e && atoi(e) != 0; }();
    uint32_t iterations = 0;
    size_t nev = 0;
    if (!timing && !sorting && !force_host_loop) {
        // ---- asynchronous form: the wavefront loop is a CUDA graph with a conditional WHILE node, launched into the caller's stream.
        // Nothing here waits for the device: optixLaunch is asynchronous by contract (SURVEY 8(b) "Threading"; the reference's one-thread
        // loop over devices, optixMultiGPU.cpp:562-594, relies on it), and so is this launch unless statistics are asked for.
        PTLoopArgs la;
        la.mode = MODE; la.params = params; la.ln = ln; la.travstats = travstats;
        la.trace_grid = travstats ? trace_grid_stats : trace_grid; la.shade_grid = grid;
        la.hg_base = (const char*)sbt->hitgroupRecordBase; la.hg_stride = sbt->hitgroupRecordStrideInBytes; la.hg_count = sbt->hitgroupRecordCount;
        la.miss_base = (const char*)sbt->missRecordBase;
        rc = launch_pt_loop_graph<MODE>(ctx, s, la);
        if (rc) return rc;
    } else {
        // ---- host-driven form (per-stage CUDA events, ray sorting): the host reads the queue counts, so this form synchronises
        auto next_event = [&]() -> cudaEvent_t {
            if (nev == ctx->timing_events.size()) {
                cudaEvent_t e = nullptr;
                cudaEventCreate(&e);
                ctx->timing_events.push_back(e);
            }
            return ctx->timing_events[nev++];
        };
        int cur = 0;
        const int CHECK_EVERY = sorting ? 1 : 8;
        for (;;) {
            for (int k = 0; k < CHECK_EVERY; ++k) {
                ln.items = nullptr;
                if (sorting && iterations > 0) {
                    // the counts of this iteration's rays are on the host (read below); camera rays (iteration 0) are coherent as they are
                    const uint32_t n_rays = h_cnt->n_ext[cur] + h_cnt->n_shd[cur];
                    if (n_rays >= 65536u) {
                        pt_build_items_kernel<<<persistent_grid(ctx, n_rays, 256, 8), 256, 0, s>>>(ln, cur, sort_keys[0], sort_vals[0]);
                        B2_LAUNCH_CHECK(ctx);
                        int res = 0;
                        rc = radix_sort_pairs32(ctx, s, sort_keys, sort_vals, n_rays, SORT_PASSES, (uint32_t*)(W + o_hist), (uint32_t*)(W + o_scan), &res);
                        if (rc) return rc;
                        ln.items = sort_vals[res];
                    }
                }
                if (timing) cudaEventRecord(next_event(), s);
                if (travstats) pt_trace_kernel<MODE, true><<<trace_grid_stats, COOP_BLOCK, 0, s>>>(params, ln, cur);
                else pt_trace_kernel<MODE, false><<<trace_grid, COOP_BLOCK, 0, s>>>(params, ln, cur);
                B2_LAUNCH_CHECK(ctx);
                if (timing) cudaEventRecord(next_event(), s);
                pt_shade_kernel<MODE><<<grid, 256, 0, s>>>(params, ln, cur, (const char*)sbt->hitgroupRecordBase, sbt->hitgroupRecordStrideInBytes,
                                                           sbt->hitgroupRecordCount, (const char*)sbt->missRecordBase);
                B2_LAUNCH_CHECK(ctx);
                if (timing) cudaEventRecord(next_event(), s);
                cur ^= 1;
                ++iterations;
            }
            B2_CUDA(ctx, cudaMemcpyAsync(h_cnt, ln.counters, sizeof(Counters), cudaMemcpyDeviceToHost, s));
            B2_CUDA(ctx, cudaStreamSynchronize(s));
            if (h_cnt->qcount[cur] == 0) break;
            if (iterations > 100000) return set_error(ctx, B200RT_ERROR_LAUNCH_FAILURE, "path tracer did not terminate");
        }
    }
    if (groups > 1) {
        pt_resolve_kernel<MODE><<<div_up(nlaunch, 256), 256, 0, s>>>(params, ln);
        B2_LAUNCH_CHECK(ctx);
    }
    ws_release(ctx, s);
    if (want) {
        // statistics are read back: a launch that asks for them returns when it has finished
        B2_CUDA(ctx, cudaMemcpyAsync(h_cnt, ln.counters, sizeof(Counters), cudaMemcpyDeviceToHost, s));
        B2_CUDA(ctx, cudaStreamSynchronize(s));
        if (iterations == 0) iterations = h_cnt->iterations;
        b200rt_pt_stats* st = opt->stats;
        memset(st, 0, sizeof(*st));
        st->radiance_segments = h_cnt->radiance_segments;
        st->shadow_segments = h_cnt->shadow_segments;
        st->iterations = iterations;
        st->kernel_launches = (uint32_t)(ctx->launches - launches0) + (nev == 0 && !sorting && h_cnt->iterations ? 5u * (h_cnt->iterations / 2u) : 0u);
        st->nodes_fetched = h_cnt->nodes_fetched;
        st->tris_tested = h_cnt->tris_tested;
        if (timing) {
            for (size_t i = 0; i + 2 < nev; i += 3) {
                float a = 0.f, b = 0.f;
                cudaEventElapsedTime(&a, ctx->timing_events[i], ctx->timing_events[i + 1]);
                cudaEventElapsedTime(&b, ctx->timing_events[i + 1], ctx->timing_events[i + 2]);
                st->trace_ms += a;
                st->shade_ms += b;
            }
            st->trace_launches = iterations;
        }
    }
    return 0;
}

int launch_pathtracer(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr d_params, const b200rt_shader_binding_table* sbt, unsigned width,
                      unsigned height, const b200rt_pt_options* opt, int multigpu)
{
    B2_REQUIRE(ctx, d_params && sbt, "null argument");
    B2_REQUIRE(ctx, sbt->hitgroupRecordBase && sbt->hitgroupRecordCount > 0 && sbt->hitgroupRecordStrideInBytes >= 64, "hit-group records required");
    B2_REQUIRE(ctx, sbt->missRecordBase && sbt->missRecordCount > 0, "miss record required");
    const uint64_t nl = (uint64_t)width * height;
    B2_REQUIRE(ctx, nl < (1ull << 31), "launch too large");
    if (nl == 0) return 0;
    DeviceGuard guard(ctx->device);
    return multigpu ? run_pathtracer<1>(ctx, s, d_params, sbt, (uint32_t)nl, opt) : run_pathtracer<0>(ctx, s, d_params, sbt, (uint32_t)nl, opt);
}

int fill_samples(b200rt_context ctx, cudaStream_t s, int gpu_idx, int num_gpus, int width, int height, b200rt_deviceptr out, int num_samples)
{
    (void)height;
    B2_REQUIRE(ctx, out && num_gpus > 0 && gpu_idx >= 0 && gpu_idx < num_gpus && width > 0 && num_samples >= 0, "bad argument");
    DeviceGuard guard(ctx->device);
    if (num_samples == 0) return 0;
    fill_samples_kernel<<<div_up(num_samples, 256), 256, 0, s>>>(gpu_idx, num_gpus, width, (int2*)out, num_samples);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int deinterleave(b200rt_context ctx, cudaStream_t s, b200rt_deviceptr gathered, int num_gpus, int num_samples, int width, int height,
                 b200rt_deviceptr accum, b200rt_deviceptr frame)
{
    B2_REQUIRE(ctx, gathered && num_gpus > 0 && num_samples >= 0 && width > 0 && height > 0, "bad argument");
    DeviceGuard guard(ctx->device);
    const long long n = (long long)num_gpus * num_samples;
    if (n == 0) return 0;
    deinterleave_kernel<<<div_up(n, 256), 256, 0, s>>>((const float4*)gathered, num_gpus, num_samples, width, height, (float4*)accum, (uchar4*)frame);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int generate_synthetic_mesh(b200rt_context ctx, cudaStream_t s, uint64_t num_triangles, uint32_t seed, b200rt_deviceptr verts, b200rt_deviceptr mats,
                            float* bounds_out)
{
    B2_REQUIRE(ctx, verts && mats && num_triangles >= 2 && num_triangles < (1ull << 30), "bad argument");
    DeviceGuard guard(ctx->device);
    const SynthLayout lay = synth_layout(num_triangles);
    synth_mesh_kernel<<<div_up(num_triangles, 256), 256, 0, s>>>(lay, seed, (float4*)verts, (uint32_t*)mats);
    B2_LAUNCH_CHECK(ctx);
    if (bounds_out) {
        bounds_out[0] = 0.f; bounds_out[1] = 0.f; bounds_out[2] = 0.f;
        bounds_out[3] = 556.0f; bounds_out[4] = 548.8f; bounds_out[5] = 559.2f;
    }
    return 0;
}

}  // namespace b200rt
