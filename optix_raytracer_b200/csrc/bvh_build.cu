// bvh_build.cu — acceleration-structure build on the GPU (sm_100a), replacing the closed
// optixAccelComputeMemoryUsage / optixAccelBuild / optixAccelCompact the reference calls at
// SDK/optixPathTracer/optixPathTracer.cpp:627-684, SDK/sutil/Scene.cpp:970,1043-1054,1106,1195 and
// SDK/imgui_test/triangle_gas.cpp:213-234.
//
// Pipeline (all kernels hand written, no CUB/Thrust):
//   1 gather      build inputs (strided float3 vertices, optional u16/u32 indices, optional 3x4
//                 pre-transform, per-primitive SBT index) -> 48-byte triangle records + scene bounds
//   2 morton      centroid -> 30/48/63-bit Morton key
//   3 radix sort  LSD, 8-bit digits, stable (per-block histogram -> scan -> ranked scatter)
//   4 hierarchy   Karras 2012 binary radix tree over the sorted keys (ties broken by position)
//   5 refit       bottom-up AABBs with one atomic arrival counter per internal node
//   6 collapse    level-synchronous, deterministic conversion to the 8-wide quantised layout of
//                 accel.h (greedy surface-area expansion, octant-aware slot assignment, conservative
//                 8-bit boxes), then the triangle records are written in leaf order
// The blob is self-contained: nothing in it points into the temp or input buffers.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "accel.h"
#include "common.h"
#include "loop_graph.h"
#include "rt_math.cuh"

namespace b200rt {

void retire_loops(b200rt_context ctx, bool all)
{
    size_t keep = 0;
    for (LoopGraph* g : ctx->loops) {
        if (all || g->finished()) delete g;
        else ctx->loops[keep++] = g;
    }
    ctx->loops.resize(keep);
}
void park_loop(b200rt_context ctx, LoopGraph* g) { ctx->loops.push_back(g); }

// ---------------------------------------------------------------------------------------------
// device-side description of one triangle build input
// ---------------------------------------------------------------------------------------------
struct DevInput {
    const char* verts;
    const char* indices;
    const float* xform;
    const char* sbt_index;
    uint32_t vstride, istride, iformat;  // iformat: 0 none, 2 u16x3, 4 u32x3
    uint32_t sbt_size, sbt_stride;
    uint32_t prim_offset, sbt_base, num_sbt;
    uint32_t tri_start, ntris;
    uint32_t flags_off;                  // into the flat geometry-flag array
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------
// scan (exclusive, in place), templated on element type
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(T* data, size_t n, T* tile_sums)
{
    __shared__ T warp_sums[SCAN_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? data[base + i] : (T)0;
        sum += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : (T)0;
        T wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            T o = __shfl_up_sync(0xffffffffu, wi, off);
            if (lane >= off) wi += o;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    T run = warp_sums[warp] + (incl - sum);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(T* data, size_t n, const T* tile_offsets)
{
    const T off = tile_offsets[blockIdx.x];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) data[base + i] += off;
}

template <typename T>
static size_t scan_temp_elems(size_t n)
{
    size_t total = 0;
    while (n > SCAN_TILE) {
        n = div_up(n, SCAN_TILE);
        total += n;
    }
    return total + 1;
}

// exclusive scan of data[0..n) in place; tmp must hold scan_temp_elems(n) elements
template <typename T>
static int exclusive_scan(b200rt_context ctx, T* data, size_t n, T* tmp, cudaStream_t s)
{
    if (n == 0) return 0;
    const unsigned tiles = div_up(n, SCAN_TILE);
    if (tiles == 1) {
        scan_tiles_kernel<T><<<1, SCAN_THREADS, 0, s>>>(data, n, (T*)nullptr);
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
    scan_tiles_kernel<T><<<tiles, SCAN_THREADS, 0, s>>>(data, n, tmp);
    B2_LAUNCH_CHECK(ctx);
    int rc = exclusive_scan<T>(ctx, tmp, tiles, tmp + tiles, s);
    if (rc) return rc;
    scan_add_kernel<T><<<tiles, SCAN_THREADS, 0, s>>>(data, n, tmp);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// scan whose length lives on the device (exclusive, in place): the loops of the build (clustering rounds, collapse levels) run as CUDA
// graphs whose kernels have fixed launch shapes and read their problem size from device memory.  DS_BLOCKS blocks own one contiguous
// chunk each: (1) chunk totals, (2) one block scans the totals and writes the grand total, (3) every block rescans its chunk.
// Deterministic (no atomics), so the hierarchy and the node order are the same from run to run.
// ---------------------------------------------------------------------------------------------
constexpr int DS_BLOCKS = 592, DS_THREADS = 256;  // 4 blocks per SM on 148 SMs

__device__ __forceinline__ void ds_chunk(uint32_t n, uint32_t& first, uint32_t& last)
{
    const uint32_t chunk = ((n + DS_BLOCKS - 1) / DS_BLOCKS + DS_THREADS - 1) / DS_THREADS * DS_THREADS;
    first = min(n, blockIdx.x * chunk);
    last = min(n, first + chunk);
}

template <typename T>
__device__ __forceinline__ T ds_block_scan(T v, T* wsum, T& total)  // exclusive scan over the block's threads; total = block sum
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    __syncthreads();  // wsum may still be read by the previous call
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    T before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < DS_THREADS / 32; ++w) { const T x = wsum[w]; before += w < warp ? x : (T)0; tot += x; }
    total = tot;
    return before + incl - v;
}

template <typename T>
__global__ void __launch_bounds__(DS_THREADS) dscan_reduce_kernel(const T* __restrict__ data, const uint32_t* __restrict__ n_dev, T* __restrict__ block_sums)
{
    __shared__ T wsum[DS_THREADS / 32];
    uint32_t first, last;
    ds_chunk(*n_dev, first, last);
    T sum = 0;
    for (uint32_t i = first + threadIdx.x; i < last; i += DS_THREADS) sum += data[i];
    T total;
    ds_block_scan<T>(sum, wsum, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(1024) dscan_spine_kernel(T* __restrict__ block_sums, T* __restrict__ total_out)
{
    __shared__ T wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T v = threadIdx.x < DS_BLOCKS ? block_sums[threadIdx.x] : (T)0;
    T incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    T before = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { const T x = wsum[w]; before += w < warp ? x : (T)0; tot += x; }
    if (threadIdx.x < DS_BLOCKS) block_sums[threadIdx.x] = before + incl - v;
    if (threadIdx.x == 0) *total_out = tot;
}

template <typename T>
__global__ void __launch_bounds__(DS_THREADS) dscan_apply_kernel(T* __restrict__ data, const uint32_t* __restrict__ n_dev, const T* __restrict__ block_sums)
{
    __shared__ T wsum[DS_THREADS / 32];
    uint32_t first, last;
    ds_chunk(*n_dev, first, last);
    T run = block_sums[blockIdx.x];
    for (uint32_t base = first; base < last; base += DS_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const T v = i < last ? data[i] : (T)0;
        T total;
        const T ex = ds_block_scan<T>(v, wsum, total);
        if (i < last) data[i] = run + ex;
        run += total;
    }
}

// appends the three kernels to a loop body; block_sums: DS_BLOCKS elements of T; *total_out = sum of data[0 .. *n_dev)
template <typename T>
static int dscan_add(LoopGraph& g, T* data, const uint32_t* n_dev, T* block_sums, T* total_out)
{
    static_assert(DS_BLOCKS <= 1024, "one block scans the chunk totals");
    if (int rc = g.add((const void*)dscan_reduce_kernel<T>, DS_BLOCKS, DS_THREADS, 0, (const T*)data, n_dev, block_sums)) return rc;
    if (int rc = g.add((const void*)dscan_spine_kernel<T>, 1, 1024, 0, block_sums, total_out)) return rc;
    return g.add((const void*)dscan_apply_kernel<T>, DS_BLOCKS, DS_THREADS, 0, data, n_dev, (const T*)block_sums);
}

// ---------------------------------------------------------------------------------------------
// 1. gather
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_to_ordered(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u)
{
    const uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

__global__ void init_bounds_kernel(uint32_t* b)
{
    if (threadIdx.x < 3) b[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) b[threadIdx.x] = 0u;
}

// One thread per triangle, grid-stride over a persistent grid: the scene bounds are reduced in registers over all the triangles a
// thread sees, then per warp, then per block, and only then go to the six global words — one atomic per warp and word (1.5 M warps
// hitting six addresses of one L2 slice for 50 M triangles) used to be most of this kernel's time.  Vertex records of 16 bytes (the
// samples' float4 / Vertex{x, y, z, pad}) are read with one 128-bit load each.
__global__ void __launch_bounds__(256) gather_tris_kernel(const DevInput* __restrict__ inputs, int num_inputs, uint32_t ntris,
                                                           const uint32_t* __restrict__ geom_flags, float4* __restrict__ tri_tmp,
                                                           uint32_t* __restrict__ bounds)
{
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < ntris; g += gridDim.x * blockDim.x) {
        int k = 0;
        while (k + 1 < num_inputs && g >= inputs[k + 1].tri_start) ++k;
        const DevInput in = inputs[k];
        const uint32_t p = g - in.tri_start;
        uint32_t i0, i1, i2;
        if (in.iformat == 4) {
            const uint32_t* ip = (const uint32_t*)(in.indices + (size_t)p * in.istride);
            i0 = ip[0]; i1 = ip[1]; i2 = ip[2];
        } else if (in.iformat == 2) {
            const uint16_t* ip = (const uint16_t*)(in.indices + (size_t)p * in.istride);
            i0 = ip[0]; i1 = ip[1]; i2 = ip[2];
        } else {
            i0 = 3 * p; i1 = 3 * p + 1; i2 = 3 * p + 2;
        }
        float3 v0, v1, v2;
        if ((((uintptr_t)in.verts | (uintptr_t)in.vstride) & 15u) == 0u) {
            const float4 a = __ldg((const float4*)(in.verts + (size_t)i0 * in.vstride));
            const float4 b = __ldg((const float4*)(in.verts + (size_t)i1 * in.vstride));
            const float4 c = __ldg((const float4*)(in.verts + (size_t)i2 * in.vstride));
            v0 = f3(a.x, a.y, a.z); v1 = f3(b.x, b.y, b.z); v2 = f3(c.x, c.y, c.z);
        } else {
            const float* a = (const float*)(in.verts + (size_t)i0 * in.vstride);
            const float* b = (const float*)(in.verts + (size_t)i1 * in.vstride);
            const float* c = (const float*)(in.verts + (size_t)i2 * in.vstride);
            v0 = f3(a[0], a[1], a[2]); v1 = f3(b[0], b[1], b[2]); v2 = f3(c[0], c[1], c[2]);
        }
        if (in.xform) {
            v0 = xform_point(in.xform, v0);
            v1 = xform_point(in.xform, v1);
            v2 = xform_point(in.xform, v2);
        }
        uint32_t local_sbt = 0;
        if (in.sbt_index) {
            const char* sp = in.sbt_index + (size_t)p * in.sbt_stride;
            local_sbt = in.sbt_size == 4 ? *(const uint32_t*)sp : (in.sbt_size == 2 ? *(const uint16_t*)sp : *(const uint8_t*)sp);
        }
        if (local_sbt >= in.num_sbt) local_sbt = in.num_sbt ? in.num_sbt - 1 : 0;
        const uint32_t gf = geom_flags[in.flags_off + local_sbt] & 0xffu;
        const uint32_t sbt = ((in.sbt_base + local_sbt) & TRI_SBT_MASK) | (gf << TRI_FLAG_SHIFT);
        tri_tmp[3 * (size_t)g + 0] = make_float4(v0.x, v0.y, v0.z, __uint_as_float(in.prim_offset + p));
        tri_tmp[3 * (size_t)g + 1] = make_float4(v1.x, v1.y, v1.z, __uint_as_float(sbt));
        tri_tmp[3 * (size_t)g + 2] = make_float4(v2.x, v2.y, v2.z, __uint_as_float(g));
        lo[0] = fminf(lo[0], fminf(v0.x, fminf(v1.x, v2.x))); hi[0] = fmaxf(hi[0], fmaxf(v0.x, fmaxf(v1.x, v2.x)));
        lo[1] = fminf(lo[1], fminf(v0.y, fminf(v1.y, v2.y))); hi[1] = fmaxf(hi[1], fmaxf(v0.y, fmaxf(v1.y, v2.y)));
        lo[2] = fminf(lo[2], fminf(v0.z, fminf(v1.z, v2.z))); hi[2] = fmaxf(hi[2], fmaxf(v0.z, fmaxf(v1.z, v2.z)));
    }
    __shared__ float sh_lo[3][8], sh_hi[3][8];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float l = lo[a], h = hi[a];
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, off));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, off));
        }
        if (lane == 0) { sh_lo[a][wid] = l; sh_hi[a][wid] = h; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int a = (int)threadIdx.x;
        float l = sh_lo[a][0], h = sh_hi[a][0];
        for (int w = 1; w < 8; ++w) { l = fminf(l, sh_lo[a][w]); h = fmaxf(h, sh_hi[a][w]); }
        if (l <= h) {
            atomicMin(&bounds[a], float_to_ordered(l));
            atomicMax(&bounds[3 + a], float_to_ordered(h));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 2. morton keys
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread3(uint64_t x)
{
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void __launch_bounds__(256) morton_kernel(const float4* __restrict__ tri_tmp, uint32_t ntris,
                                                      const uint32_t* __restrict__ bounds, int bits_per_axis,
                                                      uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ntris) return;
    const float4 a = tri_tmp[3 * (size_t)g], b = tri_tmp[3 * (size_t)g + 1], c = tri_tmp[3 * (size_t)g + 2];
    const float lox = ordered_to_float(bounds[0]), loy = ordered_to_float(bounds[1]), loz = ordered_to_float(bounds[2]);
    const float hix = ordered_to_float(bounds[3]), hiy = ordered_to_float(bounds[4]), hiz = ordered_to_float(bounds[5]);
    const float cx = 0.5f * (fminf(a.x, fminf(b.x, c.x)) + fmaxf(a.x, fmaxf(b.x, c.x)));
    const float cy = 0.5f * (fminf(a.y, fminf(b.y, c.y)) + fmaxf(a.y, fmaxf(b.y, c.y)));
    const float cz = 0.5f * (fminf(a.z, fminf(b.z, c.z)) + fmaxf(a.z, fmaxf(b.z, c.z)));
    const float cells = (float)(1u << bits_per_axis);
    const uint32_t maxc = (1u << bits_per_axis) - 1u;
    auto quant = [&](float v, float lo, float hi) -> uint64_t {
        const float ext = hi - lo;
        float u = ext > 0.0f ? (v - lo) / ext : 0.0f;
        u = fminf(fmaxf(u * cells, 0.0f), (float)maxc);
        return (uint64_t)min((uint32_t)u, maxc);
    };
    const uint64_t key = (spread3(quant(cx, lox, hix)) << 2) | (spread3(quant(cy, loy, hiy)) << 1) | spread3(quant(cz, loz, hiz));
    keys[g] = key;
    vals[g] = g;
}

// ---------------------------------------------------------------------------------------------
// 3. LSD radix sort, 8-bit digits, stable.  One block owns one tile of RS_TILE consecutive keys.
// ---------------------------------------------------------------------------------------------
#ifndef B200RT_RS_IPT
#define B200RT_RS_IPT 16
#endif
#ifndef B200RT_RS_MIN_CTAS
#define B200RT_RS_MIN_CTAS 3
#endif
constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_IPT = B200RT_RS_IPT, RS_TILE = RS_THREADS * RS_IPT;

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const K* __restrict__ keys, uint32_t n, int shift,
                                                              uint32_t* __restrict__ hist, uint32_t nblocks)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_IPT; ++i) {
        const uint32_t idx = base + i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(uint32_t)(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Scatter of one tile.  Ranks come from __match_any_sync within a warp plus per-(warp, digit) counters; the tile is then put in digit
// order in shared memory and written out by consecutive threads, so that every run of equal digits (16 keys on average for uniformly
// distributed digits) leaves the SM as whole 32-byte sectors instead of one 8- or 4-byte store per bucket and lane.
template <typename K>
constexpr size_t rs_scatter_smem() { return (size_t)RS_TILE * (sizeof(K) + sizeof(uint32_t)); }

template <typename K>
__global__ void __launch_bounds__(RS_THREADS, B200RT_RS_MIN_CTAS) rs_scatter_kernel(const K* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                                 K* __restrict__ kout, uint32_t* __restrict__ vout, uint32_t n,
                                                                 int shift, const uint32_t* __restrict__ offs, uint32_t nblocks)
{
    extern __shared__ __align__(16) unsigned char rs_stage[];
    K* skeys = (K*)rs_stage;
    uint32_t* svals = (uint32_t*)(skeys + RS_TILE);
    __shared__ uint32_t wcount[RS_WARPS][256];
    __shared__ uint32_t dstart[256];  // position of the tile's first key with digit d in the staged tile
    __shared__ uint32_t gdelta[256];  // global position of that key minus dstart[d] (modulo 2^32)
    __shared__ uint32_t wsum[RS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const uint32_t wbase = blockIdx.x * RS_TILE + warp * (32 * RS_IPT);
    K k[RS_IPT];
    uint16_t rank[RS_IPT];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) {
        const uint32_t idx = wbase + r * 32 + lane;
        const bool valid = idx < n;
        k[r] = valid ? kin[idx] : (K)~(K)0;
        const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
        const uint32_t m = __match_any_sync(0xffffffffu, valid ? d : (256u + lane));
        const uint32_t cnt = valid ? wcount[warp][d] : 0u;
        rank[r] = (uint16_t)(cnt + __popc(m & lt));
        __syncwarp();
        if (valid && (m & lt) == 0u) wcount[warp][d] = cnt + __popc(m);
        __syncwarp();
    }
    __syncthreads();
    {
        // thread d: counts of digit d over the warps -> start of every warp's share; tile totals -> exclusive scan over the digits
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) before += w < warp ? wsum[w] : 0u;
        const uint32_t excl = before + incl - run;
        dstart[d] = excl;
        gdelta[d] = offs[(size_t)d * nblocks + blockIdx.x] - excl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) {
        const uint32_t idx = wbase + r * 32 + lane;
        if (idx < n) {
            const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
            const uint32_t pos = dstart[d] + wcount[warp][d] + rank[r];
            skeys[pos] = k[r];
            svals[pos] = vin[idx];
        }
    }
    __syncthreads();
    const uint32_t tile_first = blockIdx.x * RS_TILE;
    const uint32_t tile_n = n - tile_first < (uint32_t)RS_TILE ? n - tile_first : (uint32_t)RS_TILE;
    for (uint32_t pos = threadIdx.x; pos < tile_n; pos += RS_THREADS) {
        const K key = skeys[pos];
        const uint32_t dst = gdelta[(uint32_t)(key >> shift) & 255u] + pos;
        kout[dst] = key;
        vout[dst] = svals[pos];
    }
}

template <typename K>
static int rs_scatter_prepare(b200rt_context ctx)
{
    // per device function and per device: set on every sort (microseconds), the process may drive several devices
    B2_CUDA(ctx, cudaFuncSetAttribute(rs_scatter_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_scatter_smem<K>()));
    return 0;
}

// Stable LSD sort of (32-bit key, 32-bit value) pairs, `passes` digits of 8 bits from bit 0 — the ray reordering of pathtracer.cu.
// keys / vals are ping-pong buffers; returns the index of the buffer holding the result.  hist: 256 * div_up(n, RS_TILE) words,
// scan_tmp: radix_sort_scan_words(n) words.
size_t radix_sort_hist_words(size_t n) { return 256 * (size_t)std::max(1u, div_up(n, RS_TILE)); }
size_t radix_sort_scan_words(size_t n) { return scan_temp_elems<uint32_t>(radix_sort_hist_words(n)); }
int radix_sort_pairs32(b200rt_context ctx, cudaStream_t s, uint32_t* keys[2], uint32_t* vals[2], uint32_t n, int passes, uint32_t* hist, uint32_t* scan_tmp,
                       int* result)
{
    int cur = 0;
    const uint32_t blocks = std::max(1u, div_up(n, RS_TILE));
    if (int rc = rs_scatter_prepare<uint32_t>(ctx)) return rc;
    for (int pass = 0; pass < passes && n > 1; ++pass) {
        rs_hist_kernel<uint32_t><<<blocks, RS_THREADS, 0, s>>>(keys[cur], n, pass * 8, hist, blocks);
        B2_LAUNCH_CHECK(ctx);
        int rc = exclusive_scan<uint32_t>(ctx, hist, 256 * (size_t)blocks, scan_tmp, s);
        if (rc) return rc;
        rs_scatter_kernel<uint32_t><<<blocks, RS_THREADS, rs_scatter_smem<uint32_t>(), s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, pass * 8, hist, blocks);
        B2_LAUNCH_CHECK(ctx);
        cur ^= 1;
    }
    *result = cur;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// 4. Karras hierarchy.  Node ids: internal i in [0, N-1), leaf at sorted position s is (N-1)+s.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

// Boxes of the binary hierarchy: {lo.xyz, left child} {hi.xyz, right child} of node id side by side (32 bytes, one DRAM sector), reached
// through two strided views so that kernels keep writing box_lo[id] / box_hi[id].  Side by side because every consumer (bottom-up
// union, clustering, collapse) wants both halves of a box it fetches by a data-dependent id: two arrays cost two sectors per box.
struct BoxArr {
    float4* p;
    __device__ __forceinline__ float4& operator[](size_t i) const { return p[2 * i]; }
};
// The two child words also say how many leaves the node holds, as far as the collapse cares (a subtree of at most three triangles
// becomes a leaf child, anything larger an inner child): bit 31 of lo.w = "at most three", bit 31 of hi.w = "exactly three" (else two).
// With that the collapse never touches a second per-node array: one sector per node it looks at.
constexpr uint32_t CHILD_ID_MASK = 0x7fffffffu;
__device__ __forceinline__ int child_id(float w) { return (int)(__float_as_uint(w) & CHILD_ID_MASK); }
__device__ __forceinline__ int leaves_upto4(float lo_w, float hi_w)  // 2, 3, or 4 for "more than three"
{
    return (__float_as_uint(lo_w) >> 31) ? 2 + (int)(__float_as_uint(hi_w) >> 31) : 4;
}
__device__ __forceinline__ float child_word_lo(int left, int leaves) { return __uint_as_float((uint32_t)left | (leaves <= 3 ? 0x80000000u : 0u)); }
__device__ __forceinline__ float child_word_hi(int right, int leaves) { return __uint_as_float((uint32_t)right | (leaves == 3 ? 0x80000000u : 0u)); }

// The radix tree over the sorted keys (Karras 2012: node = maximal range of leaves with a common key prefix; equal keys are told apart by
// their position) AND its boxes in ONE bottom-up pass (Apetrei 2014): a thread starts at leaf s with the range [s, s] and its box in
// registers.  The parent of a range [l, r] is the split position next to it whose keys differ least — p = r (the range is p's left
// child) when key[r] ^ key[r+1] < key[l-1] ^ key[l], else p = l - 1 (right child); for a valid range the two differences never have the
// same leading bit, so the comparison is strict.  The thread leaves its node id and its far bound in p's record, then swaps its
// subtree height into p's arrival word: the first of the two children to arrive stops, the second finds its sibling's id, bound and
// height there, reads the sibling's box (one sector), and goes on as node p with the union in registers.  Compared with the two-kernel
// form this replaces (top-down binary searches for every node, then a bottom-up pass re-reading both children's boxes through parent
// pointers) a node costs one box read instead of two, no searches and no parent array: leaf boxes + hierarchy 6.35 -> 4.34 ms for 50 M
// triangles (profiles/r02_build.md).  Same tree: internal node p's children are [l, p] and [p + 1, r]; only the numbering differs (p is
// the split position, the root is whichever node ends up with [0, n - 1] and goes to *root_out).  `range` is only the place where the two
// children of a node leave their far bounds for each other.  A subtree of two or three leaves starts at the sorted position its left
// child names (a leaf's own position; an inner left child of two leaves [q, q + 1] is node q): the collapse reads it from there
// (collapse_emit_kernel).
__device__ __forceinline__ bool split_less(const uint64_t* __restrict__ keys, int a, int b)  // is the key step at position a smaller than the one at b?
{
    const uint64_t xa = __ldg(keys + a) ^ __ldg(keys + a + 1), xb = __ldg(keys + b) ^ __ldg(keys + b + 1);
    if (xa != xb) return xa < xb;
    return (uint32_t)(a ^ (a + 1)) < (uint32_t)(b ^ (b + 1));  // both steps zero (duplicate keys): the positions decide, as in Karras's delta
}

constexpr uint32_t ARRIVED = 0x80000000u;

__global__ void __launch_bounds__(256) radix_tree_kernel(const uint64_t* __restrict__ keys, const float4* __restrict__ tri_tmp, const uint32_t* __restrict__ vals, int n,
                                                          const BoxArr box_lo, const BoxArr box_hi, int2* range, uint32_t* arrive, uint8_t* __restrict__ height,
                                                          uint32_t* __restrict__ root_out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const size_t g = vals[s];
    const float4 a = tri_tmp[3 * g], b = tri_tmp[3 * g + 1], c = tri_tmp[3 * g + 2];
    float lx = fminf(a.x, fminf(b.x, c.x)), ly = fminf(a.y, fminf(b.y, c.y)), lz = fminf(a.z, fminf(b.z, c.z));
    float hx = fmaxf(a.x, fmaxf(b.x, c.x)), hy = fmaxf(a.y, fmaxf(b.y, c.y)), hz = fmaxf(a.z, fmaxf(b.z, c.z));
    int id = n - 1 + s, l = s, r = s, h = 0;
    box_lo[id] = make_float4(lx, ly, lz, 0.f);
    box_hi[id] = make_float4(hx, hy, hz, 0.f);
    for (;;) {
        if (l == 0 && r == n - 1) { *root_out = (uint32_t)id; return; }
        const bool left_child = l == 0 || (r != n - 1 && split_less(keys, r, l - 1));
        const int p = left_child ? r : l - 1;
        // what the sibling needs from this subtree: its node id (into the parent's child slot) and its far bound
        volatile float* slot_w = &(left_child ? box_lo[p] : box_hi[p]).w;
        volatile int* bound = left_child ? &range[p].x : &range[p].y;
        *slot_w = __int_as_float(id);
        *bound = left_child ? l : r;
        const uint32_t old = atomic_exch_release(&arrive[p], ARRIVED | (uint32_t)h);   // release: the two stores above are visible first (common.h)
        if (!(old & ARRIVED)) return;  // first arrival: the sibling subtree is not finished yet, its thread will carry on
        // second arrival: the sibling's stores are visible (its fence precedes its exchange); read them past the L1
        const int sib = __float_as_int(__ldcg(&(left_child ? box_hi[p] : box_lo[p]).w));
        const int far = __ldcg(left_child ? &range[p].y : &range[p].x);
        const float4 slo = __ldcg(&box_lo[sib]), shi = __ldcg(&box_hi[sib]);
        lx = fminf(lx, slo.x); ly = fminf(ly, slo.y); lz = fminf(lz, slo.z);
        hx = fmaxf(hx, shi.x); hy = fmaxf(hy, shi.y); hz = fmaxf(hz, shi.z);
        h = min(max(h, (int)(old & 0xffu)) + 1, 255);
        if (left_child) r = far; else l = far;
        box_lo[p] = make_float4(lx, ly, lz, child_word_lo(left_child ? id : sib, r - l + 1));
        box_hi[p] = make_float4(hx, hy, hz, child_word_hi(left_child ? sib : id, r - l + 1));
        height[p] = (uint8_t)h;   // subtree height (edges to the deepest leaf): what the collapse's depth guard reads
        id = p;
    }
}

// ---------------------------------------------------------------------------------------------
// 4b. PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) — the quality alternative to the Karras tree: bottom-up
// agglomeration of the Morton-ordered leaves.  Every cluster looks PLOC_RADIUS positions to either side for the neighbour with which
// it would form the smallest box; mutual nearest neighbours merge into a new internal node; the cluster array is compacted; repeat
// until one cluster is left.  Prefix sums place both the survivors and the new nodes, so the tree is deterministic.  Boxes and leaf
// counts are final when a node is created (no refit pass).  Internal ids are handed out in creation order: the root is the last one.
// ---------------------------------------------------------------------------------------------
#ifndef B200RT_PLOC_RADIUS
#define B200RT_PLOC_RADIUS 16
#endif
constexpr int PLOC_RADIUS = B200RT_PLOC_RADIUS, PLOC_THREADS = 256;

// The clustering rounds run as a device-side loop (loop_graph.h): every kernel is grid-stride over the live cluster count in PlocState.
struct PlocState {
    uint32_t n;           // clusters alive
    uint32_t node_base;   // internal nodes created so far
    uint32_t cc;          // which of the two cluster arrays is current
    uint32_t round;
    unsigned long long total;  // this round's flag totals: survivors << 32 | nodes created
    uint32_t error, pad;
};
// Rounds from this one on pair the clusters by position (2 i with 2 i + 1) instead of by nearest neighbour: a round adds at most one
// level to the tree, clustering by neighbour may take arbitrarily many rounds on adversarial input, and the collapse needs a bounded
// height to keep the wide tree within the traversal stack (collapse_plan_kernel).  Typical inputs finish in about 40 rounds.
constexpr uint32_t PLOC_POSITIONAL_ROUND = 64;

__global__ void __launch_bounds__(PLOC_THREADS) ploc_nearest_kernel(const uint32_t* __restrict__ cl0, const uint32_t* __restrict__ cl1, const PlocState* __restrict__ st,
                                                                     const BoxArr box_lo, const BoxArr box_hi, uint32_t* __restrict__ nearest)
{
    __shared__ float4 slo[PLOC_THREADS + 2 * PLOC_RADIUS], shi[PLOC_THREADS + 2 * PLOC_RADIUS];
    const uint32_t n = st->n;
    const uint32_t* __restrict__ clusters = st->cc ? cl1 : cl0;
    if (st->round >= PLOC_POSITIONAL_ROUND) {
        for (uint32_t i = blockIdx.x * PLOC_THREADS + threadIdx.x; i < n; i += gridDim.x * PLOC_THREADS) nearest[i] = (i ^ 1u) < n ? (i ^ 1u) : 0xffffffffu;
        return;
    }
    for (uint32_t tile = blockIdx.x; tile * PLOC_THREADS < n; tile += gridDim.x) {
        const int base = (int)(tile * PLOC_THREADS) - PLOC_RADIUS;
        __syncthreads();
        for (int k = threadIdx.x; k < PLOC_THREADS + 2 * PLOC_RADIUS; k += PLOC_THREADS) {
            const int g = base + k;
            if (g >= 0 && g < (int)n) { const uint32_t id = clusters[g]; slo[k] = box_lo[id]; shi[k] = box_hi[id]; }
        }
        __syncthreads();
        const int i = (int)(tile * PLOC_THREADS + threadIdx.x);
        if (i >= (int)n) continue;
        const int me = threadIdx.x + PLOC_RADIUS;
        const float4 alo = slo[me], ahi = shi[me];
        float best = INFINITY;
        int bj = -1;
        for (int d = -PLOC_RADIUS; d <= PLOC_RADIUS; ++d) {
            const int j = i + d;
            if (d == 0 || j < 0 || j >= (int)n) continue;
            const float4 blo = slo[me + d], bhi = shi[me + d];
            const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
            const float a = dx * dy + dy * dz + dz * dx;
            if (a < best) { best = a; bj = j; }  // ties: the lower position (scan order), so (i, j) and (j, i) agree on equal areas
        }
        nearest[i] = (uint32_t)bj;
    }
}

// flags: high word = this position survives into the next round, low word = it creates a node
__global__ void __launch_bounds__(256) ploc_flag_kernel(const uint32_t* __restrict__ nearest, const PlocState* __restrict__ st, unsigned long long* __restrict__ flags)
{
    const uint32_t n = st->n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t j = nearest[i];
        const bool mutual = j < n && nearest[j] == i;
        const bool leader = mutual && i < j;
        flags[i] = ((unsigned long long)((!mutual || leader) ? 1u : 0u) << 32) | (leader ? 1u : 0u);
    }
}

__global__ void __launch_bounds__(256) ploc_merge_kernel(uint32_t* __restrict__ cl0, uint32_t* __restrict__ cl1, const uint32_t* __restrict__ nearest,
                                                          const PlocState* __restrict__ st, const unsigned long long* __restrict__ excl, int ninternal,
                                                          const BoxArr box_lo, const BoxArr box_hi, uint8_t* __restrict__ height)
{
    const uint32_t n = st->n, node_base = st->node_base;
    const uint32_t* __restrict__ clusters = st->cc ? cl1 : cl0;
    uint32_t* __restrict__ out = st->cc ? cl0 : cl1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t j = nearest[i];
        const bool mutual = j < n && nearest[j] == i;
        if (mutual && i > j) continue;  // absorbed by its partner
        const unsigned long long e = excl[i];
        const uint32_t pos = (uint32_t)(e >> 32);
        if (!mutual) { out[pos] = clusters[i]; continue; }
        const uint32_t id = node_base + (uint32_t)(e & 0xffffffffu);
        const uint32_t a = clusters[i], b = clusters[j];
        const float4 alo = box_lo[a], ahi = box_hi[a], blo = box_lo[b], bhi = box_hi[b];
        const int ca = (int)a < ninternal ? leaves_upto4(alo.w, ahi.w) : 1, cb = (int)b < ninternal ? leaves_upto4(blo.w, bhi.w) : 1;
        const int ha = (int)a < ninternal ? height[a] : 0, hb = (int)b < ninternal ? height[b] : 0;
        // (the leaves of a clustered node are not a range of sorted positions: the collapse walks its small subtrees)
        box_lo[id] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), child_word_lo((int)a, min(ca + cb, 4)));
        box_hi[id] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), child_word_hi((int)b, min(ca + cb, 4)));
        height[id] = (uint8_t)min(max(ha, hb) + 1, 255);
        out[pos] = id;
    }
}

__global__ void ploc_advance_kernel(PlocState* st, cudaGraphConditionalHandle cond)
{
    const unsigned long long tot = st->total;
    const uint32_t survivors = (uint32_t)(tot >> 32), created = (uint32_t)(tot & 0xffffffffu);
    if (created == 0u || survivors != st->n - created) st->error = 1u;  // cannot happen: the pair with the smallest union is mutual
    st->node_base += created;
    st->n = survivors;
    st->cc ^= 1u;
    st->round += 1u;
    cudaGraphSetConditional(cond, (st->n > 1u && !st->error && st->round < 4096u) ? 1u : 0u);
}

__global__ void ploc_init_kernel(uint32_t* clusters, uint32_t n, PlocState* st)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) clusters[i] = n - 1u + i;  // leaf ids in Morton order
    if (i == 0) { PlocState z = {}; z.n = n; *st = z; }
}

// ---------------------------------------------------------------------------------------------
// 5. refit
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) leaf_boxes_kernel(const float4* __restrict__ tri_tmp, const uint32_t* __restrict__ vals, int n,
                                                          const BoxArr box_lo, const BoxArr box_hi)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const size_t g = vals[s];
    const float4 a = tri_tmp[3 * g], b = tri_tmp[3 * g + 1], c = tri_tmp[3 * g + 2];
    box_lo[n - 1 + s] = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.f);
    box_hi[n - 1 + s] = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.f);
}

// ---------------------------------------------------------------------------------------------
// 6. collapse to the 8-wide layout
// ---------------------------------------------------------------------------------------------
// The levels run as a device-side loop (loop_graph.h); every kernel is grid-stride over the level's work list, whose length lives here.
struct CollapseState {
    uint32_t nwork;        // wide nodes of the current level
    uint32_t level_start;  // index of the level's first node
    uint32_t tri_cursor;   // triangles placed by the levels above
    uint32_t wcur;         // which of the two work lists holds the current level
    uint32_t depth;        // levels finished
    uint32_t total_nodes;
    uint32_t error;        // 1: node capacity exceeded, 2: triangle count mismatch, 3: clustering failed (internal errors)
    uint32_t pad;
    unsigned long long total;  // this level's totals: inner children << 32 | triangles in leaf children
};

__device__ __forceinline__ float box_area(const float4 lo, const float4 hi)
{
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// Wide-tree depth the traversal stack is good for (two entries per level, accel.h) and the levels a binary subtree of height h needs
// at most when every wide node below opens its tallest children first (7 openings take 3 off the height of a full binary tree, more
// off anything thinner).  Binary heights are bounded by the key length: 63 Morton bits + 30 index bits for the radix tree,
// PLOC_POSITIONAL_ROUND + 23 rounds for the clustering, so levels_needed(root) <= MAX_WIDE_DEPTH for every input.
constexpr int MAX_WIDE_DEPTH = TRAV_STACK / 2;
__device__ __forceinline__ int levels_needed(int h) { return (h + 2) / 3 + 1; }
static_assert((93 + 2) / 3 + 1 <= MAX_WIDE_DEPTH, "the traversal stack must hold the deepest tree the builder can make");

__global__ void __launch_bounds__(128) collapse_plan_kernel(const CollapseState* __restrict__ st, const uint32_t* __restrict__ work0,
                                                             const uint32_t* __restrict__ work1, int n, const BoxArr box_lo,
                                                             const BoxArr box_hi, const uint8_t* __restrict__ height, int* __restrict__ child_tmp,
                                                             unsigned long long* __restrict__ counts, int max_depth)
{
    const uint32_t nwork = st->nwork;
    const uint32_t* __restrict__ work = st->wcur ? work1 : work0;
    const int ninternal = n - 1;
    const int levels_left = max_depth - (int)st->depth;  // levels available from this one down, this one included
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nwork; i += gridDim.x * blockDim.x) {
        int ids[8];
        float area[8];
        int cnt[8];
        int m = 0;
        auto push = [&](int id) {
            ids[m] = id;
            if (id < ninternal) { const float4 lo = box_lo[id], hi = box_hi[id]; cnt[m] = leaves_upto4(lo.w, hi.w); area[m] = box_area(lo, hi); }
            else { cnt[m] = 1; area[m] = -1.0f; }
            ++m;
        };
        const int root = (int)work[i];
        int hroot = 0;
        if (root >= ninternal) push(root);  // single-triangle scene
        else { hroot = height[root]; push(child_id(box_lo[root].w)); push(child_id(box_hi[root].w)); }
        // depth guard: a subtree too tall for the levels that are left opens its tallest children first until every child is at least 3
        // lower than the root (7 openings always suffice); this takes precedence over the surface-area order below
        // (heights are read here only: one load per wide node on the ordinary path)
        if (levels_needed(hroot) > levels_left - 2) {
            while (m < 8) {
                int best = -1, bh = hroot - 3;
                for (int k = 0; k < m; ++k) {
                    if (ids[k] >= ninternal) continue;  // a leaf cannot be opened (and hroot - 3 may be negative)
                    const int hk = (int)height[ids[k]];
                    if (hk > bh) { bh = hk; best = k; }
                }
                if (best < 0) break;
                const int id = ids[best];
                const int l = child_id(box_lo[id].w), r = child_id(box_hi[id].w);
                const int save = m;
                m = best; push(l);
                m = save; push(r);
            }
        }
        // phase 1: open the largest inner subtree with more than 3 triangles; phase 2: split small multi-triangle leaves
        for (int phase = 0; phase < 2; ++phase) {
            while (m < 8) {
                int best = -1;
                float ba = -1.0f;
                for (int k = 0; k < m; ++k) {
                    const bool inner = ids[k] < ninternal;
                    const bool ok = phase == 0 ? (inner && cnt[k] > 3) : inner;
                    if (ok && area[k] > ba) { ba = area[k]; best = k; }
                }
                if (best < 0) break;
                const int id = ids[best];
                const int l = child_id(box_lo[id].w), r = child_id(box_hi[id].w);
                const int save = m;
                m = best; push(l);
                m = save; push(r);
            }
        }
        uint32_t nint = 0, ntri = 0;
        for (int k = 0; k < 8; ++k) {
            if (k < m) {
                child_tmp[(size_t)i * 8 + k] = ids[k];
                if (ids[k] < ninternal && cnt[k] > 3) ++nint; else ntri += cnt[k];
            } else child_tmp[(size_t)i * 8 + k] = -1;
        }
        counts[i] = ((unsigned long long)nint << 32) | ntri;
    }
}

// B200RT_BUILD_TIMING=2: a one-thread kernel between the kernels of the collapse loop's body adds the time since the previous stamp to
// acc[1 + k] (nanoseconds of %globaltimer); acc[0] is the previous stamp.  Debug aid: the kernels of a WHILE body cannot be bracketed by events.
__global__ void stamp_kernel(unsigned long long* acc, int k)
{
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (k >= 0) acc[1 + k] += now - acc[0];
    acc[0] = now;
}

__global__ void collapse_advance_kernel(CollapseState* st, uint32_t max_nodes, cudaGraphConditionalHandle cond)
{
    const unsigned long long tot = st->total;
    const uint32_t next_nodes = (uint32_t)(tot >> 32), level_tris = (uint32_t)(tot & 0xffffffffu);
    const uint32_t next_level_start = st->level_start + st->nwork;
    st->total_nodes = next_level_start;
    if ((unsigned long long)next_level_start + next_nodes > max_nodes) st->error = 1u;
    st->level_start = next_level_start;
    st->tri_cursor += level_tris;
    st->nwork = next_nodes;
    st->wcur ^= 1u;
    st->depth += 1u;
    cudaGraphSetConditional(cond, (st->nwork > 0u && !st->error) ? 1u : 0u);
}

// Eight lanes per wide node, one per child: the child boxes arrive as one coalesced read of the plan's ids and eight independent box
// fetches; the union is the box of the binary node the wide node stands for (read, not reduced); every lane scores the eight slots for
// its own child and the greedy assignment costs one shuffle per child; the slot-ordered masks (inner children, triangle counts) are one
// packed OR-reduction; and every lane stores its child's bytes straight into the node image (the byte stores of a group coalesce).
// (One thread per node held all eight boxes and the whole node image in registers and local memory: 136 registers, 752 bytes of
// stack, 3.6 of the 6.4 ms the collapse took on 50 M triangles; profiles/r02_build.md.)
__global__ void __launch_bounds__(128) collapse_emit_kernel(CollapseState* __restrict__ st, uint32_t max_nodes, int n,
                                                             const BoxArr box_lo, const BoxArr box_hi, int scattered_leaves,
                                                             const int* __restrict__ child_tmp,
                                                             const unsigned long long* __restrict__ excl, uint32_t* __restrict__ work0,
                                                             uint32_t* __restrict__ work1, uint32_t* __restrict__ dest, uint4* __restrict__ nodes_out,
                                                             uint32_t node_bytes)
{
    const uint32_t nwork = st->nwork, level_start = st->level_start, next_level_start = level_start + nwork, tri_cursor = st->tri_cursor;
    const uint32_t* __restrict__ work = st->wcur ? work1 : work0;
    uint32_t* __restrict__ next_work = st->wcur ? work0 : work1;
    const int ninternal = n - 1;
    const uint32_t sub = threadIdx.x & 7u;               // lane within the node's group = child index k
    const uint32_t gbase = threadIdx.x & 24u;            // first lane of the group within the warp
    const uint32_t gmask = 0xffu << gbase;
    constexpr uint32_t GROUPS = 128 / 8;
    for (uint32_t i = blockIdx.x * GROUPS + (threadIdx.x >> 3); i < nwork; i += gridDim.x * GROUPS) {
        const uint32_t node_index = level_start + i;
        if (node_index >= max_nodes) { st->error = 1u; continue; }
        const unsigned long long ex = excl[i];
        const uint32_t child_base = next_level_start + (uint32_t)(ex >> 32);
        const uint32_t tri_base = tri_cursor + (uint32_t)(ex & 0xffffffffu);
        if (child_base > max_nodes - min(max_nodes, 8u)) { st->error = 1u; continue; }  // its children would not fit: the level check reports it

        const int kid = child_tmp[(size_t)i * 8 + sub];   // the plan lists the children densely from 0
        const bool used = kid >= 0;
        const int m = __popc(__ballot_sync(gmask, used) & gmask);
        float4 klo = make_float4(0.f, 0.f, 0.f, 0.f), khi = klo;
        if (used) { klo = box_lo[kid]; khi = box_hi[kid]; }
        // the children's union is the box of the binary node this wide node stands for (min / max are exact: bit for bit the same)
        const uint32_t root = work[i];
        const float4 rlo = box_lo[root], rhi = box_hi[root];
        const float3 Lo = f3(rlo.x, rlo.y, rlo.z), Hi = f3(rhi.x, rhi.y, rhi.z);
        // conservative padding: 2^-16 of the largest extent on every side (covers the rounding of the
        // slab arithmetic in traverse.cuh; see DESIGN.md "why box tests never cull a true hit")
        const float maxext = fmaxf(fmaxf(Hi.x - Lo.x, Hi.y - Lo.y), Hi.z - Lo.z);
        // The pad is applied with directed rounding: far from the origin it is smaller than the spacing of the coordinates (extent 4 at 1e6:
        // pad 6e-5 against an ulp of 0.0625) and a round-to-nearest `hi + pad` would hand back `hi` — an unpadded box, which the slab test
        // with clamped zero direction components rejects for a ray lying exactly in one of its face planes.  Rounded outwards, every box
        // grows by at least one ulp on every side.
        const float pad = fmaxf(maxext * 1.52587890625e-5f, 1e-30f);
        const float3 P = f3(__fsub_rd(Lo.x, pad), __fsub_rd(Lo.y, pad), __fsub_rd(Lo.z, pad));
        // grid exponent per axis: smallest e with 255 * 2^e >= padded extent
        uint32_t eb[3];
        float scale[3], inv_scale[3];
        const float Pa[3] = {P.x, P.y, P.z};
        const float Ha[3] = {__fadd_ru(Hi.x, pad), __fadd_ru(Hi.y, pad), __fadd_ru(Hi.z, pad)};
        const float ext3[3] = {__fsub_ru(Ha[0], P.x), __fsub_ru(Ha[1], P.y), __fsub_ru(Ha[2], P.z)};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = ext3[a] * (1.0000002f / 255.0f);
            const uint32_t bits = __float_as_uint(v);
            uint32_t e = (bits >> 23) + ((bits & 0x7fffffu) ? 1u : 0u);
            e = min(max(e, 1u), 254u);
            while (e < 254u && fm(255.0f, __uint_as_float(e << 23), Pa[a]) < Ha[a]) ++e;
            eb[a] = e;
            scale[a] = __uint_as_float(e << 23);
            inv_scale[a] = e < 254u ? __uint_as_float((254u - e) << 23) : __uint_as_float(0x00400000u);   // 2^(127 - e), exact: x * inv == x / scale
        }
        // ---- slot assignment: greedy in child order, each child takes the free slot whose octant sign vector best matches its centroid
        // offset (slot bit 4/2/1 set = child on the +x/+y/+z side; ties: the lower slot).  Every lane scores the eight slots for its
        // own child; child k's pick goes round with one shuffle.
        const float3 ctr = f3(0.5f * (Lo.x + Hi.x), 0.5f * (Lo.y + Hi.y), 0.5f * (Lo.z + Hi.z));
        const float ox = 0.5f * (klo.x + khi.x) - ctr.x, oy = 0.5f * (klo.y + khi.y) - ctr.y, oz = 0.5f * (klo.z + khi.z) - ctr.z;
        float score[8];
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) score[sl] = ((sl & 4) ? ox : -ox) + ((sl & 2) ? oy : -oy) + ((sl & 1) ? oz : -oz);
        uint32_t taken = 0u;
        int slot = -1;
        for (int k = 0; k < m; ++k) {
            int bs = -1;
            float bc = -INFINITY;
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
                if (!((taken >> sl) & 1u) && score[sl] > bc) { bc = score[sl]; bs = sl; }
            if (bs < 0) bs = __ffs(~taken & 0xffu) - 1;   // scores that are not numbers: any free slot
            const int pick = __shfl_sync(gmask, bs, gbase + k);
            taken |= 1u << pick;
            if ((int)sub == k) slot = pick;
        }
        if (!used) slot = __fns(~taken & 0xffu, 0u, (int)sub - m + 1);   // the lanes without a child fill the empty slots with "no child"
        // ---- this lane's child in its slot
        uint32_t qlo[3] = {255u, 255u, 255u}, qhi[3] = {0u, 0u, 0u};
        float flo[3] = {0.f, 0.f, 0.f}, fhi[3] = {0.f, 0.f, 0.f};  // Node8F: offsets from P, rounded outwards
        int count = 0, first = 0;
        bool inner = false;
        if (used) {
            const float lo3[3] = {__fsub_rd(klo.x, pad), __fsub_rd(klo.y, pad), __fsub_rd(klo.z, pad)};
            const float hi3[3] = {__fadd_ru(khi.x, pad), __fadd_ru(khi.y, pad), __fadd_ru(khi.z, pad)};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (node_bytes == NODE8F_BYTES) {
                    // plane = P + off in exact arithmetic (the traversal evaluates (P + off - o) / d as fma(off, 1/d, (P - o) * (1/d)) and
                    // never forms P + off in floating point): rounding the differences outwards keeps the padded box enclosed
                    flo[a] = __fsub_rd(lo3[a], Pa[a]);
                    fhi[a] = __fsub_ru(hi3[a], Pa[a]);
                    continue;
                }
                int ql = (int)floorf((lo3[a] - Pa[a]) * inv_scale[a]);
                ql = min(max(ql, 0), 255);
                while (ql > 0 && fm((float)ql, scale[a], Pa[a]) > lo3[a]) --ql;
                int qh = (int)ceilf((hi3[a] - Pa[a]) * inv_scale[a]);
                qh = min(max(qh, 0), 255);
                while (qh < 255 && fm((float)qh, scale[a], Pa[a]) < hi3[a]) ++qh;
                qlo[a] = (uint32_t)ql;
                qhi[a] = (uint32_t)qh;
            }
            if (kid < ninternal) {
                count = leaves_upto4(klo.w, khi.w);
                // two or three leaves: consecutive sorted positions from the one the left child names (radix tree), or wherever the
                // clustering found them (scattered_leaves: walk the subtree)
                const int left = child_id(klo.w);
                first = scattered_leaves ? -1 : (left >= ninternal ? left - ninternal : left);
            } else { first = kid - ninternal; count = 1; }
            inner = kid < ninternal && count > 3;
        }
        // slot-ordered masks in one OR-reduction: inner children | leaf children with an odd count << 8 | with two or three << 16
        const uint32_t lcount = (used && !inner) ? (uint32_t)count : 0u;   // 0 .. 3 triangles in a leaf child
        uint32_t masks = ((inner ? 1u : 0u) | ((lcount & 1u) << 8) | ((lcount >> 1) << 16)) << slot;
#pragma unroll
        for (int d = 4; d >= 1; d >>= 1) masks |= __shfl_xor_sync(gmask, masks, d);
        const uint32_t imask = masks & 0xffu, c0 = (masks >> 8) & 0xffu, c1 = (masks >> 16) & 0xffu;
        const uint32_t below = (1u << slot) - 1u;
        const uint32_t tri_off = (uint32_t)__popc(c0 & below) + 2u * (uint32_t)__popc(c1 & below);   // triangles of the leaf children in lower slots
        uint32_t meta = 0u;
        if (inner) {
            meta = (1u << 5) | (24u + (uint32_t)slot);
            next_work[(child_base - next_level_start) + (uint32_t)__popc(imask & below)] = (uint32_t)kid;
        } else if (used) {
            const uint32_t unary = count == 1 ? 1u : (count == 2 ? 3u : 7u);
            meta = (unary << 5) | tri_off;
            if (first >= 0) {
                for (int j = 0; j < count; ++j) dest[first + j] = tri_base + tri_off + (uint32_t)j;
            } else {
                // PLOC subtree of 2 or 3 leaves: its leaves are not consecutive sorted positions, walk it (left to right)
                int stack2[4] = {kid, -1, -1, -1};
                int sp2 = 1, j = 0;
                while (sp2 > 0) {
                    const int nd = stack2[--sp2];
                    if (nd >= ninternal) { dest[nd - ninternal] = tri_base + tri_off + (uint32_t)j; ++j; continue; }
                    stack2[sp2++] = child_id(box_hi[nd].w);
                    stack2[sp2++] = child_id(box_lo[nd].w);
                }
            }
        }
        // ---- the node image.  Words 0-5 (both formats): origin, exponents | inner mask, child base, triangle base; bytes 24-31: meta
        uint32_t* out = (uint32_t*)((char*)nodes_out + (size_t)node_index * node_bytes);
        uint8_t* outb = (uint8_t*)out;
        const uint32_t w3 = node_bytes == NODE8F_BYTES ? imask : (eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24));
        const uint32_t head = sub == 0u ? __float_as_uint(P.x) : sub == 1u ? __float_as_uint(P.y) : sub == 2u ? __float_as_uint(P.z) : sub == 3u ? w3 : sub == 4u ? child_base : tri_base;
        if (sub < 6u) out[sub] = head;
        outb[24 + slot] = (uint8_t)meta;
        if (node_bytes == NODE8F_BYTES) {
            // plane-major: [lo x][lo y][lo z][hi x][hi y][hi z], eight floats each
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                out[8 + 8 * a + slot] = __float_as_uint(flo[a]);
                out[32 + 8 * a + slot] = __float_as_uint(fhi[a]);
            }
            continue;
        }
        // Node8 bytes 32-79: six planes (lo x, lo y, lo z, hi x, hi y, hi z) of eight slot bytes
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            outb[32 + 8 * a + slot] = (uint8_t)qlo[a];
            outb[56 + 8 * a + slot] = (uint8_t)qhi[a];
        }
    }
}

__global__ void __launch_bounds__(256) scatter_tris_kernel(const float4* __restrict__ tri_tmp, const uint32_t* __restrict__ vals,
                                                            const uint32_t* __restrict__ dest, uint32_t n, float4* __restrict__ tris_out)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const size_t g = vals[s];
    const size_t d = dest[s];
    tris_out[3 * d + 0] = tri_tmp[3 * g + 0];
    tris_out[3 * d + 1] = tri_tmp[3 * g + 1];
    tris_out[3 * d + 2] = tri_tmp[3 * g + 2];
}

// The header is completed on the device from the collapse's final state; the compacted size goes to the caller's emit address.
__global__ void write_header_kernel(AccelHeader* h, AccelHeader v, const uint32_t* bounds, const CollapseState* st, unsigned long long* emit0,
                                    unsigned long long* emit1)
{
    for (int a = 0; a < 6; ++a) v.bounds[a] = ordered_to_float(bounds[a]);
    if (st) {
        v.num_nodes = st->total_nodes;
        v.depth = st->depth;
        v.error = st->error ? st->error : (st->tri_cursor != v.num_tris ? 2u : 0u);
    }
    v.total_bytes = HEADER_BYTES + (uint64_t)v.num_nodes * v.node_bytes + (uint64_t)v.num_tris * TRI_BYTES;
    *h = v;
    const unsigned long long compacted = (v.total_bytes + 127ull) / 128ull * 128ull;
    if (emit0) *emit0 = compacted;
    if (emit1) *emit1 = compacted;
}

// first work item of the collapse: the root of the binary hierarchy (radix tree: *root_dev; clustering: the last cluster standing)
__global__ void collapse_init_kernel(CollapseState* st, uint32_t* work0, const uint32_t* root_dev, const uint32_t* cl0, const uint32_t* cl1, const PlocState* ploc)
{
    CollapseState z = {};
    z.nwork = 1u;
    uint32_t root = ploc ? 0u : *root_dev;  // radix tree: the node that ended up with the whole range (leaf 0 of a one-triangle input)
    if (ploc) {
        root = (ploc->cc ? cl1 : cl0)[0];
        if (ploc->error || ploc->n != 1u) z.error = 3u;
    }
    work0[0] = root;
    *st = z;
}

// ---------------------------------------------------------------------------------------------
// IAS
// ---------------------------------------------------------------------------------------------
__global__ void build_ias_kernel(const char* __restrict__ instances, uint32_t n, uint32_t stride, AccelHeader* h, InstanceRecord* recs)
{
    // one block; few instances on this path (one per glTF mesh node, SDK/sutil/Scene.cpp:1134-1155)
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        const b200rt_instance* in = (const b200rt_instance*)(instances + (size_t)k * stride);
        InstanceRecord r;
        for (int j = 0; j < 12; ++j) r.m[j] = in->transform[j];
        invert34(r.m, r.inv);
        r.gas = in->traversableHandle;
        r.instance_id = in->instanceId;
        r.sbt_offset = in->sbtOffset;
        r.mask = in->visibilityMask;
        r.flags = in->flags;
        r.wbounds[0] = r.wbounds[1] = 0.f;
        recs[k] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // bounds: serial pass (n is small on this path)
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        uint32_t anyhit = 0;
        for (uint32_t k = 0; k < n; ++k) {
            const AccelHeader* g = (const AccelHeader*)recs[k].gas;
            if (!g || g->magic != ACCEL_MAGIC || !g->num_tris) continue;
            // OPTIX_INSTANCE_FLAG_DISABLE_ANYHIT (1<<2) / ENFORCE_ANYHIT (1<<3) override the geometry flags of the instanced GAS
            anyhit |= (recs[k].flags & 4u) ? 0u : (recs[k].flags & 8u) ? 1u : g->anyhit;
            for (int c = 0; c < 8; ++c) {
                const float3 p = f3(g->bounds[(c & 1) ? 3 : 0], g->bounds[(c & 2) ? 4 : 1], g->bounds[(c & 4) ? 5 : 2]);
                const float3 w = xform_point(recs[k].m, p);
                lo[0] = fminf(lo[0], w.x); lo[1] = fminf(lo[1], w.y); lo[2] = fminf(lo[2], w.z);
                hi[0] = fmaxf(hi[0], w.x); hi[1] = fmaxf(hi[1], w.y); hi[2] = fmaxf(hi[2], w.z);
            }
        }
        AccelHeader v;
        memset(&v, 0, sizeof(v));
        v.magic = ACCEL_MAGIC;
        v.kind = ACCEL_KIND_IAS;
        v.num_instances = n;
        v.inst_off = HEADER_BYTES;
        v.total_bytes = HEADER_BYTES + (uint64_t)n * INSTREC_BYTES;
        v.anyhit = anyhit;
        for (int a = 0; a < 3; ++a) { v.bounds[a] = lo[a]; v.bounds[3 + a] = hi[a]; }
        *h = v;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct BuildPlan {
    uint32_t ntris = 0;
    uint32_t num_inputs = 0;
    uint32_t total_sbt = 0;
    uint32_t max_nodes = 0;
    uint32_t rs_blocks = 0;
    int morton_bits = 21;
    // temp offsets
    size_t off_inputs, off_flags, off_bounds, off_state, off_tri_tmp, off_keys0, off_keys1, off_vals0, off_vals1, off_box,
        off_parent, off_range, off_arrive, off_dest, off_hist, off_scan_tmp, off_work0, off_work1, off_child_tmp, off_counts, off_height, off_dsums;
    size_t temp_bytes = 0, out_bytes = 0;
    uint32_t node_bytes = NODE8_BYTES;
};

static uint32_t count_tris(const b200rt_build_input_triangle_array& t)
{
    return t.indexFormat != B200RT_INDICES_FORMAT_NONE && t.indexBuffer ? t.numIndexTriplets : t.numVertices / 3;
}

// Node encoding by scene size: fp32 child boxes while nodes + triangles stay within about half of the 126 MB L2 (the traversal
// is then issue-bound and skipping the 8-bit decode roughly halves a node visit), 8-bit boxes beyond (HBM-bound: 80 B per visit).
// B200RT_NODE_FORMAT=q8|f32 overrides (A/B measurements, tests of both encodings).
static uint32_t choose_node_bytes(uint64_t ntris)
{
    const char* e = getenv("B200RT_NODE_FORMAT");
    if (e && !strcmp(e, "q8")) return NODE8_BYTES;
    if (e && !strcmp(e, "f32")) return NODE8F_BYTES;
    return ntris <= 600000ull ? NODE8F_BYTES : NODE8_BYTES;
}

// Which binary hierarchy the 8-wide collapse starts from.  PLOC costs a few dozen clustering rounds (each with a host read-back of the
// cluster count) and pays where Morton order is a poor guide — few, large or unevenly sized triangles: measured (profiles/
// r01_hierarchy.md) Cornell 7.95 -> 7.29 ms per launch, the imgui_test scene (1.74 M triangles) 20.9 -> 17.0 ms per frame.  On the
// regularly tessellated 50 M-triangle bench scene the Karras tree over 63-bit Morton codes is already the better one (15.6 vs 16.8
// node visits per ray) and twice as fast to build, so big inputs keep it.  B200RT_HIERARCHY=ploc|lbvh overrides (A/B runs, tests).
constexpr uint32_t PLOC_MAX_TRIANGLES = 1u << 23;
static bool use_ploc(const b200rt_accel_build_options* options, uint32_t ntris)
{
    const char* e = getenv("B200RT_HIERARCHY");  // read per build: tests switch it
    if (e && !strcmp(e, "ploc")) return true;
    if (e && !strcmp(e, "lbvh")) return false;
    (void)options;
    return ntris <= PLOC_MAX_TRIANGLES;
}

// B200RT_MAX_WIDE_DEPTH=<n> lowers the depth the collapse aims for (tests of the depth guard; it holds whenever the binary hierarchy is
// at most 3 (n - 1) levels tall).  Never above what the traversal stack is good for.
static int max_wide_depth()
{
    const char* e = getenv("B200RT_MAX_WIDE_DEPTH");
    const int v = e ? atoi(e) : MAX_WIDE_DEPTH;
    return std::min(std::max(v, 3), MAX_WIDE_DEPTH);
}

static int make_plan(b200rt_context ctx, const b200rt_build_input* inputs, unsigned num_inputs, BuildPlan& p)
{
    uint64_t n = 0, sbt = 0;
    for (unsigned i = 0; i < num_inputs; ++i) {
        B2_REQUIRE(ctx, inputs[i].type == B200RT_BUILD_INPUT_TYPE_TRIANGLES, "build input %u: only triangle inputs can be mixed in a GAS", i);
        const auto& t = inputs[i].triangleArray;
        B2_REQUIRE(ctx, t.vertexFormat == B200RT_VERTEX_FORMAT_FLOAT3, "build input %u: vertexFormat 0x%x not supported (FLOAT3 only)", i, t.vertexFormat);
        B2_REQUIRE(ctx, t.indexFormat == B200RT_INDICES_FORMAT_NONE || t.indexFormat == B200RT_INDICES_FORMAT_UNSIGNED_SHORT3 ||
                            t.indexFormat == B200RT_INDICES_FORMAT_UNSIGNED_INT3, "build input %u: bad indexFormat 0x%x", i, t.indexFormat);
        B2_REQUIRE(ctx, t.numSbtRecords >= 1, "build input %u: numSbtRecords must be >= 1", i);
        n += count_tris(t);
        sbt += t.numSbtRecords;
    }
    B2_REQUIRE(ctx, n < (1ull << 30), "too many triangles (%llu)", (unsigned long long)n);
    B2_REQUIRE(ctx, sbt <= TRI_SBT_MASK, "too many SBT records");
    p.ntris = (uint32_t)n;
    p.num_inputs = num_inputs;
    p.total_sbt = (uint32_t)sbt;
    p.max_nodes = (uint32_t)(2 * n / 3 + 16);
    p.rs_blocks = std::max(1u, div_up(n, RS_TILE));
    // bits per axis of the Morton key = radix passes of the sort (8 bits each): 30 / 48 / 63 bits -> 4 / 6 / 8 passes.  2^16 cells per axis
    // separate the centroids of 5 * 10^7 triangles as well as 2^18 do (same node visits per ray to five digits, one pass = 0.8 ms less:
    // profiles/r02_build.md); the 63-bit key is for inputs beyond 2^27 triangles.
    p.morton_bits = n < (1u << 14) ? 10 : (n < (1u << 27) ? 16 : 21);
    if (const char* e = getenv("B200RT_MORTON_BITS")) p.morton_bits = std::min(std::max(atoi(e), 4), 21);  // A/B runs
    const size_t N = std::max<size_t>(n, 1);
    const size_t W = N / 4 + 2;  // widest possible level (every wide node roots >= 4 triangles)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    p.off_inputs = take(sizeof(DevInput) * std::max(1u, num_inputs));
    p.off_flags = take(4 * std::max<size_t>(sbt, 1));
    p.off_bounds = take(64);
    p.off_state = take(256);      // PlocState at 0, CollapseState at 128, root of the radix tree at 192
    p.off_dsums = take(8 * DS_BLOCKS);
    p.off_tri_tmp = take(48 * N);
    p.off_keys0 = take(8 * N);
    p.off_keys1 = take(8 * N);
    p.off_vals0 = take(4 * N);
    p.off_vals1 = take(4 * N);
    p.off_box = take(32 * 2 * N);
    p.off_range = take(8 * N);
    p.off_arrive = take(4 * N);
    p.off_dest = take(4 * N);
    p.off_height = take(N);
    p.off_hist = take(4 * 256 * (size_t)p.rs_blocks);
    const size_t scan_elems = scan_temp_elems<uint32_t>(256 * (size_t)p.rs_blocks);
    p.off_scan_tmp = take(8 * scan_elems);
    p.off_work0 = take(4 * W);
    p.off_work1 = take(4 * W);
    p.off_child_tmp = take(4 * 8 * W);
    p.off_counts = take(8 * W);
    p.temp_bytes = off;
    p.node_bytes = choose_node_bytes(n);
    p.out_bytes = align_up(HEADER_BYTES + (size_t)p.max_nodes * p.node_bytes + (size_t)n * TRI_BYTES, 128);
    return 0;
}

int accel_compute_memory_usage(b200rt_context ctx, const b200rt_accel_build_options* options, const b200rt_build_input* inputs,
                               unsigned num_inputs, b200rt_accel_buffer_sizes* sizes)
{
    B2_REQUIRE(ctx, options && inputs && sizes && num_inputs >= 1, "null argument");
    B2_REQUIRE(ctx, options->operation == B200RT_BUILD_OPERATION_BUILD, "only OPERATION_BUILD is supported (static scenes)");
    if (inputs[0].type == B200RT_BUILD_INPUT_TYPE_INSTANCES) {
        B2_REQUIRE(ctx, num_inputs == 1, "an instance build takes exactly one build input");
        sizes->outputSizeInBytes = align_up(HEADER_BYTES + (size_t)inputs[0].instanceArray.numInstances * INSTREC_BYTES, 128);
        sizes->tempSizeInBytes = 128;
        sizes->tempUpdateSizeInBytes = 0;
        return 0;
    }
    BuildPlan p;
    int rc = make_plan(ctx, inputs, num_inputs, p);
    if (rc) return rc;
    sizes->outputSizeInBytes = p.out_bytes;
    sizes->tempSizeInBytes = p.temp_bytes;
    sizes->tempUpdateSizeInBytes = 0;
    return 0;
}

int accel_build(b200rt_context ctx, cudaStream_t s, const b200rt_accel_build_options* options, const b200rt_build_input* inputs,
                unsigned num_inputs, b200rt_deviceptr temp, size_t temp_bytes, b200rt_deviceptr out, size_t out_bytes,
                b200rt_traversable* handle, const b200rt_accel_emit_desc* emitted, unsigned num_emitted)
{
    B2_REQUIRE(ctx, options && inputs && handle && num_inputs >= 1, "null argument");
    B2_REQUIRE(ctx, options->operation == B200RT_BUILD_OPERATION_BUILD, "only OPERATION_BUILD is supported (static scenes)");
    B2_REQUIRE(ctx, out && (out % B200RT_ACCEL_BUFFER_BYTE_ALIGNMENT) == 0, "outputBuffer must be 128-byte aligned");
    for (unsigned i = 0; i < num_emitted; ++i) {
        B2_REQUIRE(ctx, emitted && emitted[i].type == B200RT_PROPERTY_TYPE_COMPACTED_SIZE, "only COMPACTED_SIZE can be emitted");
        // same rule (and same error code) as optixAccelBuild: "Querying compacted size, but build flag ALLOW_COMPACTION is not set"
        B2_REQUIRE(ctx, options->buildFlags & B200RT_BUILD_FLAG_ALLOW_COMPACTION, "emittedProperties[%u]: compacted size queried without BUILD_FLAG_ALLOW_COMPACTION", i);
        B2_REQUIRE(ctx, emitted[i].result, "emittedProperties[%u].result is null", i);
    }
    DeviceGuard guard(ctx->device);
    retire_loops(ctx, false);
    uint64_t exact_bytes = 0;
    // what the whitted launch learnt about the scene (light count, BLEND materials) belongs to the scene that was there before:
    // a new scene always comes with a build, and its buffers may well reuse the old addresses
    ctx->w_params = 0;

    if (inputs[0].type == B200RT_BUILD_INPUT_TYPE_INSTANCES) {
        B2_REQUIRE(ctx, num_inputs == 1, "an instance build takes exactly one build input");
        const auto& ia = inputs[0].instanceArray;
        const size_t need = align_up(HEADER_BYTES + (size_t)ia.numInstances * INSTREC_BYTES, 128);
        B2_REQUIRE(ctx, out_bytes >= need, "outputBuffer too small (%zu < %zu)", out_bytes, need);
        B2_REQUIRE(ctx, ia.instances || ia.numInstances == 0, "null instances");
        build_ias_kernel<<<1, 128, 0, s>>>((const char*)ia.instances, ia.numInstances, ia.instanceStride ? ia.instanceStride : 80u,
                                           (AccelHeader*)out, (InstanceRecord*)(out + HEADER_BYTES));
        B2_LAUNCH_CHECK(ctx);
        exact_bytes = HEADER_BYTES + (uint64_t)ia.numInstances * INSTREC_BYTES;
    } else {
        BuildPlan p;
        int rc = make_plan(ctx, inputs, num_inputs, p);
        if (rc) return rc;
        B2_REQUIRE(ctx, out_bytes >= p.out_bytes, "outputBuffer too small (%zu < %zu)", out_bytes, p.out_bytes);
        B2_REQUIRE(ctx, temp_bytes >= p.temp_bytes && temp, "tempBuffer too small (%zu < %zu)", temp_bytes, p.temp_bytes);
        char* T = (char*)temp;
        const uint32_t N = p.ntris;
        // ---- upload input descriptors + geometry flags (pageable -> staged synchronously by the runtime)
        std::vector<DevInput> dev(num_inputs);
        std::vector<uint32_t> gflags(std::max(1u, p.total_sbt), 0u);
        uint32_t any_anyhit = 0;
        uint32_t tri_start = 0, sbt_base = 0;
        for (unsigned i = 0; i < num_inputs; ++i) {
            const auto& t = inputs[i].triangleArray;
            DevInput& d = dev[i];
            memset(&d, 0, sizeof(d));
            // an input without triangles may come with a null buffer (an empty cudaMalloc / tensor); one with triangles may not
            B2_REQUIRE(ctx, count_tris(t) == 0 || (t.vertexBuffers && t.vertexBuffers[0]), "build input %u: null vertex buffer", i);
            d.verts = t.vertexBuffers ? (const char*)t.vertexBuffers[0] : nullptr;
            d.vstride = t.vertexStrideInBytes ? t.vertexStrideInBytes : 12u;
            const bool indexed = t.indexFormat != B200RT_INDICES_FORMAT_NONE && t.indexBuffer;
            d.indices = indexed ? (const char*)t.indexBuffer : nullptr;
            d.iformat = !indexed ? 0u : (t.indexFormat == B200RT_INDICES_FORMAT_UNSIGNED_SHORT3 ? 2u : 4u);
            d.istride = t.indexStrideInBytes ? t.indexStrideInBytes : 3u * d.iformat;
            d.xform = (t.transformFormat == B200RT_TRANSFORM_FORMAT_MATRIX_FLOAT12) ? (const float*)t.preTransform : nullptr;
            d.sbt_index = (const char*)t.sbtIndexOffsetBuffer;
            d.sbt_size = t.sbtIndexOffsetSizeInBytes ? t.sbtIndexOffsetSizeInBytes : 4u;
            d.sbt_stride = t.sbtIndexOffsetStrideInBytes ? t.sbtIndexOffsetStrideInBytes : d.sbt_size;
            d.prim_offset = t.primitiveIndexOffset;
            d.sbt_base = sbt_base;
            d.num_sbt = t.numSbtRecords;
            d.tri_start = tri_start;
            d.ntris = count_tris(t);
            d.flags_off = sbt_base;
            for (unsigned k = 0; k < t.numSbtRecords; ++k) {
                gflags[sbt_base + k] = t.flags ? t.flags[k] : 0u;
                if (!(gflags[sbt_base + k] & 1u)) any_anyhit = 1u;  // OPTIX_GEOMETRY_FLAG_DISABLE_ANYHIT not set
            }
            tri_start += d.ntris;
            sbt_base += t.numSbtRecords;
        }
        B2_CUDA(ctx, cudaMemcpyAsync(T + p.off_inputs, dev.data(), sizeof(DevInput) * num_inputs, cudaMemcpyHostToDevice, s));
        B2_CUDA(ctx, cudaMemcpyAsync(T + p.off_flags, gflags.data(), 4 * gflags.size(), cudaMemcpyHostToDevice, s));

        uint32_t* d_bounds = (uint32_t*)(T + p.off_bounds);
        float4* tri_tmp = (float4*)(T + p.off_tri_tmp);
        uint64_t* keys[2] = {(uint64_t*)(T + p.off_keys0), (uint64_t*)(T + p.off_keys1)};
        uint32_t* vals[2] = {(uint32_t*)(T + p.off_vals0), (uint32_t*)(T + p.off_vals1)};
        const BoxArr box_lo{(float4*)(T + p.off_box)}, box_hi{(float4*)(T + p.off_box) + 1};   // interleaved: see BoxArr
        uint32_t* d_root = (uint32_t*)(T + p.off_state + 192);
        int2* range = (int2*)(T + p.off_range);
        uint32_t* arrive = (uint32_t*)(T + p.off_arrive);
        uint32_t* dest = (uint32_t*)(T + p.off_dest);
        uint32_t* hist = (uint32_t*)(T + p.off_hist);
        void* scan_tmp = (void*)(T + p.off_scan_tmp);
        uint32_t* work[2] = {(uint32_t*)(T + p.off_work0), (uint32_t*)(T + p.off_work1)};
        int* child_tmp = (int*)(T + p.off_child_tmp);
        unsigned long long* counts = (unsigned long long*)(T + p.off_counts);
        uint8_t* height = (uint8_t*)(T + p.off_height);
        PlocState* d_ploc = (PlocState*)(T + p.off_state);
        CollapseState* d_cst = (CollapseState*)(T + p.off_state + 128);
        unsigned long long* dsums = (unsigned long long*)(T + p.off_dsums);

        AccelHeader hv;
        memset(&hv, 0, sizeof(hv));
        hv.magic = ACCEL_MAGIC;
        hv.kind = ACCEL_KIND_GAS;
        hv.num_tris = N;
        hv.nodes_off = HEADER_BYTES;
        hv.max_nodes = p.max_nodes;
        uint4* nodes_out = (uint4*)(out + HEADER_BYTES);

        init_bounds_kernel<<<1, 32, 0, s>>>(d_bounds);
        B2_LAUNCH_CHECK(ctx);
        // B200RT_BUILD_TIMING=1: CUDA events between the phases, read (with a synchronisation) and logged at the end of the call
        static const bool phase_timing = [] { const char* e = getenv("B200RT_BUILD_TIMING"); return e && atoi(e) != 0; }();
        cudaEvent_t pev[8];
        int npev = 0;
        auto mark = [&]() { if (phase_timing && npev < 8) { cudaEventCreate(&pev[npev]); cudaEventRecord(pev[npev], s); ++npev; } };
        mark();
        if (N > 0) {
            gather_tris_kernel<<<persistent_grid(ctx, N, 256, 8), 256, 0, s>>>((const DevInput*)(T + p.off_inputs), (int)num_inputs, N,
                                                              (const uint32_t*)(T + p.off_flags), tri_tmp, d_bounds);
            B2_LAUNCH_CHECK(ctx);
            morton_kernel<<<div_up(N, 256), 256, 0, s>>>(tri_tmp, N, d_bounds, p.morton_bits, keys[0], vals[0]);
            B2_LAUNCH_CHECK(ctx);
            int cur = 0;
            if ((rc = rs_scatter_prepare<uint64_t>(ctx))) return rc;
            const int passes = (3 * p.morton_bits + 7) / 8;
            for (int pass = 0; pass < passes && N > 1; ++pass) {
                rs_hist_kernel<uint64_t><<<p.rs_blocks, RS_THREADS, 0, s>>>(keys[cur], N, pass * 8, hist, p.rs_blocks);
                B2_LAUNCH_CHECK(ctx);
                rc = exclusive_scan<uint32_t>(ctx, hist, 256 * (size_t)p.rs_blocks, (uint32_t*)scan_tmp, s);
                if (rc) return rc;
                rs_scatter_kernel<uint64_t><<<p.rs_blocks, RS_THREADS, rs_scatter_smem<uint64_t>(), s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], N, pass * 8, hist,
                                                                               p.rs_blocks);
                B2_LAUNCH_CHECK(ctx);
                cur ^= 1;
            }
            mark();  // gather + Morton + sort
            // No host read-back from here on: optixAccelBuild is asynchronous, so the rounds of the clustering and the levels of
            // the collapse, whose counts only the device knows, run as device-side loops (CUDA graphs with a conditional WHILE node).
            const unsigned wide_grid = std::max(1u, std::min(div_up(N, 128), (unsigned)ctx->sm_count * 16u));
            const bool ploc = use_ploc(options, N) && N > 1;
            uint32_t* cl[2] = {nullptr, nullptr};
            if (ploc) {
                leaf_boxes_kernel<<<div_up(N, 256), 256, 0, s>>>(tri_tmp, vals[cur], (int)N, box_lo, box_hi);
                B2_LAUNCH_CHECK(ctx);
                // the sort's key buffers are free now: cluster ping-pong in one, the scan words in the other
                cl[0] = (uint32_t*)keys[cur]; cl[1] = (uint32_t*)keys[cur] + N;
                unsigned long long* flags = (unsigned long long*)keys[cur ^ 1];
                uint32_t* nearest = arrive;
                ploc_init_kernel<<<div_up(N, 256), 256, 0, s>>>(cl[0], N, d_ploc);
                B2_LAUNCH_CHECK(ctx);
                const unsigned pgrid = std::max(1u, std::min(div_up(N, 256), (unsigned)ctx->sm_count * 8u));
                LoopGraph& g = *new LoopGraph(ctx);
                park_loop(ctx, &g);
                if ((rc = g.begin())) return rc;
                if ((rc = g.add((const void*)ploc_nearest_kernel, pgrid, PLOC_THREADS, 0, (const uint32_t*)cl[0], (const uint32_t*)cl[1], (const PlocState*)d_ploc,
                                box_lo, box_hi, nearest))) return rc;
                if ((rc = g.add((const void*)ploc_flag_kernel, pgrid, 256, 0, (const uint32_t*)nearest, (const PlocState*)d_ploc, flags))) return rc;
                if ((rc = dscan_add<unsigned long long>(g, flags, &d_ploc->n, dsums, &d_ploc->total))) return rc;
                if ((rc = g.add((const void*)ploc_merge_kernel, pgrid, 256, 0, cl[0], cl[1], (const uint32_t*)nearest, (const PlocState*)d_ploc,
                                (const unsigned long long*)flags, (int)N - 1, box_lo, box_hi, height))) return rc;
                if ((rc = g.add((const void*)ploc_advance_kernel, 1, 1, 0, d_ploc, g.cond()))) return rc;
                if ((rc = g.launch(s))) return rc;
                ctx->launches += 1;
            } else {
                B2_CUDA(ctx, cudaMemsetAsync(arrive, 0, 4 * (size_t)N, s));
                radix_tree_kernel<<<div_up(N, 256), 256, 0, s>>>(keys[cur], tri_tmp, vals[cur], (int)N, box_lo, box_hi, range, arrive, height, d_root);
                B2_LAUNCH_CHECK(ctx);
            }
            mark();  // leaf boxes + binary hierarchy
            collapse_init_kernel<<<1, 1, 0, s>>>(d_cst, work[0], d_root, cl[0], cl[1], ploc ? d_ploc : nullptr);
            B2_LAUNCH_CHECK(ctx);
            // ---- collapse, level by level
            static const bool body_timing = [] { const char* e = getenv("B200RT_BUILD_TIMING"); return e && atoi(e) >= 2; }();
            unsigned long long* d_stamps = nullptr;
            if (body_timing) {
                cudaMalloc(&d_stamps, 8 * sizeof(unsigned long long));
                cudaMemsetAsync(d_stamps, 0, 8 * sizeof(unsigned long long), s);
                stamp_kernel<<<1, 1, 0, s>>>(d_stamps, -1);
            }
            {
                LoopGraph& g = *new LoopGraph(ctx);
                park_loop(ctx, &g);
                if ((rc = g.begin())) return rc;
                if ((rc = g.add((const void*)collapse_plan_kernel, wide_grid, 128, 0, (const CollapseState*)d_cst, (const uint32_t*)work[0], (const uint32_t*)work[1], (int)N,
                                box_lo, box_hi, (const uint8_t*)height, child_tmp, counts, max_wide_depth()))) return rc;
                if (body_timing && (rc = g.add((const void*)stamp_kernel, 1, 1, 0, d_stamps, 0))) return rc;
                if ((rc = dscan_add<unsigned long long>(g, counts, &d_cst->nwork, dsums, &d_cst->total))) return rc;
                if (body_timing && (rc = g.add((const void*)stamp_kernel, 1, 1, 0, d_stamps, 1))) return rc;
                const unsigned emit_grid = std::max(1u, std::min(div_up(N, 64), (unsigned)ctx->sm_count * 8u));  // 16 nodes per CTA and sweep
                if ((rc = g.add((const void*)collapse_emit_kernel, emit_grid, 128, 0, d_cst, p.max_nodes, (int)N, box_lo, box_hi, ploc ? 1 : 0,
                                (const int*)child_tmp, (const unsigned long long*)counts, work[0], work[1], dest, nodes_out, p.node_bytes))) return rc;
                if (body_timing && (rc = g.add((const void*)stamp_kernel, 1, 1, 0, d_stamps, 2))) return rc;
                if ((rc = g.add((const void*)collapse_advance_kernel, 1, 1, 0, d_cst, p.max_nodes, g.cond()))) return rc;
                if (body_timing && (rc = g.add((const void*)stamp_kernel, 1, 1, 0, d_stamps, 3))) return rc;
                if ((rc = g.launch(s))) return rc;
                ctx->launches += 1;
            }
            if (body_timing) {
                unsigned long long h[8];
                cudaMemcpy(h, d_stamps, sizeof(h), cudaMemcpyDeviceToHost);
                cudaFree(d_stamps);
                fprintf(stderr, "[b200rt build] collapse body: plan %.3f ms, scan %.3f ms, emit %.3f ms, advance + loop %.3f ms\n", h[1] * 1e-6, h[2] * 1e-6, h[3] * 1e-6,
                        h[4] * 1e-6);
            }
            mark();  // collapse
            // triangles go right after the node capacity region; compaction later closes the gap
            hv.tris_off = HEADER_BYTES + (uint64_t)p.max_nodes * p.node_bytes;
            scatter_tris_kernel<<<div_up(N, 256), 256, 0, s>>>(tri_tmp, vals[cur], dest, N, (float4*)(out + hv.tris_off));
            B2_LAUNCH_CHECK(ctx);
        } else {
            hv.tris_off = HEADER_BYTES;
        }
        hv.node_bytes = p.node_bytes;
        hv.anyhit = any_anyhit;
        B2_REQUIRE(ctx, num_emitted <= 2, "at most two emitted properties");
        write_header_kernel<<<1, 1, 0, s>>>((AccelHeader*)out, hv, d_bounds, N > 0 ? d_cst : nullptr, num_emitted > 0 ? (unsigned long long*)emitted[0].result : nullptr,
                                            num_emitted > 1 ? (unsigned long long*)emitted[1].result : nullptr);
        B2_LAUNCH_CHECK(ctx);
        mark();
        if (phase_timing && npev == 5) {
            cudaEventSynchronize(pev[4]);
            float t[4];
            for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], pev[i], pev[i + 1]);
            fprintf(stderr, "[b200rt build] %u triangles: keys+sort %.3f ms, hierarchy %.3f ms, collapse %.3f ms, triangles+header %.3f ms\n", N, t[0], t[1], t[2], t[3]);
        }
        for (int i = 0; i < npev; ++i) cudaEventDestroy(pev[i]);
        log_msg(ctx, 4, "accel", "GAS build enqueued: %u triangles, node capacity %u, %u-byte nodes", N, p.max_nodes, p.node_bytes);
        *handle = out;
        return 0;
    }
    for (unsigned i = 0; i < num_emitted; ++i) {  // instance acceleration structures: the size is known here
        const size_t v = align_up(exact_bytes, 128);
        B2_CUDA(ctx, cudaMemcpyAsync((void*)emitted[i].result, &v, sizeof(size_t), cudaMemcpyHostToDevice, s));
    }
    *handle = out;
    return 0;
}

// optixAccelCompact (optixPathTracer.cpp:671-683) is asynchronous like the build: the sizes are in the header, which only the device
// has, so one grid-stride kernel reads it there, copies the used part of the node array and the triangle records behind it, and
// writes the patched header.  A blob that is not a b200rt traversable, or an output buffer that is too small, leaves a header with the
// error field set and no handle contents (error 4 / 5); accel_get_info reports it.
__global__ void __launch_bounds__(256) compact_kernel(const char* __restrict__ in, char* __restrict__ out, size_t out_bytes)
{
    const AccelHeader h = *(const AccelHeader*)in;
    const bool ok = h.magic == ACCEL_MAGIC && out_bytes >= h.total_bytes;
    if (!ok) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && out_bytes >= HEADER_BYTES) {
            AccelHeader e = {};
            e.magic = ACCEL_MAGIC; e.kind = ACCEL_KIND_GAS; e.nodes_off = HEADER_BYTES; e.tris_off = HEADER_BYTES; e.total_bytes = HEADER_BYTES;
            e.node_bytes = NODE8_BYTES; e.error = h.magic == ACCEL_MAGIC ? 5u : 4u;
            *(AccelHeader*)out = e;
        }
        return;
    }
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    if (h.kind == ACCEL_KIND_IAS) {
        const size_t n16 = (size_t)(h.total_bytes / 16);
        for (size_t i = tid; i < n16; i += nthreads) ((uint4*)out)[i] = __ldg((const uint4*)in + i);
        return;
    }
    const size_t node_bytes = (size_t)h.num_nodes * (h.node_bytes ? h.node_bytes : NODE8_BYTES);
    const size_t nn16 = node_bytes / 16, nt16 = (size_t)h.num_tris * (TRI_BYTES / 16);
    const uint4* src_nodes = (const uint4*)(in + HEADER_BYTES);
    const uint4* src_tris = (const uint4*)(in + h.tris_off);
    uint4* dst_nodes = (uint4*)(out + HEADER_BYTES);
    uint4* dst_tris = (uint4*)(out + HEADER_BYTES + node_bytes);
    for (size_t i = tid; i < nn16; i += nthreads) dst_nodes[i] = __ldcs(src_nodes + i);
    for (size_t i = tid; i < nt16; i += nthreads) dst_tris[i] = __ldcs(src_tris + i);
    if (tid == 0) {
        AccelHeader v = h;
        v.tris_off = HEADER_BYTES + node_bytes;
        v.max_nodes = h.num_nodes;
        *(AccelHeader*)out = v;
    }
}

int accel_compact(b200rt_context ctx, cudaStream_t s, b200rt_traversable input, b200rt_deviceptr out, size_t out_bytes,
                  b200rt_traversable* handle)
{
    B2_REQUIRE(ctx, input && out && handle, "null argument");
    B2_REQUIRE(ctx, (out % B200RT_ACCEL_BUFFER_BYTE_ALIGNMENT) == 0, "outputBuffer must be 128-byte aligned");
    B2_REQUIRE(ctx, out_bytes >= HEADER_BYTES, "outputBuffer too small (%zu bytes)", out_bytes);
    DeviceGuard guard(ctx->device);
    compact_kernel<<<(unsigned)ctx->sm_count * 8u, 256, 0, s>>>((const char*)input, (char*)out, out_bytes);
    B2_LAUNCH_CHECK(ctx);
    *handle = out;
    return 0;
}

// optixAccelEmitProperty: the compacted size (or the bounds) of a finished acceleration structure, written to device memory by a
// one-thread kernel that reads the header where it lives
__global__ void emit_property_kernel(const AccelHeader* h, unsigned type, void* result)
{
    if (type == B200RT_PROPERTY_TYPE_COMPACTED_SIZE) *(unsigned long long*)result = (h->total_bytes + 127ull) / 128ull * 128ull;
    else { float* b = (float*)result; for (int a = 0; a < 6; ++a) b[a] = h->bounds[a]; }  // OPTIX_PROPERTY_TYPE_AABBS: one OptixAabb
}

int accel_emit_property(b200rt_context ctx, cudaStream_t s, b200rt_traversable handle, const b200rt_accel_emit_desc* emitted, unsigned num_emitted)
{
    B2_REQUIRE(ctx, handle && (emitted || num_emitted == 0), "null argument");
    DeviceGuard guard(ctx->device);
    for (unsigned i = 0; i < num_emitted; ++i) {
        B2_REQUIRE(ctx, emitted[i].result && (emitted[i].type == B200RT_PROPERTY_TYPE_COMPACTED_SIZE || emitted[i].type == B200RT_PROPERTY_TYPE_AABBS),
                   "emittedProperties[%u]: COMPACTED_SIZE or AABBS with a result address", i);
        emit_property_kernel<<<1, 1, 0, s>>>((const AccelHeader*)handle, emitted[i].type, (void*)emitted[i].result);
        B2_LAUNCH_CHECK(ctx);
    }
    return 0;
}

int accel_get_info(b200rt_context ctx, b200rt_traversable handle, b200rt_accel_info* info)
{
    B2_REQUIRE(ctx, handle && info, "null argument");
    DeviceGuard guard(ctx->device);
    AccelHeader h;
    B2_CUDA(ctx, cudaMemcpy(&h, (const void*)handle, sizeof(h), cudaMemcpyDeviceToHost));
    B2_REQUIRE(ctx, h.magic == ACCEL_MAGIC, "not a b200rt traversable");
    // builds and compactions are asynchronous: what went wrong on the device is in the header
    B2_REQUIRE(ctx, h.error == 0, "the acceleration structure is invalid: %s",
               h.error == 1 ? "node capacity exceeded during the build (internal error)" : h.error == 2 ? "triangle count mismatch after the build (internal error)" :
               h.error == 3 ? "clustering failed during the build (internal error)" : h.error == 4 ? "compaction source was not a b200rt traversable" :
               h.error == 5 ? "compaction output buffer too small" : "unknown error");
    info->kind = h.kind;
    info->num_triangles = h.num_tris;
    info->num_nodes = h.num_nodes;
    info->num_instances = h.num_instances;
    info->total_bytes = h.total_bytes;
    memcpy(info->bounds, h.bounds, sizeof(h.bounds));
    info->depth = h.depth;
    info->reserved = h.kind == ACCEL_KIND_GAS ? (h.node_bytes ? h.node_bytes : NODE8_BYTES) : 0;  // bytes per wide node (80 or 224)
    return 0;
}

}  // namespace b200rt
