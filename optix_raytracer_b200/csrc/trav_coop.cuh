// trav_coop.cuh — persistent-warp traversal driver: dynamic ray fetch, batched commits, and a
// WARP-COOPERATIVE triangle phase.
//
// The per-ray traversal of traverse.cuh is re-expressed as a state machine (`Trav`).  One warp keeps its 32
// lanes busy: a lane whose ray terminates fetches the next work item from a global cursor (Aila/Laine 2009
// persistent threads; Ylitie/Karras/Laine 2017 dynamic fetch).  What is new relative to that literature is
// the triangle phase.  ncu on the first two versions (profiles/r01_trace_before.md) showed the node test
// running at 23.6 of 32 lanes but the per-lane triangle loop at 3.5 lanes — with only ~3 triangles per ray
// among ~15 nodes, a lane rarely has company when it reaches a leaf — so ~1/3 of all warp instructions were
// triangle tests at 11 % SIMD efficiency.  Here a lane that reaches a leaf *parks* its triangle group and keeps
// descending; when the warp has collected enough (ray, triangle) units they are spread over all 32 lanes
// through shared memory, tested, and min-reduced back to the owning lane:
//
//   * node step     : pop the best pending child of the current node group, fetch its 80-byte node, test the 8
//                     quantised child boxes -> new node group (+ a triangle group that is parked)
//   * triangle round: prefix-sum the parked triangle counts, owners publish (triangle index, owner lane) units,
//                     every lane tests one unit against its owner's ray constants (kept in shared memory since
//                     fetch), owners take the lexicographic minimum (t, ordinal) of their units
//   * refill        : when fewer than REFILL_THRESHOLD lanes are active, finished lanes commit their results
//                     together and fetch new items with one warp-aggregated atomicAdd
//
// Semantics are exactly those of traverse.cuh (same tri_test arithmetic, same conservative box test, closest =
// min (t, instance, ordinal)); the hit rule is independent of visit order, so results are bit-identical.
#pragma once
#include "traverse.cuh"

namespace b200rt {

#ifndef B200RT_LOOKTHROUGH
#define B200RT_LOOKTHROUGH 0
#endif
#ifndef B200RT_FETCH_CHUNK
#define B200RT_FETCH_CHUNK 128   // 0: one atomic on the global item cursor per refill
#endif
#ifndef B200RT_STACK_TOP
#define B200RT_STACK_TOP 0   // measured slower at 64 registers (the two extra live registers turn into 190 bytes of spills): see stack_push
#endif
// Triangle records of a scene that lives in HBM are read once per ray and not again soon, while nodes (the upper levels at least) are
// read again and again: such launches load triangles with the streaming (evict-first) policy so they do not push nodes out of L1 / L2
// (measured on the 50 M-triangle bench: 2033 -> 2045 Mrays/s).  Cache-resident scenes keep the default policy.
#define TRI_LOAD(p) (tri_stream ? __ldcs(p) : __ldg(p))
constexpr uint32_t NODE_BITS = 0xff000000u;
#ifndef B200RT_REFILL_THRESHOLD
#define B200RT_REFILL_THRESHOLD 24
#endif
#ifndef B200RT_TRI_TRIGGER
#define B200RT_TRI_TRIGGER 24
#endif
constexpr int REFILL_THRESHOLD = B200RT_REFILL_THRESHOLD;   // refill when fewer lanes than this hold a ray
constexpr int TRI_TRIGGER = B200RT_TRI_TRIGGER;             // run a triangle round once this many (ray, triangle) units are parked
constexpr int COOP_BLOCK = 128;        // CTA size of every kernel built on trace_persistent
constexpr int COOP_WARPS = COOP_BLOCK / 32;
#ifndef B200RT_COOP_MIN_CTAS
#define B200RT_COOP_MIN_CTAS 8
#endif
constexpr int COOP_MIN_CTAS = B200RT_COOP_MIN_CTAS;  // resident CTAs per SM the path-tracing trace kernel is compiled for (8 -> 64 registers)
// Lane slot in shared memory (floats): ox, oy, oz, Sx, Sy, Sz, pack, tmin, cull — what the lane that tests one of this ray's triangles
// needs.  Kernels whose Work sets SMEM_STATE keep more of the lane's state there: the triangle base pointer of the GAS being traversed
// (read by the triangle rounds only — by other lanes, which otherwise fetch it with two shuffles) and the primitive / SBT / instance /
// barycentrics of the best hit so far (written when a hit is accepted, read at commit).  That takes registers off kernels whose commit
// needs the whole hit record: the optixRaycasting kernel spilled 158 bytes at the 64-register cap and spills 16 with it (Duck ray
// buffers 5666 -> 6389 Mrays/s).  The path-tracing trace kernel only carries prim and sbt and does not spill; for it the larger slot
// costs occupancy through shared memory (8 CTAs x 10.7 KB no longer fit the 28 % carve-out: 2138 -> 1645 Mrays/s on the bench), so it
// stays off there.  Keeping the popped stack entry in registers on top of it (B200RT_STACK_TOP) measured slower in both settings.
constexpr int RAY_S_STRIDE_BASE = 9, RAY_S_STRIDE_SMEM = 17;  // odd strides: lanes reading different owners hit different banks
constexpr int RS_TRIS = 9, RS_PRIM = 11, RS_SBT = 12, RS_INST = 13, RS_B1 = 14, RS_B2 = 15;
template <bool SM> constexpr int ray_s_stride() { return SM ? RAY_S_STRIDE_SMEM : RAY_S_STRIDE_BASE; }

// pack word layout
constexpr uint32_t TP_ANY = 1u << 16;        // any-hit (terminate on first hit)
constexpr uint32_t TP_FOUND = 1u << 17;      // a hit was accepted in the GAS currently being traversed (tie-break scope)
constexpr uint32_t TP_FOUND_ANY = 1u << 18;  // a hit was accepted in any instance so far
constexpr uint32_t TP_F32 = 1u << 19;        // the GAS being traversed has Node8F nodes (fp32 child boxes, 224 B)
constexpr uint32_t TP_H64 = 0x64000000u;     // top byte = high byte of the fp16 patterns 0x64qq = 1024 + q the 8-bit decode builds with one PRMT

// 8-bit plane decode of the Node8 visit.  ncu's source page (profiles/r01_trace_kernel.md) showed ~50 of the 352 instructions of a
// node visit to be constant traffic: PRMT takes only one immediate, so with the constant 0x64646464 as an operand every PRMT needed
// its selector moved into a register, and every FHADD its -1024.  Here the 0x64 byte comes from the ray's live `pack` word (so the
// selector is the immediate) and the 1024 bias is folded into the plane offset once per axis (aox - 1024 aix), so the conversion
// is FHADD with RZ.  Both are exact except for the one extra rounding of the folded offset: <= 2^-24 |aox| + 2^-14 of a grid cell,
// far inside the build-time padding (2^-16 of the node extent) and the slab slack.
__device__ __forceinline__ uint32_t q8_pair01(uint32_t w, uint32_t pack) { return __byte_perm(w, pack, 0x7170u); }  // {1024 + b0, 1024 + b1} as half2 bits
__device__ __forceinline__ uint32_t q8_pair23(uint32_t w, uint32_t pack) { return __byte_perm(w, pack, 0x7372u); }  // {1024 + b2, 1024 + b3}
__device__ __forceinline__ float q8_lo(uint32_t h)
{
    float f;
    asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(f) : "h"((unsigned short)(h & 0xffffu)), "f"(0.0f));
    return f;
}
__device__ __forceinline__ float q8_hi(uint32_t h)
{
    float f;
    asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(f) : "h"((unsigned short)(h >> 16)), "f"(0.0f));
    return f;
}

// L2 eviction priorities for the node fetches of a scene that lives in HBM (B200RT_NODE_KEEP_MB > 0): nodes are laid out level by
// level, so the first megabytes of the node array are the upper levels of the tree.  A range policy (createpolicy.range: one 64-bit
// descriptor, held in uniform registers) gives them evict_last and the deep levels — hundreds of megabytes read once per ray that
// reaches them — evict_first, so that the deep levels stop pushing the middle of the tree out of L2.
#ifndef B200RT_NODE_KEEP_MB
#define B200RT_NODE_KEEP_MB 0
#endif
__device__ __forceinline__ uint64_t node_policy_for(const AccelHeader* __restrict__ h)
{
    uint64_t pol;
    if (h->kind == ACCEL_KIND_GAS && h->node_bytes == NODE8_BYTES) {
        const char* base = (const char*)h + h->nodes_off;
        asm("createpolicy.range.global.L2::evict_last.L2::evict_first.b64 %0, [%1], %2, %3;"
            : "=l"(pol) : "l"(base), "r"((uint32_t)B200RT_NODE_KEEP_MB << 20), "r"(0xffffff00u));
    } else {
        asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    }
    return pol;
}
template <bool POLICY>
__device__ __forceinline__ uint4 node_load(const uint4* p, uint64_t pol)
{
    if constexpr (POLICY) {
        uint4 v;
        asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
        return v;
    } else {
        return __ldg(p);
    }
}

// Prefetches into L2 for the scene that lives in HBM (Node8 hierarchies): see trav_node_step_q8 and the parking of triangle groups.
#ifndef B200RT_SIB_PREFETCH
#define B200RT_SIB_PREFETCH 0   // sectors asked for per deferred sibling node: 1 = first, 2 = first and last, 3 = all three of the 80 bytes
#endif
#ifndef B200RT_TRI_PREFETCH
#define B200RT_TRI_PREFETCH 0   // 1: the first and the last record of a parked triangle group; 2: every 64 bytes of the span (at most 4)
#endif
__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct Trav {
    const uint4* nodes;
    const float4* tris;     // unused (and dead) in kernels that keep it in the lane's shared slot (SMEM_STATE)
    float ox, oy, oz;       // origin in the space of the GAS being traversed
    float idx, idy, idz;    // reciprocal (clamped) direction for the box tests
    float tmin;
    uint32_t pack;          // kx | ky<<2 | kz<<4 | octinv<<8 | negx<<11 | negy<<12 | negz<<13 | TP_* | 0x64 << 24 (fp16 exponent byte, see q8_pair*)
    uint32_t inst;          // index of the instance being traversed (0 for a bare GAS)
    uint2 ngroup, tgroup;   // current node group; parked triangle group (base, 24-bit mask)
    int sp;
#if B200RT_STACK_TOP
    uint2 top;              // the top stack entry (index sp - 1) lives here; stack[0 .. sp-2] are in local memory
#endif
    RayHit best;
};

// Traversal stack.  ncu's source page (profiles/r01_trace_kernel.md, v5) charged 10.7 % of the kernel's stall samples to the line that
// looks at the entry just popped: the pop is a local-memory load (a miss in the node-filled L1 more often than not) whose value the very
// next instructions need.  With B200RT_STACK_TOP the top entry is kept in registers: a pop hands it over at once and issues the load of
// the entry below, which has the whole next node visit to arrive; a push stores the old top and keeps the new one.  Measured on the
// bench: 2083 -> 1931 Mrays/s — under the 64-register cap of 8 CTAs per SM the kernel spills 190 instead of 10 bytes per thread, which
// costs more than the exposed pop.  Off by default; the helpers keep both forms.
// The stack itself: entries [0, NS) live in shared memory (one column per thread, rows COOP_BLOCK apart: conflict-free 64-bit accesses),
// the rest in local memory.  NS = 0: all of it in local memory.  A pop from local memory goes through the node-filled L1 and misses it
// more often than not; a pop from shared memory takes a fixed ~25 cycles and leaves the L1 to the nodes (Work::SMEM_STACK).
template <int NS>
struct TStack {
    uint2 loc[TRAV_STACK - NS];
    uint2* shm;
    __device__ __forceinline__ void put(int i, uint2 e)
    {
        if (NS > 0 && i < NS) shm[i * COOP_BLOCK] = e;
        else loc[i - NS] = e;
    }
    __device__ __forceinline__ uint2 get(int i) const
    {
        if (NS > 0 && i < NS) return shm[i * COOP_BLOCK];
        return loc[i - NS];
    }
};
template <int NS>
__device__ __forceinline__ void stack_push(Trav& s, TStack<NS>& stack, uint2 e)
{
#if B200RT_STACK_TOP
    if (s.sp > 0) stack.put(s.sp - 1, s.top);
    s.top = e;
    ++s.sp;
#else
    stack.put(s.sp++, e);
#endif
}
template <int NS>
__device__ __forceinline__ uint2 stack_peek(const Trav& s, const TStack<NS>& stack)
{
#if B200RT_STACK_TOP
    return s.top;
#else
    return stack.get(s.sp - 1);
#endif
}
template <int NS>
__device__ __forceinline__ void stack_drop(Trav& s, const TStack<NS>& stack)  // remove the top entry
{
    --s.sp;
#if B200RT_STACK_TOP
    if (s.sp > 0) s.top = stack.get(s.sp - 1);
#endif
}

template <bool SM>
__device__ __forceinline__ const float4* lane_tris(const Trav& s, const float* __restrict__ slot)
{
    if constexpr (SM) return (const float4*)(uintptr_t)(((unsigned long long)__float_as_uint(slot[RS_TRIS + 1]) << 32) | __float_as_uint(slot[RS_TRIS]));
    else return s.tris;
}
template <bool SM>
__device__ __forceinline__ void set_lane_tris(Trav& s, float* __restrict__ slot, const float4* p)
{
    if constexpr (SM) {
        slot[RS_TRIS] = __uint_as_float((uint32_t)(uintptr_t)p);
        slot[RS_TRIS + 1] = __uint_as_float((uint32_t)((unsigned long long)(uintptr_t)p >> 32));
    } else s.tris = p;
}
// best-hit fields that only commit reads
template <bool SM>
__device__ __forceinline__ void store_hit(Trav& s, float* __restrict__ slot, float b1, float b2, uint32_t prim, uint32_t sbt)
{
    if constexpr (SM) {
        slot[RS_B1] = b1; slot[RS_B2] = b2; slot[RS_PRIM] = __uint_as_float(prim); slot[RS_SBT] = __uint_as_float(sbt); slot[RS_INST] = __uint_as_float(s.inst);
    } else { s.best.b1 = b1; s.best.b2 = b2; s.best.prim = prim; s.best.sbt = sbt; s.best.inst = s.inst; }
}
template <bool SM>
__device__ __forceinline__ void load_hit(Trav& s, const float* __restrict__ slot)
{
    if constexpr (SM) {
        s.best.b1 = slot[RS_B1]; s.best.b2 = slot[RS_B2]; s.best.prim = __float_as_uint(slot[RS_PRIM]); s.best.sbt = __float_as_uint(slot[RS_SBT]);
        s.best.inst = __float_as_uint(slot[RS_INST]);
    }
}

template <int STRIDE>
struct CoopShared {
    float ray[COOP_WARPS][32 * STRIDE];  // per lane: ox, oy, oz, Sx, Sy, Sz, pack, tmin, cull (+ triangle base, best-hit fields: SMEM_STATE)
    uint32_t unit_tri[COOP_WARPS][32];
    uint32_t unit_owner[COOP_WARPS][32];
    float res_t[COOP_WARPS][32];
    uint32_t res_ord[COOP_WARPS][32];
};
template <bool ANYHIT> struct CoopSharedAnyHit {};
template <> struct CoopSharedAnyHit<true> {
    float res_fac[COOP_WARPS][32];  // factor the unit's any-hit program returned for the owner's pending attenuation
};

// Does the ray reach the (padded) bounds of the traversable at all?  For ray-generation kernels that finish the rays passing the scene
// by on the spot (one thread per ray) instead of making them work items of the persistent traversal.  Conservative like every box test
// here: the pad (2^-14 of the largest extent; an IAS's bounds are the box of its instances' transformed corners, one more rounding than
// a GAS's) and the slack keep every ray the triangle test could accept.
__device__ __forceinline__ bool ray_reaches_bounds(const AccelHeader* __restrict__ h, float3 o, float3 d, float tmin, float tmax)
{
    if (h->kind == ACCEL_KIND_IAS ? h->num_instances == 0u : h->num_tris == 0u) return false;
    const float bx = fabsf(d.x) < DIR_EPS ? copysignf(DIR_EPS, d.x) : d.x;
    const float by = fabsf(d.y) < DIR_EPS ? copysignf(DIR_EPS, d.y) : d.y;
    const float bz = fabsf(d.z) < DIR_EPS ? copysignf(DIR_EPS, d.z) : d.z;
    const float idx = fdiv(1.0f, bx), idy = fdiv(1.0f, by), idz = fdiv(1.0f, bz);
    const float lx = h->bounds[0], ly = h->bounds[1], lz = h->bounds[2], hx = h->bounds[3], hy = h->bounds[4], hz = h->bounds[5];
    // padded with directed rounding: far from the origin the pad is below the spacing of the coordinates and must still move the plane
    const float pad = fmaxf(fmaxf(hx - lx, hy - ly), hz - lz) * 6.103515625e-05f;
    const float ax = (__fsub_rd(lx, pad) - o.x) * idx, cx = (__fadd_ru(hx, pad) - o.x) * idx;
    const float ay = (__fsub_rd(ly, pad) - o.y) * idy, cy = (__fadd_ru(hy, pad) - o.y) * idy;
    const float az = (__fsub_rd(lz, pad) - o.z) * idz, cz = (__fadd_ru(hz, pad) - o.z) * idz;
    const float tn = fmaxf(fmaxf(fminf(ax, cx), fminf(ay, cy)), fmaxf(fminf(az, cz), tmin));
    const float tf = fminf(fminf(fmaxf(ax, cx), fmaxf(ay, cy)), fminf(fmaxf(az, cz), tmax));
    return tn <= tf * BOX_SLACK;
}

// Prepare the per-ray constants for a GAS (object-space origin/direction).  best.t must hold tmax.
// BOUNDS: first test the ray against the GAS bounds and return false when it passes them by — for launches where most rays miss the
// scene (camera rays around a model) this replaces the ray set-up (three more IEEE divisions), the root fetch and a full node visit by
// some twenty instructions.  The box is padded like every node box of the BVH (2^-16 of the largest extent) and compared with the same
// slack, so it is exactly as conservative as the traversal it stands in for: the hit rule (traverse.cuh) still decides alone.
template <bool BOUNDS = false, bool SM = false>
__device__ __forceinline__ bool trav_begin(Trav& s, float* __restrict__ my_ray, const AccelHeader* __restrict__ gas, float3 o, float3 d, float tmin,
                                           uint32_t keep_flags, uint32_t cull)
{
    const float bx = fabsf(d.x) < DIR_EPS ? copysignf(DIR_EPS, d.x) : d.x;
    const float by = fabsf(d.y) < DIR_EPS ? copysignf(DIR_EPS, d.y) : d.y;
    const float bz = fabsf(d.z) < DIR_EPS ? copysignf(DIR_EPS, d.z) : d.z;
    s.idx = fdiv(1.0f, bx); s.idy = fdiv(1.0f, by); s.idz = fdiv(1.0f, bz);
    if (BOUNDS) {
        const float lx = gas->bounds[0], ly = gas->bounds[1], lz = gas->bounds[2], hx = gas->bounds[3], hy = gas->bounds[4], hz = gas->bounds[5];
        const float pad = fmaxf(fmaxf(hx - lx, hy - ly), hz - lz) * 1.52587890625e-05f;
        const float ax = (__fsub_rd(lx, pad) - o.x) * s.idx, cx = (__fadd_ru(hx, pad) - o.x) * s.idx;
        const float ay = (__fsub_rd(ly, pad) - o.y) * s.idy, cy = (__fadd_ru(hy, pad) - o.y) * s.idy;
        const float az = (__fsub_rd(lz, pad) - o.z) * s.idz, cz = (__fadd_ru(hz, pad) - o.z) * s.idz;
        const float tn = fmaxf(fmaxf(fminf(ax, cx), fminf(ay, cy)), fmaxf(fminf(az, cz), tmin));
        const float tf = fminf(fminf(fmaxf(ax, cx), fmaxf(ay, cy)), fminf(fmaxf(az, cz), s.best.t));
        if (!(tn <= tf * BOX_SLACK)) return false;
    }
    const char* base = (const char*)gas;
    s.nodes = (const uint4*)(base + gas->nodes_off);
    set_lane_tris<SM>(s, my_ray, (const float4*)(base + gas->tris_off));
    const TriRay tr = make_tri_ray(o, d);
    s.ox = o.x; s.oy = o.y; s.oz = o.z;
    const uint32_t nx = bx < 0.0f, ny = by < 0.0f, nz = bz < 0.0f;
    const uint32_t oct = (nx << 2) | (ny << 1) | nz;
    s.pack = (uint32_t)tr.kx | ((uint32_t)tr.ky << 2) | ((uint32_t)tr.kz << 4) | ((7u - oct) << 8) | (nx << 11) | (ny << 12) | (nz << 13) | keep_flags | TP_H64 |
             (gas->node_bytes == NODE8F_BYTES ? TP_F32 : 0u);
    s.tmin = tmin;
    s.ngroup = gas->num_tris ? make_uint2(0u, 0x80000000u) : make_uint2(0u, 0u);
    s.tgroup = make_uint2(0u, 0u);
    s.sp = 0;
    // ray constants of the triangle test, read by whichever lane tests this ray's triangles
    my_ray[0] = o.x; my_ray[1] = o.y; my_ray[2] = o.z;
    my_ray[3] = tr.Sx; my_ray[4] = tr.Sy; my_ray[5] = tr.Sz;
    my_ray[6] = __uint_as_float(s.pack & 0x3fu);
    my_ray[7] = tmin;
    my_ray[8] = __uint_as_float(cull);
    return true;
}

// Set up traversal of `h` (GAS, or the first usable instance >= first_inst of an IAS) for the world-space ray.
// keep = TP_* bits to carry over; cull = the ray's OptixRayFlags (the CULL_* and DISABLE/ENFORCE_ANYHIT bits are used).
// Returns false when there is nothing (more) to traverse.
template <bool BOUNDS = false, bool SM = false>
__device__ __forceinline__ bool trav_begin_handle(Trav& s, float* __restrict__ my_ray, const AccelHeader* __restrict__ h, float3 o, float3 d,
                                                  float tmin, uint32_t keep, uint32_t cull, uint32_t first_inst)
{
    if (h->kind == ACCEL_KIND_GAS) {
        if (first_inst > 0u) return false;
        s.inst = 0u;
        if (!trav_begin<BOUNDS, SM>(s, my_ray, h, o, d, tmin, keep, cull_word(cull, 0u))) { s.pack = keep; return false; }
        return true;
    }
    const InstanceRecord* recs = (const InstanceRecord*)((const char*)h + h->inst_off);
    const uint32_t n = h->num_instances;
    for (uint32_t k = first_inst; k < n; ++k) {
        const InstanceRecord* ir = recs + k;
        if (!(ir->mask & ray_visibility(cull))) continue;
        s.inst = k;
        const uint32_t c = cull_word(cull, ir->flags);  // the instance's face-culling / facing / any-hit flags
        // instances are tested at their bounds always: with several of them a ray passes most of them by
        if (trav_begin<true, SM>(s, my_ray, (const AccelHeader*)ir->gas, xform_point(ir->inv, o), xform_vec(ir->inv, d), tmin, keep, c)) return true;
    }
    s.pack = keep;  // nothing (more) to traverse: the flags carried so far are what commit sees
    return false;
}

// One node visit; returns the triangle group of the visited node (mask 0 = none).
template <bool POLICY = false, int NS = 0>
__device__ __forceinline__ uint2 trav_node_step_q8(Trav& s, TStack<NS>& stack, TravStats* st, uint64_t pol = 0)
{
    const uint32_t hits_imask = s.ngroup.y;
    const uint32_t bit = 31u - __clz(hits_imask);
    const uint32_t child_base = s.ngroup.x;
    s.ngroup.y &= ~(1u << bit);
    const uint32_t octinv = (s.pack >> 8) & 7u;
    if (s.ngroup.y & NODE_BITS) {
        if (s.sp < TRAV_STACK) stack_push(s, stack, s.ngroup);
#if B200RT_SIB_PREFETCH
        // The group's next child WILL be visited (a popped group's best child is fetched without another look at its distance), but only
        // after the whole subtree entered now: its node is asked into L2 here, a subtree's time ahead of the fetch.
        const uint32_t bit2 = 31u - __clz(s.ngroup.y);
        const uint32_t slot2 = (bit2 - 24u) ^ octinv;
        const uint32_t rel2 = __popc(hits_imask & ~(0xffffffffu << slot2));
        const char* pp = (const char*)(s.nodes + (size_t)(child_base + rel2) * 5u);
        prefetch_l2_line(pp);
        if (B200RT_SIB_PREFETCH >= 2) prefetch_l2_line(pp + 64);
        if (B200RT_SIB_PREFETCH >= 3) prefetch_l2_line(pp + 32);
#endif
    }
    const uint32_t slot = (bit - 24u) ^ octinv;
    const uint32_t rel = __popc(hits_imask & ~(0xffffffffu << slot));
    const uint4* np = s.nodes + (size_t)(child_base + rel) * 5u;
    const uint4 n0 = node_load<POLICY>(np, pol), n1 = node_load<POLICY>(np + 1, pol), n2 = node_load<POLICY>(np + 2, pol),
                n3 = node_load<POLICY>(np + 3, pol), n4 = node_load<POLICY>(np + 4, pol);
    if (st) st->nodes++;
    const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
    const uint32_t e_imask = n0.w;
    const float aix = __uint_as_float((e_imask & 0xffu) << 23) * s.idx;
    const float aiy = __uint_as_float(((e_imask >> 8) & 0xffu) << 23) * s.idy;
    const float aiz = __uint_as_float(((e_imask >> 16) & 0xffu) << 23) * s.idz;
    // plane t = (1024 + q) * ai + (ao - 1024 ai): the fp16 patterns carry the 1024 bias, the offset takes it back out
    const float aox = fm(-1024.0f, aix, (px - s.ox) * s.idx), aoy = fm(-1024.0f, aiy, (py - s.oy) * s.idy), aoz = fm(-1024.0f, aiz, (pz - s.oz) * s.idz);
    const float tfar = s.best.t, tmin = s.tmin;
    const uint32_t pack = s.pack;
    const bool negx = (pack >> 11) & 1u, negy = (pack >> 12) & 1u, negz = (pack >> 13) & 1u;
    const uint32_t octinv4 = octinv * 0x01010101u;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = byte_mask_from_bit4(is_inner4);
        const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
        const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
        const uint32_t xn = negx ? qhix : qlox, xf = negx ? qlox : qhix;
        const uint32_t yn = negy ? qhiy : qloy, yf = negy ? qloy : qhiy;
        const uint32_t zn = negz ? qhiz : qloz, zf = negz ? qloz : qhiz;
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
            const uint32_t pxn = jp ? q8_pair23(xn, pack) : q8_pair01(xn, pack), pxf = jp ? q8_pair23(xf, pack) : q8_pair01(xf, pack);
            const uint32_t pyn = jp ? q8_pair23(yn, pack) : q8_pair01(yn, pack), pyf = jp ? q8_pair23(yf, pack) : q8_pair01(yf, pack);
            const uint32_t pzn = jp ? q8_pair23(zn, pack) : q8_pair01(zn, pack), pzf = jp ? q8_pair23(zf, pack) : q8_pair01(zf, pack);
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {
                const int j = 2 * jp + jh;
                const float tnx = fm(jh ? q8_hi(pxn) : q8_lo(pxn), aix, aox), tfx = fm(jh ? q8_hi(pxf) : q8_lo(pxf), aix, aox);
                const float tny = fm(jh ? q8_hi(pyn) : q8_lo(pyn), aiy, aoy), tfy = fm(jh ? q8_hi(pyf) : q8_lo(pyf), aiy, aoy);
                const float tnz = fm(jh ? q8_hi(pzn) : q8_lo(pzn), aiz, aoz), tfz = fm(jh ? q8_hi(pzf) : q8_lo(pzf), aiz, aoz);
                const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
                const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tfar));
                if (cmin <= cmax * BOX_SLACK) {
                    const uint32_t cb = __byte_perm(child_bits4, 0u, 0x4440u | (uint32_t)j);   // byte j, one PRMT (the shift-and-mask form is two instructions)
                    const uint32_t bi = __byte_perm(bit_index4, 0u, 0x4440u | (uint32_t)j);
                    hitmask |= cb << bi;
                }
            }
        }
    }
    s.ngroup = make_uint2(n1.x, (hitmask & NODE_BITS) | (e_imask >> 24));
    return make_uint2(n1.y, hitmask & 0x00ffffffu);
}

// Node8F visit: the child planes are fp32 offsets from the node origin, stored plane-major; the near / far plane arrays are picked
// by ADDRESS from the ray's direction signs, so there is neither a decode nor a select per plane — 6 FFMA + 4 FMNMX per child.
template <int NS = 0>
__device__ __forceinline__ uint2 trav_node_step_f32(Trav& s, TStack<NS>& stack, TravStats* st)
{
    const uint32_t hits_imask = s.ngroup.y;
    const uint32_t bit = 31u - __clz(hits_imask);
    const uint32_t child_base = s.ngroup.x;
    s.ngroup.y &= ~(1u << bit);
    if (s.ngroup.y & NODE_BITS) {
        if (s.sp < TRAV_STACK) stack_push(s, stack, s.ngroup);
    }
    const uint32_t octinv = (s.pack >> 8) & 7u;
    const uint32_t slot = (bit - 24u) ^ octinv;
    const uint32_t rel = __popc(hits_imask & ~(0xffffffffu << slot));
    const uint4* np = s.nodes + (size_t)(child_base + rel) * (NODE8F_BYTES / 16u);
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1);
    if (st) st->nodes++;
    const float aox = (__uint_as_float(n0.x) - s.ox) * s.idx, aoy = (__uint_as_float(n0.y) - s.oy) * s.idy, aoz = (__uint_as_float(n0.z) - s.oz) * s.idz;
    const float tfar = s.best.t, tmin = s.tmin;
    // plane arrays (uint4 units): lo x/y/z at 2/4/6, hi x/y/z at 8/10/12
    const float4* xn = (const float4*)(np + (((s.pack >> 11) & 1u) ? 8u : 2u));
    const float4* xf = (const float4*)(np + (((s.pack >> 11) & 1u) ? 2u : 8u));
    const float4* yn = (const float4*)(np + (((s.pack >> 12) & 1u) ? 10u : 4u));
    const float4* yf = (const float4*)(np + (((s.pack >> 12) & 1u) ? 4u : 10u));
    const float4* zn = (const float4*)(np + (((s.pack >> 13) & 1u) ? 12u : 6u));
    const float4* zf = (const float4*)(np + (((s.pack >> 13) & 1u) ? 6u : 12u));
    const uint32_t octinv4 = octinv * 0x01010101u;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = byte_mask_from_bit4(is_inner4);
        const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const float4 vxn = __ldg(xn + half), vxf = __ldg(xf + half), vyn = __ldg(yn + half), vyf = __ldg(yf + half), vzn = __ldg(zn + half),
                     vzf = __ldg(zf + half);
        const float axn[4] = {vxn.x, vxn.y, vxn.z, vxn.w}, axf[4] = {vxf.x, vxf.y, vxf.z, vxf.w};
        const float ayn[4] = {vyn.x, vyn.y, vyn.z, vyn.w}, ayf[4] = {vyf.x, vyf.y, vyf.z, vyf.w};
        const float azn[4] = {vzn.x, vzn.y, vzn.z, vzn.w}, azf[4] = {vzf.x, vzf.y, vzf.z, vzf.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float tnx = fm(axn[j], s.idx, aox), tfx = fm(axf[j], s.idx, aox);
            const float tny = fm(ayn[j], s.idy, aoy), tfy = fm(ayf[j], s.idy, aoy);
            const float tnz = fm(azn[j], s.idz, aoz), tfz = fm(azf[j], s.idz, aoz);
            const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
            const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tfar));
            if (cmin <= cmax * BOX_SLACK) {
                const uint32_t cb = __byte_perm(child_bits4, 0u, 0x4440u | (uint32_t)j);   // byte j, one PRMT (the shift-and-mask form is two instructions)
                const uint32_t bi = __byte_perm(bit_index4, 0u, 0x4440u | (uint32_t)j);
                hitmask |= cb << bi;
            }
        }
    }
    s.ngroup = make_uint2(n1.x, (hitmask & NODE_BITS) | (n0.w & 0xffu));
    return make_uint2(n1.y, hitmask & 0x00ffffffu);
}

template <bool POLICY = false, int NS = 0>
__device__ __forceinline__ uint2 trav_node_step(Trav& s, TStack<NS>& stack, TravStats* st, uint64_t pol = 0)
{
    return (s.pack & TP_F32) ? trav_node_step_f32<NS>(s, stack, st) : trav_node_step_q8<POLICY, NS>(s, stack, st, pol);
}

// TERMINATE_ON_FIRST_HIT rays stop at the first accepted hit: no further instance is traversed
__device__ __forceinline__ bool any_ray_done(const Trav& s) { return (s.pack & TP_ANY) && (s.pack & TP_FOUND_ANY); }

// Arithmetic of tri_test (traverse.cuh) for one (ray, triangle) unit, without the acceptance bookkeeping: returns
// true when tmin < t <= tfar and the face-cull flags let the triangle through; the owner decides ties.
__device__ __forceinline__ bool tri_unit(const float* __restrict__ r, const float4 q0, const float4 q1, const float4 q2, float tfar, float& t_out,
                                         float& b1, float& b2)
{
    TriRay tr;
    tr.o = f3(r[0], r[1], r[2]);
    tr.Sx = r[3]; tr.Sy = r[4]; tr.Sz = r[5];
    const uint32_t k = __float_as_uint(r[6]);
    tr.kx = (int)(k & 3u); tr.ky = (int)((k >> 2) & 3u); tr.kz = (int)((k >> 4) & 3u);
    RayHit tmp;
    tmp.t = tfar;
    // found = true with the largest ordinal makes tri_test's tie rule accept every tmin < t <= tfar; the owner re-applies the real
    // rule (strictly smaller t, or equal t and a lower ordinal than its current best within this GAS)
    tmp.ord = 0xffffffffu;
    bool found = true;
    const bool hit = tri_test<false>(tr, q0, q1, q2, r[7], tmp, found, __float_as_uint(r[8]));
    t_out = tmp.t; b1 = tmp.b1; b2 = tmp.b2;
    return hit;
}

// tri_unit with the ray's axis permutation as compile-time constants: the nine component selections of the watertight test (sel3: two
// predicated moves each) fold away.  Same operations on the same values, so the same bits.  For kernels whose warps agree on the
// permutation — coherent rays, one per thread — the six-way branch around it is uniform; the cooperative rounds, whose lanes test
// triangles for rays of any direction, keep the dynamic form.
template <int KX, int KY, int KZ>
__device__ __forceinline__ bool tri_unit_k(const float* __restrict__ r, const float4 q0, const float4 q1, const float4 q2, float tfar, float& t_out,
                                           float& b1, float& b2)
{
    TriRay tr;
    tr.o = f3(r[0], r[1], r[2]);
    tr.Sx = r[3]; tr.Sy = r[4]; tr.Sz = r[5];
    tr.kx = KX; tr.ky = KY; tr.kz = KZ;
    RayHit tmp;
    tmp.t = tfar;
    tmp.ord = 0xffffffffu;
    bool found = true;
    const bool hit = tri_test<false>(tr, q0, q1, q2, r[7], tmp, found, __float_as_uint(r[8]));
    t_out = tmp.t; b1 = tmp.b1; b2 = tmp.b2;
    return hit;
}
#ifndef B200RT_TRI_STATIC_AXES
#define B200RT_TRI_STATIC_AXES 1
#endif
__device__ __forceinline__ bool tri_unit_by_axes(const float* __restrict__ r, const float4 q0, const float4 q1, const float4 q2, float tfar, float& t_out,
                                                 float& b1, float& b2)
{
#if B200RT_TRI_STATIC_AXES
    // kx | ky << 2 | kz << 4 (make_tri_ray: kz = dominant axis, kx = kz + 1, ky = kx + 1 mod 3, swapped for a negative dominant component)
    switch (__float_as_uint(r[6]) & 0x3fu) {
    case (1u | 2u << 2 | 0u << 4): return tri_unit_k<1, 2, 0>(r, q0, q1, q2, tfar, t_out, b1, b2);
    case (2u | 1u << 2 | 0u << 4): return tri_unit_k<2, 1, 0>(r, q0, q1, q2, tfar, t_out, b1, b2);
    case (2u | 0u << 2 | 1u << 4): return tri_unit_k<2, 0, 1>(r, q0, q1, q2, tfar, t_out, b1, b2);
    case (0u | 2u << 2 | 1u << 4): return tri_unit_k<0, 2, 1>(r, q0, q1, q2, tfar, t_out, b1, b2);
    case (0u | 1u << 2 | 2u << 4): return tri_unit_k<0, 1, 2>(r, q0, q1, q2, tfar, t_out, b1, b2);
    default: return tri_unit_k<1, 0, 2>(r, q0, q1, q2, tfar, t_out, b1, b2);
    }
#else
    return tri_unit(r, q0, q1, q2, tfar, t_out, b1, b2);
#endif
}

// Work concept:
//   __device__ bool  fetch(uint32_t item, Trav& s, float* my_ray)  load item, set s.best.t = tmax, call trav_begin_handle; false = nothing to trace
//   __device__ bool  next_instance(Trav& s, float* my_ray)         the traversal set up last has ended: set up the next instance of an IAS — or,
//                                                                  for items made of several rays (whitted.cu), the item's next ray — and
//                                                                  return true; false = the item is finished.  An any-hit ray that found
//                                                                  its hit ends here too (any_ray_done)
//   __device__ void  commit(const Trav& s, bool found)             store the result of the finished item
//   __device__ bool  stream_triangles()                            uniform over the launch: true = triangle records are not worth caching
//   static constexpr bool CONTINUES                                true: instead of commit, the finished lanes call
//   __device__ bool  commit_continue(Trav& s, float* my_ray, found) together; it may start another ray of the same item (true) — the place
//                                                                  for work that should run with many lanes at once (shading, whitted.cu)
//   static constexpr bool ANYHIT                                   any-hit programs exist: a candidate hit on a triangle whose geometry
//                                                                  flags lack DISABLE_ANYHIT is put to
//   __device__ bool  anyhit(prim, sbt, inst, pack, b1, b2, factor) (called by whichever lane tested the triangle: it may only use state
//                                                                  shared by all lanes plus its arguments; false = ignore the hit), and
//   __device__ void  attenuate(float factor)                       is called on the OWNING lane for every factor != 1
//   __device__ bool  anyhit_enabled()                              uniform over the launch: false skips all of it at run time
//   static constexpr bool NODE_POLICY (optional)                   the launch gives its node fetches an L2 eviction policy:
//   __device__ uint64_t node_policy()                              uniform over the launch (node_policy_for)
template <class Work, class = void> struct CoopHasNodePolicy { static constexpr bool value = false; };
template <class Work> struct CoopHasNodePolicy<Work, decltype((void)Work::NODE_POLICY)> { static constexpr bool value = Work::NODE_POLICY; };
template <class Work> constexpr bool coop_node_policy() { return CoopHasNodePolicy<Work>::value; }
//   static constexpr bool SMEM_STATE (optional)                    the lane's shared slot also holds the triangle base and the best hit's
//                                                                  fields; the Work's trav_begin* calls must pass the same value
template <class Work, class = void> struct CoopHasSmemState { static constexpr bool value = false; };
template <class Work> struct CoopHasSmemState<Work, decltype((void)Work::SMEM_STATE)> { static constexpr bool value = Work::SMEM_STATE; };
template <class Work> constexpr bool coop_smem_state() { return CoopHasSmemState<Work>::value; }

//   static constexpr bool COHERENT_AXES (optional)                 the rays of a warp mostly agree on their dominant axis (camera rays, probes
//                                                                  towards one light): the rounds test with tri_unit_by_axes
template <class Work, class = void> struct CoopHasCoherentAxes { static constexpr bool value = false; };
template <class Work> struct CoopHasCoherentAxes<Work, decltype((void)Work::COHERENT_AXES)> { static constexpr bool value = Work::COHERENT_AXES; };
template <class Work> constexpr bool coop_coherent_axes() { return CoopHasCoherentAxes<Work>::value; }
//   static constexpr int SMEM_STACK (optional)                     entries of the traversal stack kept in shared memory (TStack)
template <class Work, class = void> struct CoopHasSmemStack { static constexpr int value = 0; };
template <class Work> struct CoopHasSmemStack<Work, decltype((void)Work::SMEM_STACK)> { static constexpr int value = Work::SMEM_STACK; };
template <class Work> constexpr int coop_smem_stack() { return CoopHasSmemStack<Work>::value; }
template <int NS> struct CoopSharedStack { uint2 e[NS][COOP_BLOCK]; };
template <> struct CoopSharedStack<0> {};

template <class Work>
__device__ __forceinline__ void trace_persistent(Work& work, uint32_t n_items, unsigned int* __restrict__ fetch_counter, TravStats* st)
{
    constexpr bool SM = coop_smem_state<Work>();
    constexpr int NS = coop_smem_stack<Work>();
    __shared__ CoopSharedStack<NS> shs;
    constexpr int RAY_S_STRIDE = ray_s_stride<SM>();
    __shared__ CoopShared<RAY_S_STRIDE> sh;
    __shared__ CoopSharedAnyHit<Work::ANYHIT> sha;
    const bool tri_stream = work.stream_triangles();  // uniform over the launch
    constexpr bool NODE_POLICY = coop_node_policy<Work>();
    uint64_t node_pol = 0;  // uniform over the launch
    if constexpr (NODE_POLICY) node_pol = work.node_policy();
    bool ah_on = false;  // uniform: the launch has any-hit programs AND the traversable holds geometry that runs them
    if constexpr (Work::ANYHIT) ah_on = work.anyhit_enabled();
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    float* my_ray = &sh.ray[wid][lane * RAY_S_STRIDE];
    TStack<NS> stack;
    if constexpr (NS > 0) stack.shm = &shs.e[0][threadIdx.x];
    Trav s;
    s.tgroup = make_uint2(0u, 0u);
    s.ngroup = make_uint2(0u, 0u);
    s.pack = 0u;
    s.best.t = 0.f;
    bool has = false, fin = false, exhausted = false;
#if B200RT_FETCH_CHUNK
    // How much of the global cursor a warp takes at once.  Launches with at least 64 items per lane of the grid take B200RT_FETCH_CHUNK
    // items (bench, 30-60 M items per launch: 2092 -> 2143 Mrays/s); smaller launches take what the refill needs (chunk 0), because the
    // warp that holds the last chunk of expensive rays is their tail (measured with chunks throughout: Duck ray buffers 0.348 ->
    // 0.424 ms, Cornell 6.54 -> 7.03 ms).
    uint32_t loc_next = 0, loc_end = 0;
    const uint32_t chunk = n_items / (gridDim.x * (uint32_t)COOP_BLOCK) >= 64u ? (uint32_t)B200RT_FETCH_CHUNK : 0u;
    // (Measured and taken out again: a warp's first 32 items assigned by position, the cursor handing out what lies beyond the grid's
    // first round — one atomic round trip less per warp, but one more live value in a kernel capped at 64 registers: 8 -> 62 bytes
    // spilled, bench 2184 -> 2015 Mrays/s, nothing gained on the viewer's probes it was meant for.)
#endif
    for (;;) {
        // ---- commit finished rays together, then every lane without a ray takes the next work item
        if (fin) {
            load_hit<SM>(s, my_ray);
            if constexpr (Work::CONTINUES) { if (work.commit_continue(s, my_ray, (s.pack & TP_FOUND_ANY) != 0u)) has = true; }
            else work.commit(s, (s.pack & TP_FOUND_ANY) != 0u);
            fin = false;
        }
        unsigned need = __ballot_sync(FULL, !has && !exhausted);
        while (need) {
#if B200RT_FETCH_CHUNK
            // items come from a warp-private range [loc_next, loc_end) that is topped up from the global cursor: with chunks, one atomic
            // per `chunk` items instead of one per refill (a refill takes ~8 items; every warp of the launch hits the same word)
            if (loc_next == loc_end) {
                const int leader = __ffs(need) - 1;
                uint32_t base = 0;
                const uint32_t take = chunk ? chunk : (uint32_t)__popc(need);
                if ((int)lane == leader) base = atomicAdd(fetch_counter, take);
                base = __shfl_sync(FULL, base, leader);
                if (base >= n_items) {  // nothing left anywhere: every lane still waiting is done for good
                    if (!has) exhausted = true;
                    break;
                }
                loc_next = base;
                loc_end = base + take < n_items ? base + take : n_items;
            }
            const uint32_t avail = loc_end - loc_next;
            const uint32_t rank = (uint32_t)__popc(need & lt);
            if (!has && !exhausted && rank < avail) has = work.fetch(loc_next + rank, s, my_ray);
            const uint32_t want = (uint32_t)__popc(need);
            loc_next += want < avail ? want : avail;
#else
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(fetch_counter, (unsigned)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (!has && !exhausted) {
                const uint32_t item = base + __popc(need & lt);
                if (item >= n_items) exhausted = true;
                else has = work.fetch(item, s, my_ray);
            }
#endif
            need = __ballot_sync(FULL, !has && !exhausted);
        }
        __syncwarp();
        unsigned act = __ballot_sync(FULL, has);
        if (act == 0) break;
        const bool can_refill = !__any_sync(FULL, exhausted);
        // ---- traverse until the warp drains below the refill threshold
        for (;;) {
            // ---- phase A (divergent, short): lanes whose node group is used up take the next one off the stack, finish, or block.
            // Kept apart from the node step by __syncwarp(): the node step below must run ONCE per iteration with every lane that
            // has node work in it (without the barrier the compiler threads the "popped a node group" path straight into its own
            // copy of the node step and the warp executes that expensive block twice with complementary halves).
            bool blocked = false;  // this lane can only continue after a triangle round
            if (has && !(s.ngroup.y & NODE_BITS)) {
                // triangle groups met on the way down the stack are parked, or — slot taken — skipped over when a node group sits
                // right below them (B200RT_LOOKTHROUGH)
                while (s.sp > 0) {
                    const uint2 e = stack_peek(s, stack);
                    if (e.y & NODE_BITS) { s.ngroup = e; stack_drop(s, stack); break; }
                    if (s.tgroup.y == 0u) { s.tgroup = e; stack_drop(s, stack); continue; }
                    if (B200RT_LOOKTHROUGH && s.sp >= 2) {
                        const uint2 below = stack.get(s.sp - 2);
                        if (below.y & NODE_BITS) {
                            // take the node group from under the triangle group, which stays on top
                            s.ngroup = below;
#if !B200RT_STACK_TOP
                            stack.put(s.sp - 2, e);
#endif
                            --s.sp;
                            break;
                        }
                    }
                    blocked = true;  // both triangle slots taken and no node group within reach
                    break;
                }
                if (!(s.ngroup.y & NODE_BITS) && !blocked) {
                    if (s.tgroup.y != 0u) blocked = true;  // only the parked triangles are left
                    else if (work.next_instance(s, my_ray)) { /* next instance (or the item's next ray) set up */ }
                    else { has = false; fin = true; }
                }
            }
            __syncwarp();
            // ---- phase B: one node visit for every lane that has node work
            if (has && (s.ngroup.y & NODE_BITS)) {
                const uint2 nt = trav_node_step<NODE_POLICY, NS>(s, stack, st, node_pol);
                if (nt.y) {
#if B200RT_TRI_PREFETCH
                    // the parked group's records are read by other lanes, several node visits from now: ask them into L2 meanwhile
                    if (tri_stream) {
                        const uint32_t hi = 31u - __clz(nt.y), lo = (uint32_t)__ffs((int)nt.y) - 1u;
                        const char* t0 = (const char*)(lane_tris<SM>(s, my_ray) + (size_t)(nt.x + lo) * 3u);
                        const uint32_t span = (hi - lo) * 48u + 48u;   // bytes from the first record's start to the last one's end
                        prefetch_l2_line(t0);
                        if (B200RT_TRI_PREFETCH == 1) {
                            prefetch_l2_line(t0 + span - 16u);
                        } else {
                            if (span > 64u) prefetch_l2_line(t0 + 64);
                            if (span > 128u) prefetch_l2_line(t0 + 128);
                            prefetch_l2_line(t0 + span - 16u);
                        }
                    }
#endif
                    if (s.tgroup.y == 0u) s.tgroup = nt;
                    else if (s.sp < TRAV_STACK) stack_push(s, stack, nt);  // second parked group: goes on the stack (no NODE_BITS marks it)
                    else {
                        // stack full (pathological depth): test the group right here, one lane
                        uint2 g = nt;
                        while (g.y) {
                            const uint32_t ti = 31u - __clz(g.y);
                            g.y &= ~(1u << ti);
                            const float4* tp = lane_tris<SM>(s, my_ray) + (size_t)(g.x + ti) * 3u;
                            const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
                            if (st) st->tris++;
                            float t, b1, b2;
                            bool uh = tri_unit(my_ray, q0, q1, q2, s.best.t, t, b1, b2);
                            if constexpr (Work::ANYHIT) {
                                if (ah_on && uh && !anyhit_off(__float_as_uint(q1.w) >> TRI_FLAG_SHIFT, __float_as_uint(my_ray[8]))) {
                                    float fac;
                                    uh = work.anyhit(__float_as_uint(q0.w), __float_as_uint(q1.w) & TRI_SBT_MASK, s.inst, s.pack, b1, b2, fac);
                                    if (fac != 1.0f) work.attenuate(fac);
                                }
                            }
                            if (uh) {
                                const uint32_t ord = __float_as_uint(q2.w);
                                if (t < s.best.t || ((s.pack & TP_FOUND) && ord < s.best.ord)) {
                                    s.best.t = t; s.best.ord = ord;
                                    store_hit<SM>(s, my_ray, b1, b2, __float_as_uint(q0.w), __float_as_uint(q1.w));
                                    s.pack |= TP_FOUND | TP_FOUND_ANY;
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            // ---- triangle round?
            const uint32_t cnt = has ? (uint32_t)__popc(s.tgroup.y) : 0u;
            const uint32_t total = __reduce_add_sync(FULL, cnt);
            act = __ballot_sync(FULL, has);
            const unsigned blk = __ballot_sync(FULL, blocked);
            if (total >= (uint32_t)TRI_TRIGGER || (total != 0u && __popc(blk) * 2 >= __popc(act))) {
                // exclusive prefix sum of the unit counts
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, incl, d);
                    if ((int)lane >= d) incl += v;
                }
                const uint32_t first = incl - cnt;
                uint32_t pos = first, rem = has ? s.tgroup.y : 0u;
                for (uint32_t base = 0; base < total; base += 32u) {
                    while (rem != 0u && pos < base + 32u) {
                        const uint32_t ti = 31u - __clz(rem);
                        rem &= ~(1u << ti);
                        sh.unit_tri[wid][pos - base] = s.tgroup.x + ti;
                        sh.unit_owner[wid][pos - base] = lane;
                        ++pos;
                    }
                    __syncwarp();
                    const bool valid = base + lane < total;
                    const uint32_t owner = valid ? sh.unit_owner[wid][lane] : lane;
                    const float tfar = __shfl_sync(FULL, s.best.t, owner);
                    uint64_t tris_base;
                    if constexpr (SM) tris_base = (uint64_t)(uintptr_t)lane_tris<true>(s, &sh.ray[wid][owner * RAY_S_STRIDE]);
                    else tris_base = __shfl_sync(FULL, (unsigned long long)(uintptr_t)s.tris, owner);
                    float ut = 0.f, ub1 = 0.f, ub2 = 0.f;
                    uint32_t uord = 0xffffffffu, uprim = 0u, usbt = 0u;
                    bool uhit = false;
                    uint32_t opack = 0u, oinst = 0u;
                    if constexpr (Work::ANYHIT) {
                        if (ah_on) { opack = __shfl_sync(FULL, s.pack, owner); oinst = __shfl_sync(FULL, s.inst, owner); }
                    }
                    if (valid) {
                        const float4* tp = (const float4*)(uintptr_t)tris_base + (size_t)sh.unit_tri[wid][lane] * 3u;
                        const float4 q0 = TRI_LOAD(tp), q1 = TRI_LOAD(tp + 1), q2 = TRI_LOAD(tp + 2);
                        if (st) st->tris++;
                        if constexpr (coop_coherent_axes<Work>()) uhit = tri_unit_by_axes(&sh.ray[wid][owner * RAY_S_STRIDE], q0, q1, q2, tfar, ut, ub1, ub2);
                        else uhit = tri_unit(&sh.ray[wid][owner * RAY_S_STRIDE], q0, q1, q2, tfar, ut, ub1, ub2);
                        uord = __float_as_uint(q2.w); uprim = __float_as_uint(q0.w); usbt = __float_as_uint(q1.w);
                    }
                    if constexpr (Work::ANYHIT) {
                        // the any-hit program of the candidate runs on the lane that tested it; its verdict (and the occlusion factor)
                        // goes back to the owner with the unit's result
                        if (ah_on) {
                            float fac = 1.0f;
                            if (uhit && !anyhit_off(usbt >> TRI_FLAG_SHIFT, __float_as_uint(sh.ray[wid][owner * RAY_S_STRIDE + 8])))
                                uhit = work.anyhit(uprim, usbt & TRI_SBT_MASK, oinst, opack, ub1, ub2, fac);
                            sha.res_fac[wid][lane] = fac;
                        }
                    }
                    sh.res_t[wid][lane] = uhit ? ut : __int_as_float(0x7f800000);
                    sh.res_ord[wid][lane] = uord;
                    __syncwarp();
                    // owners: lexicographic minimum (t, ordinal) over their units of this chunk, same acceptance rule as tri_test
                    int win = -1;
                    if (cnt) {
                        const uint32_t lo = first > base ? first : base;
                        const uint32_t hi = (first + cnt) < (base + 32u) ? (first + cnt) : (base + 32u);
                        for (uint32_t p = lo; p < hi; ++p) {
                            const float t = sh.res_t[wid][p - base];
                            const uint32_t ord = sh.res_ord[wid][p - base];
                            if constexpr (Work::ANYHIT) {
                                if (ah_on) {
                                    const float fac = sha.res_fac[wid][p - base];
                                    if (fac != 1.0f) work.attenuate(fac);
                                }
                            }
                            if (t < s.best.t || (t == s.best.t && (s.pack & TP_FOUND) && ord < s.best.ord)) {
                                s.best.t = t; s.best.ord = ord; win = (int)(p - base);
                                s.pack |= TP_FOUND | TP_FOUND_ANY;
                            }
                        }
                    }
                    const int src = win >= 0 ? win : (int)lane;
                    const float wb1 = __shfl_sync(FULL, ub1, src), wb2 = __shfl_sync(FULL, ub2, src);
                    const uint32_t wprim = __shfl_sync(FULL, uprim, src), wsbt = __shfl_sync(FULL, usbt, src);
                    if (win >= 0) store_hit<SM>(s, my_ray, wb1, wb2, wprim, wsbt);
                    __syncwarp();
                }
                if (has) {
                    s.tgroup.y = 0u;
                    if ((s.pack & TP_ANY) && (s.pack & TP_FOUND)) { s.ngroup.y = 0u; s.sp = 0; }  // any-hit: done
                }
            }
            if (act == 0u || (can_refill && __popc(act) < REFILL_THRESHOLD)) break;
        }
    }
}

// A warp's tile of launch indices for kernels that run one launch index per thread over a width x height launch: TILE_W x TILE_H
// neighbours instead of 32 consecutive indices of a row (rays of a tile walk the same nodes more often); a CTA of 128 threads is 2 x 2 tiles.
#ifndef B200RT_TILE_W
#define B200RT_TILE_W 8
#endif
#ifndef B200RT_TILE_CTA_WARPS
#define B200RT_TILE_CTA_WARPS 4   // 4: a CTA is 2 x 2 tiles, 2: 2 x 1, 1: one tile
#endif
constexpr uint32_t TILE_W = B200RT_TILE_W, TILE_H = 32u / TILE_W, TILE_CTA_THREADS = 32u * B200RT_TILE_CTA_WARPS;
constexpr uint32_t CTA_TILE_W = (B200RT_TILE_CTA_WARPS >= 2 ? 2u : 1u) * TILE_W, CTA_TILE_H = (B200RT_TILE_CTA_WARPS >= 4 ? 2u : 1u) * TILE_H;
__device__ __forceinline__ void tile_xy(uint32_t bx, uint32_t by, uint32_t& x, uint32_t& y)   // CTA tile (bx, by)
{
    const uint32_t wrp = threadIdx.x >> 5, ln = threadIdx.x & 31u;
    x = bx * CTA_TILE_W + (wrp & 1u) * TILE_W + ln % TILE_W;
    y = by * CTA_TILE_H + (wrp >> 1) * TILE_H + ln / TILE_W;
}
__device__ __forceinline__ void tile_xy(uint32_t& x, uint32_t& y) { tile_xy(blockIdx.x, blockIdx.y, x, y); }

// The same order for launches whose lanes are a flat index: index -> pixel, bands of TILE_H rows cut into TILE_W x TILE_H tiles, a
// bijection of [0, width * height) for any size (columns beyond the last whole tile of a band and the rows of a last partial band
// follow in row order).  Work lists compacted in lane order then hold 2-D neighbours next to each other.
__device__ __forceinline__ void tile_order_xy(uint32_t i, uint32_t width, uint32_t height, uint32_t& x, uint32_t& y)
{
    const uint32_t band_sz = TILE_H * width, band = i / band_sz, r = i - band * band_sz;
    const uint32_t wt = width - width % TILE_W;
    if ((band + 1u) * TILE_H > height) { x = i % width; y = i / width; return; }
    if (r < TILE_H * wt) { const uint32_t t = r / 32u, k = r % 32u; x = t * TILE_W + k % TILE_W; y = band * TILE_H + k / TILE_W; }
    else { const uint32_t q = r - TILE_H * wt, wr = width - wt; x = wt + q % wr; y = band * TILE_H + q / wr; }
}

// ---- one ray per thread ---------------------------------------------------------------------------------------------------------------
// The same traversal state machine without the warp-cooperative machinery: every thread takes ONE work item, visits its nodes and tests
// the triangles of a leaf group as it meets them.  For COHERENT ray buffers (orthographic / camera rays in pixel order) the lanes of a
// warp walk the same nodes anyway, so there is nothing for the cooperative rounds to repair, and their bookkeeping (shared-memory unit
// lists, prefix sums, refills) is pure cost: optixRaycasting's two 1 M-ray ortho batches on the Duck took 0.35 ms on the persistent
// driver, against OptiX's 0.33 (profiles/r02_small_scenes.md).  Incoherent rays stay on trace_persistent.  Same Work concept (an any-hit program
// runs on the lane that tested the triangle — its own), same tri_unit arithmetic, same hit rule: bit-identical results.
// (Measured and not kept: postponing the leaves — a lane parks its triangle group and the warp tests once 1/4, 1/2 or 3/4 of the lanes
// still traversing hold one.  Triangle tests do run at 10 of 32 lanes here, but the warp-wide ballots, the idle parked lanes and 16 more
// registers cost more: ray buffers 0.201 -> 0.228-0.232 ms, viewer 0.191 -> 0.195-0.202 ms; profiles/r02_session4_experiments.md.)
template <class Work>
__device__ __forceinline__ void trace_one_per_thread(Work& work, uint32_t item, bool valid, TravStats* st)
{
    static_assert(!Work::CONTINUES, "continuing launches use trace_persistent");
    constexpr bool SM = coop_smem_state<Work>();
    bool ah_on = false;  // uniform: the launch has any-hit programs AND the traversable holds geometry that runs them
    if constexpr (Work::ANYHIT) ah_on = work.anyhit_enabled();
    float my_ray[ray_s_stride<SM>()];   // registers here: every index is a compile-time constant
    TStack<0> stack;
    Trav s;
    s.tgroup = make_uint2(0u, 0u);
    s.ngroup = make_uint2(0u, 0u);
    s.pack = 0u;
    s.best.t = 0.f;
    if (!valid || !work.fetch(item, s, my_ray)) return;   // fetch commits a ray that has nothing to traverse
    for (;;) {
        if (s.ngroup.y & NODE_BITS) {
            const uint2 nt = trav_node_step(s, stack, st);
            uint32_t rem = nt.y;
            while (rem) {
                const uint32_t ti = 31u - __clz(rem);
                rem &= ~(1u << ti);
                const float4* tp = lane_tris<SM>(s, my_ray) + (size_t)(nt.x + ti) * 3u;
                const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
                if (st) st->tris++;
                float t, b1, b2;
                bool uh = tri_unit_by_axes(my_ray, q0, q1, q2, s.best.t, t, b1, b2);
                if constexpr (Work::ANYHIT) {
                    // the candidate's any-hit program runs right here: this lane has everything it needs
                    if (ah_on && uh && !anyhit_off(__float_as_uint(q1.w) >> TRI_FLAG_SHIFT, __float_as_uint(my_ray[8]))) {
                        float fac;
                        uh = work.anyhit(__float_as_uint(q0.w), __float_as_uint(q1.w) & TRI_SBT_MASK, s.inst, s.pack, b1, b2, fac);
                        if (fac != 1.0f) work.attenuate(fac);
                    }
                }
                if (uh) {
                    const uint32_t ord = __float_as_uint(q2.w);
                    if (t < s.best.t || ((s.pack & TP_FOUND) && ord < s.best.ord)) {
                        s.best.t = t; s.best.ord = ord;
                        store_hit<SM>(s, my_ray, b1, b2, __float_as_uint(q0.w), __float_as_uint(q1.w));
                        s.pack |= TP_FOUND | TP_FOUND_ANY;
                        if (s.pack & TP_ANY) { s.ngroup.y = 0u; s.sp = 0; break; }
                    }
                }
            }
            continue;
        }
        if (s.sp > 0) { s.ngroup = stack_peek(s, stack); stack_drop(s, stack); continue; }   // only node groups are ever pushed here
        if (work.next_instance(s, my_ray)) continue;
        break;
    }
    load_hit<SM>(s, my_ray);
    work.commit(s, (s.pack & TP_FOUND_ANY) != 0u);
}

}  // namespace b200rt
