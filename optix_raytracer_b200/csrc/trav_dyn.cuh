// trav_dyn.cuh — persistent-warp traversal driver with dynamic ray fetch and triangle postponing.
//
// The per-ray traversal of traverse.cuh is re-expressed as a state machine (`Trav`) so that one warp
// can keep its 32 lanes busy: a lane whose ray terminates fetches the next work item from a global
// counter instead of idling until the longest ray of the warp is done (Aila/Laine 2009 "persistent
// threads", Ylitie/Karras/Laine 2017 dynamic fetch).  Measured motivation (profiles/r01_*): with one
// ray per thread and no refill the trace kernel ran at 5.6 of 32 threads per instruction.
//
// Semantics are exactly those of traverse.cuh (same tri_test, same conservative box test, closest =
// min (t, instance, ordinal)); only the scheduling differs, and since the hit rule is independent of
// visit order, results are bit-identical.
//
//   * node step   : pop the best pending child of the current node group, fetch its 80-byte node, test the
//                   8 quantised child boxes -> new node group + triangle group
//   * triangle phase: runs only when at least 1/4 of the warp's active lanes have triangles pending,
//                   otherwise the triangle group is pushed on the lane's stack (postponed)
//   * refill      : when fewer than REFILL_THRESHOLD lanes are active and work is left, terminated lanes
//                   fetch new items (one warp-aggregated atomicAdd)
#pragma once
#include "traverse.cuh"

namespace b200rt {

constexpr uint32_t NODE_BITS = 0xff000000u;
constexpr int REFILL_THRESHOLD = 24;

// pack word layout
constexpr uint32_t TP_ANY = 1u << 16;        // any-hit (terminate on first hit)
constexpr uint32_t TP_FOUND = 1u << 17;      // a hit was accepted in the GAS currently being traversed (tie-break scope)
constexpr uint32_t TP_FOUND_ANY = 1u << 18;  // a hit was accepted in any instance so far

struct Trav {
    const uint4* nodes;
    const float4* tris;
    float ox, oy, oz;       // origin in the space of the GAS being traversed
    float Sx, Sy, Sz;       // watertight shear constants
    float idx, idy, idz;    // reciprocal (clamped) direction for the box tests
    float tmin;
    uint32_t pack;          // kx | ky<<2 | kz<<4 | octinv<<8 | negx<<11 | negy<<12 | negz<<13 | TP_*
    uint32_t cull;
    uint32_t inst;          // index of the instance being traversed (0 for a bare GAS)
    uint2 ngroup, tgroup;
    int sp;
    RayHit best;
};

// Prepare the per-ray constants for a GAS (object-space origin/direction).  best.t must hold tmax.
__device__ __forceinline__ void trav_begin(Trav& s, const AccelHeader* __restrict__ gas, float3 o, float3 d, float tmin, uint32_t keep_flags)
{
    const char* base = (const char*)gas;
    s.nodes = (const uint4*)(base + gas->nodes_off);
    s.tris = (const float4*)(base + gas->tris_off);
    const TriRay tr = make_tri_ray(o, d);
    s.ox = o.x; s.oy = o.y; s.oz = o.z;
    s.Sx = tr.Sx; s.Sy = tr.Sy; s.Sz = tr.Sz;
    const float bx = fabsf(d.x) < DIR_EPS ? copysignf(DIR_EPS, d.x) : d.x;
    const float by = fabsf(d.y) < DIR_EPS ? copysignf(DIR_EPS, d.y) : d.y;
    const float bz = fabsf(d.z) < DIR_EPS ? copysignf(DIR_EPS, d.z) : d.z;
    s.idx = fdiv(1.0f, bx); s.idy = fdiv(1.0f, by); s.idz = fdiv(1.0f, bz);
    const uint32_t nx = bx < 0.0f, ny = by < 0.0f, nz = bz < 0.0f;
    const uint32_t oct = (nx << 2) | (ny << 1) | nz;
    s.pack = (uint32_t)tr.kx | ((uint32_t)tr.ky << 2) | ((uint32_t)tr.kz << 4) | ((7u - oct) << 8) | (nx << 11) | (ny << 12) | (nz << 13) | keep_flags;
    s.tmin = tmin;
    s.ngroup = gas->num_tris ? make_uint2(0u, 0x80000000u) : make_uint2(0u, 0u);
    s.tgroup = make_uint2(0u, 0u);
    s.sp = 0;
}

__device__ __forceinline__ void trav_node_step(Trav& s, uint2* __restrict__ stack, TravStats* st)
{
    const uint32_t hits_imask = s.ngroup.y;
    const uint32_t bit = 31u - __clz(hits_imask);
    const uint32_t child_base = s.ngroup.x;
    s.ngroup.y &= ~(1u << bit);
    if (s.ngroup.y & NODE_BITS) {
        if (s.sp < TRAV_STACK) stack[s.sp++] = s.ngroup;
    }
    const uint32_t octinv = (s.pack >> 8) & 7u;
    const uint32_t slot = (bit - 24u) ^ octinv;
    const uint32_t rel = __popc(hits_imask & ~(0xffffffffu << slot));
    const uint4* np = s.nodes + (size_t)(child_base + rel) * 5u;
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    if (st) st->nodes++;
    const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
    const uint32_t e_imask = n0.w;
    const float aix = __uint_as_float((e_imask & 0xffu) << 23) * s.idx;
    const float aiy = __uint_as_float(((e_imask >> 8) & 0xffu) << 23) * s.idy;
    const float aiz = __uint_as_float(((e_imask >> 16) & 0xffu) << 23) * s.idz;
    const float aox = (px - s.ox) * s.idx, aoy = (py - s.oy) * s.idy, aoz = (pz - s.oz) * s.idz;
    const float tfar = s.best.t, tmin = s.tmin;
    const bool negx = (s.pack >> 11) & 1u, negy = (s.pack >> 12) & 1u, negz = (s.pack >> 13) & 1u;
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
        for (int j = 0; j < 4; ++j) {
            const float tnx = fm(byte_f(xn, j), aix, aox), tfx = fm(byte_f(xf, j), aix, aox);
            const float tny = fm(byte_f(yn, j), aiy, aoy), tfy = fm(byte_f(yf, j), aiy, aoy);
            const float tnz = fm(byte_f(zn, j), aiz, aoz), tfz = fm(byte_f(zf, j), aiz, aoz);
            const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
            const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tfar));
            if (cmin <= cmax * BOX_SLACK) {
                const uint32_t cb = (child_bits4 >> (8 * j)) & 0xffu;
                const uint32_t bi = (bit_index4 >> (8 * j)) & 0xffu;
                hitmask |= cb << bi;
            }
        }
    }
    s.ngroup = make_uint2(n1.x, (hitmask & NODE_BITS) | (e_imask >> 24));
    s.tgroup = make_uint2(n1.y, hitmask & 0x00ffffffu);
}

// one triangle of the pending group; returns true when the ray is finished (any-hit found)
__device__ __forceinline__ bool trav_tri_step(Trav& s, TravStats* st)
{
    const uint32_t ti = 31u - __clz(s.tgroup.y);
    s.tgroup.y &= ~(1u << ti);
    const float4* tp = s.tris + (size_t)(s.tgroup.x + ti) * 3u;
    const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
    if (st) st->tris++;
    TriRay tr;
    tr.o = f3(s.ox, s.oy, s.oz);
    tr.kx = (int)(s.pack & 3u); tr.ky = (int)((s.pack >> 2) & 3u); tr.kz = (int)((s.pack >> 4) & 3u);
    tr.Sx = s.Sx; tr.Sy = s.Sy; tr.Sz = s.Sz;
    bool found = (s.pack & TP_FOUND) != 0;
    const bool hit = tri_test<false>(tr, q0, q1, q2, s.tmin, s.best, found, s.cull);
    if (hit) { s.pack |= TP_FOUND | TP_FOUND_ANY; s.best.inst = s.inst; }
    return hit && (s.pack & TP_ANY);
}

// Set up traversal of `h` (GAS, or the first usable instance >= first_inst of an IAS) for the world-space ray.
// keep = TP_* bits to carry over.  Returns false when there is nothing (more) to traverse.
__device__ __forceinline__ bool trav_begin_handle(Trav& s, const AccelHeader* __restrict__ h, float3 o, float3 d, float tmin, uint32_t keep,
                                                  uint32_t cull, uint32_t first_inst)
{
    if (h->kind == ACCEL_KIND_GAS) {
        if (first_inst > 0u) return false;
        s.inst = 0u;
        s.cull = cull;
        trav_begin(s, h, o, d, tmin, keep);
        return true;
    }
    const InstanceRecord* recs = (const InstanceRecord*)((const char*)h + h->inst_off);
    const uint32_t n = h->num_instances;
    for (uint32_t k = first_inst; k < n; ++k) {
        const InstanceRecord* ir = recs + k;
        if (!(ir->mask & 1u)) continue;
        s.inst = k;
        s.cull = (ir->flags & 1u) ? 0u : cull;  // OPTIX_INSTANCE_FLAG_DISABLE_TRIANGLE_FACE_CULLING
        trav_begin(s, (const AccelHeader*)ir->gas, xform_point(ir->inv, o), xform_vec(ir->inv, d), tmin, keep);
        return true;
    }
    return false;
}

// Work concept:
//   __device__ bool  fetch(uint32_t item, Trav& s)      load item, call trav_begin, set s.best.t = tmax; false = nothing to trace
//   __device__ bool  next_instance(Trav& s)             IAS: set up the next instance (true) or report none left (false)
//   __device__ void  commit(const Trav& s, bool found)  store the result of the finished item
template <class Work>
__device__ __forceinline__ void trace_persistent(Work& work, uint32_t n_items, unsigned int* __restrict__ fetch_counter, TravStats* st)
{
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    uint2 stack[TRAV_STACK];
    Trav s;
    bool has = false, exhausted = false;
    for (;;) {
        // ---- refill: every lane without a ray takes the next work item
        unsigned need = __ballot_sync(FULL, !has && !exhausted);
        while (need) {
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(fetch_counter, (unsigned)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (!has && !exhausted) {
                const uint32_t item = base + __popc(need & lt);
                if (item >= n_items) exhausted = true;
                else has = work.fetch(item, s);
            }
            need = __ballot_sync(FULL, !has && !exhausted);
        }
        unsigned act = __ballot_sync(FULL, has);
        if (act == 0) break;
        const bool can_refill = !__any_sync(FULL, exhausted);
        // ---- traverse until the warp drains below the refill threshold
        for (;;) {
            if (has) {
                if (s.ngroup.y & NODE_BITS) trav_node_step(s, stack, st);
                else { s.tgroup = s.ngroup; s.ngroup = make_uint2(0u, 0u); }  // a postponed triangle group was popped
            }
            // triangle phase: only with enough company, otherwise postpone
            unsigned tm = __ballot_sync(FULL, has && s.tgroup.y != 0u);
            if (tm) {
                if (__popc(tm) * 4 >= __popc(act)) {
                    while (tm) {
                        if (has && s.tgroup.y != 0u) {
                            if (trav_tri_step(s, st)) { s.tgroup.y = 0u; s.ngroup.y = 0u; s.sp = 0; }  // any-hit: done
                        }
                        tm = __ballot_sync(FULL, has && s.tgroup.y != 0u);
                    }
                } else if (has && s.tgroup.y != 0u) {
                    if (s.sp < TRAV_STACK) { stack[s.sp++] = s.tgroup; s.tgroup.y = 0u; }
                    else { while (s.tgroup.y) { if (trav_tri_step(s, st)) { s.tgroup.y = 0u; s.ngroup.y = 0u; s.sp = 0; } } }
                }
            }
            // pop / finish
            if (has && (s.ngroup.y & NODE_BITS) == 0u && s.tgroup.y == 0u) {
                if (s.sp > 0) s.ngroup = stack[--s.sp];
                else if (!((s.pack & TP_ANY) && (s.pack & TP_FOUND_ANY)) && work.next_instance(s)) { /* continue with the next instance */ }
                else { work.commit(s, (s.pack & TP_FOUND_ANY) != 0u); has = false; }
            }
            act = __ballot_sync(FULL, has);
            if (act == 0u || (can_refill && __popc(act) < REFILL_THRESHOLD)) break;
        }
    }
}

}  // namespace b200rt
