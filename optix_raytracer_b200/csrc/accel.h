// accel.h — in-memory layout of the acceleration-structure blob that lives in the caller's
// outputBuffer (the opaque bytes optixAccelBuild writes in the reference,
// SDK/optixPathTracer/optixPathTracer.cpp:650-664).  All offsets are relative to the blob base so
// b200rt_accel_compact can move it (optixPathTracer.cpp:671-683).
//
//   GAS:  [AccelHeader 128 B][Node8 x num_nodes (80 B each)][TriRecord x num_tris (48 B each)]
//   IAS:  [AccelHeader 128 B][InstanceRecord x num_instances (128 B each)]
//
// Two node encodings share the topology fields (child base, triangle base, 8 meta bytes, inner mask, octant-ordered slots):
//   Node8  (80 B)  child boxes quantised to 8 bits — for scenes whose BVH does not fit the caches: every node visit moves 80 B,
//                  at the price of decoding 48 bytes per visit;
//   Node8F (224 B) child boxes as fp32 offsets from the node origin, plane-major ([lo x][lo y][lo z][hi x][hi y][hi z], 8 floats
//                  each) — for cache-resident scenes, where the traversal is issue-bound and the decode is the larger half of a
//                  node visit (profiles/r01_trace_v3.md).  The builder picks by triangle count (bvh_build.cu: choose_node_bytes).
// Node8 is an 8-wide BVH node with child boxes quantised to 8 bits on a per-node power-of-two grid
// (compressed wide BVH, Ylitie/Karras/Laine 2017), fetched as 5 x 16-byte loads.  TriRecord is
// three float4: the xyz are the object-space vertices exactly as supplied (the watertight test runs
// on them, so results do not depend on the BVH), the w lanes carry the primitive index, SBT
// offset | geometry flags, and the GAS-global ordinal used for tie-breaking.
#pragma once
#include <stdint.h>

namespace b200rt {

constexpr uint32_t ACCEL_MAGIC = 0x54523242u;  // "B2RT"
constexpr uint32_t ACCEL_KIND_GAS = 1, ACCEL_KIND_IAS = 2;
constexpr uint32_t NODE8_BYTES = 80, NODE8F_BYTES = 224, TRI_BYTES = 48, INSTREC_BYTES = 128, HEADER_BYTES = 128;

struct AccelHeader {
    uint32_t magic;
    uint32_t kind;
    uint32_t num_tris;
    uint32_t num_nodes;      // written by the builder on the device
    uint64_t nodes_off;
    uint64_t tris_off;
    uint64_t total_bytes;    // exact size after compaction
    float bounds[6];         // lo xyz, hi xyz
    uint32_t num_instances;
    uint32_t depth;
    uint64_t inst_off;
    uint32_t max_nodes;      // node capacity of this (uncompacted) blob
    uint32_t error;          // builder overflow flag
    uint32_t node_bytes;     // 80: Node8 (8-bit quantised child boxes); 224: Node8F (fp32 child boxes), see below
    uint32_t anyhit;         // 1: some triangle (GAS) / some instanced GAS (IAS) leaves any-hit enabled (geometry flags without DISABLE_ANYHIT)
    uint32_t pad[8];
};
static_assert(sizeof(AccelHeader) == HEADER_BYTES, "header must be 128 bytes");

struct InstanceRecord {
    float m[12];             // object -> world
    float inv[12];           // world -> object
    uint64_t gas;            // device address of the GAS blob
    uint32_t instance_id;
    uint32_t sbt_offset;
    uint32_t mask;
    uint32_t flags;
    float wbounds[2];        // unused padding to 128 B
};
static_assert(sizeof(InstanceRecord) == INSTREC_BYTES, "instance record must be 128 bytes");

// Entries of the per-lane traversal stack (traverse.cuh, trav_coop.cuh).  A visited level leaves at most two entries behind (the rest of
// its node group and a parked triangle group), so a tree of depth d needs 2 d entries.  The builder keeps every tree within
// TRAV_STACK / 2 levels whatever the input (bvh_build.cu: collapse_plan_kernel opens the tallest subtrees first where the surface-area
// order would run out of levels), so no input is refused and no traversal ever drops a child.  The 50 M-triangle bench scene has depth 17.
constexpr int TRAV_STACK = 64;

// TriRecord w-lane packing
constexpr uint32_t TRI_SBT_MASK = 0x00ffffffu;   // v1.w low 24 bits: GAS-local SBT index
constexpr uint32_t TRI_FLAG_SHIFT = 24;          // v1.w high 8 bits: OptixGeometryFlags of its SBT record

// extended hit record of the C ABI (b200rt.h): {t, prim, inst, b1, b2}
struct ExtHit { float t; uint32_t prim; uint32_t inst; float b1; float b2; };

// The cull word a triangle test sees: the ray's CULL_* flags (bits 4-7, reference include/optix_types.h:1819-1839) after the instance
// flags had their say, plus the any-hit override in bits 0-1 — 1 = any-hit off for every triangle, 2 = any-hit on for every triangle,
// 0 = the triangle's own OPTIX_GEOMETRY_FLAG_DISABLE_ANYHIT decides.  Precedence as documented in include/optix_types.h:1088-1108 and
// 1794-1806: ray flags over instance flags over geometry flags.  OPTIX_INSTANCE_FLAG_DISABLE_TRIANGLE_FACE_CULLING drops the two
// face-cull bits, OPTIX_INSTANCE_FLAG_FLIP_TRIANGLE_FACING swaps them.  Same function in oracle.cpp (cull_word).
// The ray's 8-bit OptixVisibilityMask (include/optix_types.h OptixVisibilityMask; optixTrace's visibilityMask argument) travels in bits
// 16-23 of the ray-flags word, stored XOR 1 so that a flags word without the field means mask 1 — what optixPathTracer, optixRaycasting
// and whitted trace with (B200RT_RAY_VISIBILITY_MASK in b200rt.h).  An instance is traversed when (instance mask & ray mask) != 0.
__host__ __device__ __forceinline__ uint32_t ray_visibility(uint32_t ray_flags) { return ((ray_flags >> 16) ^ 1u) & 0xffu; }

__host__ __device__ __forceinline__ uint32_t cull_word(uint32_t ray_flags, uint32_t inst_flags)
{
    uint32_t c = ray_flags & 0xf0u;
    if (inst_flags & 1u) c &= ~0x30u;
    else if (inst_flags & 2u) c = (c & ~0x30u) | ((c & 0x10u) << 1) | ((c & 0x20u) >> 1);
    uint32_t force = (ray_flags & 1u) ? 1u : (ray_flags & 2u) ? 2u : 0u;
    if (!force) force = (inst_flags & 4u) ? 1u : (inst_flags & 8u) ? 2u : 0u;
    return c | force;
}

}  // namespace b200rt
