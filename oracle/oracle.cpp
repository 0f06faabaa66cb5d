// ============================================================================================
// oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Scalar C++ restatement of the reference hot path (the OptiX launch behind optixPathTracer,
// optixMultiGPU and optixRaycasting).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library, and only as the checker or the CPU
// baseline — never as the thing shipped.  The product (optix_raytracer_b200/csrc) has no CPU path.
//
// What is pinned and what is not
//   * Integer / host parts — tea<4>, lcg, rnd (SDK/cuda/random.h:30-67), StaticWorkDistribution
//     (SDK/sutil/WorkDistribution.h:50-81), Camera::UVWFrame (SDK/sutil/Camera.cpp:34-46) — are
//     pinned bit-exactly against the reference's own headers compiled in place
//     (oracle/ref_shim.cpp -> tests/golden/kat.json).
//   * Ray/triangle intersection and BVH traversal live in the closed libnvoptix.so.1 (OptiX ABI
//     87, not in /root/reference).  There is nothing to follow and no golden vector in the
//     reference: PARITY UNPINNED for that part.  The oracle *defines* the contract instead
//     (DESIGN.md "arithmetic contract"): Woop/Benthin/Wald watertight test in a fixed sequence of
//     IEEE-754 binary32 operations, closest hit = min (t, instance, primitive ordinal), which makes
//     the answer independent of the acceleration structure.  Small scenes are traced brute force.
//   * Shading follows the reference device programs line by line (citations at each function),
//     with every fused multiply-add named explicitly so the CUDA kernels (compiled -fmad=false)
//     and this file (compiled -ffp-contract=off) produce identical bits.
// ============================================================================================
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <functional>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// vector helpers — semantics of SDK/sutil/vec_math.h:500-570 with the fma placement nvcc uses
// ------------------------------------------------------------------------------------------
struct f3 { float x, y, z; };
static inline f3 mk(float x, float y, float z) { return {x, y, z}; }
static inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline f3 neg(f3 a) { return {-a.x, -a.y, -a.z}; }
static inline float fm(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
// dot = fma(z,z, fma(y,y, x*x))
static inline float dot(f3 a, f3 b) { return fm(a.z, b.z, fm(a.y, b.y, a.x * b.x)); }
// cross component = fma(a1,b2, -(a2*b1))
static inline f3 cross(f3 a, f3 b)
{
    return {fm(a.y, b.z, -(a.z * b.y)), fm(a.z, b.x, -(a.x * b.z)), fm(a.x, b.y, -(a.y * b.x))};
}
static inline float length(f3 v) { return sqrtf(dot(v, v)); }
// vec_math.h normalize: v * (1.0f / sqrtf(dot(v,v)))
static inline f3 normalize(f3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }
static inline float clampf(float x, float a, float b) { return fmaxf(a, fminf(x, b)); }
static inline float get(const f3& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// ------------------------------------------------------------------------------------------
// RNG — SDK/cuda/random.h:30-67 (integer, bit-exact)
// ------------------------------------------------------------------------------------------
static inline uint32_t tea4(uint32_t val0, uint32_t val1)
{
    uint32_t a = val0, b = val1, sum = 0;
    for (int round = 0; round < 4; ++round) {
        sum += 0x9e3779b9u;
        a += ((b << 4) + 0xa341316cu) ^ (b + sum) ^ ((b >> 5) + 0xc8013ea4u);
        b += ((a << 4) + 0xad90777du) ^ (a + sum) ^ ((a >> 5) + 0x7e95761eu);
    }
    return a;
}
static inline uint32_t lcg(uint32_t& state)
{
    state = 1664525u * state + 1013904223u;
    return state & 0x00FFFFFFu;
}
static inline float rnd(uint32_t& state) { return (float)lcg(state) / 16777216.0f; }

// ------------------------------------------------------------------------------------------
// deterministic sin/cos on [0, 2*pi] (arithmetic contract: Cody–Waite reduction by pi/2 in three
// fma steps + degree-7/8 minimax polynomials, all fma; replaces the reference's fast-math
// __sinf/__cosf of optixPathTracer.cu:149-159, which has no bit-reproducible CPU counterpart).
// ------------------------------------------------------------------------------------------
static inline void det_sincos(float phi, float& s, float& c)
{
    const float TWO_OVER_PI = 0.636619772f;
    const float P1 = 1.5703125f, P2 = 4.837512969970703125e-4f, P3 = 7.54978995489188e-8f;
    int k = (int)fm(phi, TWO_OVER_PI, 0.5f);
    float fk = (float)k;
    float r = fm(-fk, P1, phi);
    r = fm(-fk, P2, r);
    r = fm(-fk, P3, r);
    float z = r * r;
    float sp = fm(z, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fm(sp, z, -1.6666654611e-1f);
    float sr = fm(sp * z, r, r);
    float cp = fm(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fm(cp, z, 4.166664568298827e-2f);
    float cr = fm(cp * z, z, fm(-0.5f, z, 1.0f));
    switch (k & 3) {
        case 0: s = sr; c = cr; break;
        case 1: s = cr; c = -sr; break;
        case 2: s = -sr; c = -cr; break;
        default: s = -cr; c = sr; break;
    }
}

// ------------------------------------------------------------------------------------------
// Rays, triangles, instances
// ------------------------------------------------------------------------------------------
struct Tri { f3 v0, v1, v2; };

struct RayPrep {  // per-ray constants of the watertight test
    f3 o;
    int kx, ky, kz;
    float Sx, Sy, Sz;
};
static inline RayPrep prep_ray(f3 o, f3 d)
{
    RayPrep r;
    r.o = o;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0;
    float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    if (get(d, kz) < 0.0f) std::swap(kx, ky);
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = get(d, kx) / get(d, kz);
    r.Sy = get(d, ky) / get(d, kz);
    r.Sz = 1.0f / get(d, kz);
    return r;
}
// Watertight ray/triangle test (Woop, Benthin, Wald 2013) in the contract's op order.
// Returns true and (t, b1, b2) if tmin < t < tmax;  b1/b2 weight vertices 1/2 (OptiX convention,
// SDK/cuda/LocalGeometry.h:94).  `det_sign` gets sign of the determinant (for face culling).
template <bool INCLUSIVE = false>
static inline bool tri_hit(const RayPrep& r, const Tri& tr, float tmin, float tmax, float& t, float& b1, float& b2,
                           float* det_out = nullptr)
{
    const f3 A = tr.v0 - r.o, B = tr.v1 - r.o, C = tr.v2 - r.o;
    const float Akz = get(A, r.kz), Bkz = get(B, r.kz), Ckz = get(C, r.kz);
    const float Ax = fm(-r.Sx, Akz, get(A, r.kx)), Ay = fm(-r.Sy, Akz, get(A, r.ky));
    const float Bx = fm(-r.Sx, Bkz, get(B, r.kx)), By = fm(-r.Sy, Bkz, get(B, r.ky));
    const float Cx = fm(-r.Sx, Ckz, get(C, r.kx)), Cy = fm(-r.Sy, Ckz, get(C, r.ky));
    // edge functions: two rounded products and one subtraction (NOT fused) — this is what makes the value for a
    // shared edge exactly antisymmetric between the two triangles, i.e. watertight (Woop et al. 2013, sec. 3)
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = (U + V) + W;
    if (det == 0.0f) return false;
    const float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
    const float T = fm(W, Cz, fm(V, Bz, U * Az));
    const float tt = T / det;
    if (!(tt > tmin && (INCLUSIVE ? tt <= tmax : tt < tmax))) return false;
    t = tt;
    b1 = V / det;
    b2 = W / det;
    if (det_out) *det_out = det;
    return true;
}

// ------------------------------------------------------------------------------------------
// Oracle acceleration structure: median-split binary BVH with deliberately generous box margins,
// used only to make big scenes finish in seconds.  Because the hit rule is (t, ordinal)-minimal
// over all triangles, the answer does not depend on this structure; tests cross-check it against
// brute force.
// ------------------------------------------------------------------------------------------
template <class F>
static void parallel_rows(int rows, int threads, F fn)
{
    if (threads <= 1) { for (int y = 0; y < rows; ++y) fn(y, 0); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] { for (int y; (y = next.fetch_add(1)) < rows;) fn(y, t); });
    for (auto& th : pool) th.join();
}

struct BNode { f3 lo, hi; int left, right, first, count; };

struct Geometry {
    std::vector<Tri> tris;            // object space, ordinal = index
    std::vector<uint32_t> sbt;        // per-triangle SBT offset (material), may be empty
    uint32_t gflags = 1u;             // OptixGeometryFlags of the build input (default DISABLE_ANYHIT, as the samples set it)
    std::vector<uint8_t> tri_gflags;  // optional: OptixGeometryFlags per triangle (= of its SBT record); empty => gflags for all
    std::vector<BNode> nodes;
    std::vector<uint32_t> order;      // leaf triangle ordinals
    bool brute = true;
    uint64_t node_visits = 0, tri_tests = 0;  // single-thread instrumentation only
};

static void build_bvh(Geometry& g, int threads = 1)
{
    const size_t n = g.tris.size();
    g.nodes.clear();
    g.order.resize(n);
    for (size_t i = 0; i < n; ++i) g.order[i] = (uint32_t)i;
    g.brute = n <= 64;  // default: tiny scenes are traced brute force (the exact definition)
    if (n == 0) return;
    std::vector<f3> cen(n), lo(n), hi(n);
    parallel_rows((int)((n + 65535) / 65536), threads, [&](int row, int) {
        const size_t b = (size_t)row * 65536, e = std::min(n, b + 65536);
        for (size_t i = b; i < e; ++i) {
            const Tri& t = g.tris[i];
            lo[i] = mk(fminf(t.v0.x, fminf(t.v1.x, t.v2.x)), fminf(t.v0.y, fminf(t.v1.y, t.v2.y)), fminf(t.v0.z, fminf(t.v1.z, t.v2.z)));
            hi[i] = mk(fmaxf(t.v0.x, fmaxf(t.v1.x, t.v2.x)), fmaxf(t.v0.y, fmaxf(t.v1.y, t.v2.y)), fmaxf(t.v0.z, fmaxf(t.v1.z, t.v2.z)));
            cen[i] = (lo[i] + hi[i]) * 0.5f;
        }
    });
    g.nodes.assign(2 * n + 2, BNode{});
    std::atomic<int> next_node{1};
    struct Job { int node, first, count; };
    // processes one job; returns the two child jobs (count 0 => leaf)
    auto split = [&](Job j, Job& a, Job& b) -> bool {
        f3 blo = mk(INFINITY, INFINITY, INFINITY), bhi = mk(-INFINITY, -INFINITY, -INFINITY);
        f3 clo = blo, chi = bhi;
        for (int i = j.first; i < j.first + j.count; ++i) {
            uint32_t p = g.order[i];
            blo = mk(fminf(blo.x, lo[p].x), fminf(blo.y, lo[p].y), fminf(blo.z, lo[p].z));
            bhi = mk(fmaxf(bhi.x, hi[p].x), fmaxf(bhi.y, hi[p].y), fmaxf(bhi.z, hi[p].z));
            clo = mk(fminf(clo.x, cen[p].x), fminf(clo.y, cen[p].y), fminf(clo.z, cen[p].z));
            chi = mk(fmaxf(chi.x, cen[p].x), fmaxf(chi.y, cen[p].y), fmaxf(chi.z, cen[p].z));
        }
        // generous conservative padding: 1e-4 relative to extent and coordinate magnitude
        f3 ext = bhi - blo;
        auto pad = [](float e, float a, float b) { return 1e-4f * (e + fmaxf(fabsf(a), fabsf(b))) + 1e-30f; };
        f3 pd = mk(pad(ext.x, blo.x, bhi.x), pad(ext.y, blo.y, bhi.y), pad(ext.z, blo.z, bhi.z));
        BNode& nd = g.nodes[j.node];
        nd.lo = blo - pd;
        nd.hi = bhi + pd;
        nd.first = j.first;
        nd.count = j.count;
        nd.left = nd.right = -1;
        if (j.count <= 4) return false;
        f3 ce = chi - clo;
        int axis = ce.x >= ce.y ? (ce.x >= ce.z ? 0 : 2) : (ce.y >= ce.z ? 1 : 2);
        int mid = j.first + j.count / 2;
        std::nth_element(g.order.begin() + j.first, g.order.begin() + mid, g.order.begin() + j.first + j.count,
                         [&](uint32_t x, uint32_t y) {
                             float cx = get(cen[x], axis), cy = get(cen[y], axis);
                             return cx < cy || (cx == cy && x < y);
                         });
        int l = next_node.fetch_add(2);
        nd.left = l;
        nd.right = l + 1;
        nd.count = 0;
        a = {l, j.first, mid - j.first};
        b = {l + 1, mid, j.first + j.count - mid};
        return true;
    };
    auto run_serial = [&](Job root) {
        std::vector<Job> stack{root};
        while (!stack.empty()) {
            Job j = stack.back();
            stack.pop_back();
            Job a, b;
            if (split(j, a, b)) { stack.push_back(a); stack.push_back(b); }
        }
    };
    // top levels serially until there are enough independent subtrees, then one thread per subtree
    std::vector<Job> frontier{{0, 0, (int)n}};
    while (threads > 1 && frontier.size() < (size_t)threads * 4) {
        std::vector<Job> next;
        bool any = false;
        for (Job j : frontier) {
            Job a, b;
            if (j.count > 4096 && split(j, a, b)) { next.push_back(a); next.push_back(b); any = true; }
            else next.push_back(j);
        }
        frontier.swap(next);
        if (!any) break;
    }
    // every job left in the frontier is still unprocessed (split() replaces a job by its two children)
    parallel_rows((int)frontier.size(), threads, [&](int k, int) { run_serial(frontier[k]); });
    g.nodes.resize((size_t)next_node.load());
}

struct HitRec { float t; uint32_t prim; float b1, b2; float det; };

static inline bool box_hit(const BNode& n, f3 o, f3 inv, float tmin, float tmax)
{
    float t0 = tmin, t1 = tmax;
    const float lo[3] = {n.lo.x, n.lo.y, n.lo.z}, hi[3] = {n.hi.x, n.hi.y, n.hi.z};
    const float oo[3] = {o.x, o.y, o.z}, ii[3] = {inv.x, inv.y, inv.z};
    for (int a = 0; a < 3; ++a) {
        float ta = (lo[a] - oo[a]) * ii[a], tb = (hi[a] - oo[a]) * ii[a];
        if (ta != ta || tb != tb) continue;  // 0*inf: origin on a slab plane of a parallel ray -> do not cull
        float tn = fminf(ta, tb), tf = fmaxf(ta, tb);
        tn = tn - fabsf(tn) * 1e-4f;
        tf = tf + fabsf(tf) * 1e-4f;
        t0 = fmaxf(t0, tn);
        t1 = fminf(t1, tf);
    }
    return t0 <= t1;
}

// The cull word of a triangle test: the ray's CULL_* flags after the instance flags, any-hit override in bits 0-1 (1 = off, 2 = on).
// Precedence ray > instance > geometry, reference include/optix_types.h:1088-1108, 1794-1839.  Same function as traverse.cuh:cull_word.
static inline uint32_t cull_word(uint32_t ray_flags, uint32_t inst_flags)
{
    uint32_t c = ray_flags & 0xf0u;
    if (inst_flags & 1u) c &= ~0x30u;                                                                     // DISABLE_TRIANGLE_FACE_CULLING
    else if (inst_flags & 2u) c = (c & ~0x30u) | ((c & 0x10u) << 1) | ((c & 0x20u) >> 1);                   // FLIP_TRIANGLE_FACING
    uint32_t force = (ray_flags & 1u) ? 1u : (ray_flags & 2u) ? 2u : 0u;                                  // ray DISABLE / ENFORCE_ANYHIT
    if (!force) force = (inst_flags & 4u) ? 1u : (inst_flags & 8u) ? 2u : 0u;                             // instance DISABLE / ENFORCE_ANYHIT
    return c | force;
}

// closest hit in one geometry, object space.  `best` carries the current closest (t, prim).
template <bool ANY, bool STATS>
static bool trace_geom(Geometry& g, f3 o, f3 d, float tmin, HitRec& best, uint32_t cull_flags)
{
    const RayPrep rp = prep_ray(o, d);
    bool found = false;
    auto test = [&](uint32_t p) {
        float t, b1, b2, det;
        if (STATS) g.tri_tests++;
        // strict t < best.t, or equal t with a lower ordinal (order independence)
        if (tri_hit<true>(rp, g.tris[p], tmin, best.t, t, b1, b2, &det)) {
            if (cull_flags) {
                const uint32_t gf = g.tri_gflags.empty() ? g.gflags : (uint32_t)g.tri_gflags[p];
                // OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES (1<<4): det<0 is back facing for
                // counter-clockwise front faces seen along the ray in this formulation.
                if (!(gf & 4u)) {  // OPTIX_GEOMETRY_FLAG_DISABLE_TRIANGLE_FACE_CULLING
                    if ((cull_flags & 16u) && det < 0.0f) return;
                    if ((cull_flags & 32u) && det > 0.0f) return;
                }
                // OPTIX_RAY_FLAG_CULL_DISABLED_ANYHIT (1<<6) / CULL_ENFORCED_ANYHIT (1<<7) vs OPTIX_GEOMETRY_FLAG_DISABLE_ANYHIT (1<<0)
                const bool ah_off = (cull_flags & 3u) ? (cull_flags & 1u) != 0u : (gf & 1u) != 0u;
                if ((cull_flags & 64u) && ah_off) return;
                if ((cull_flags & 128u) && !ah_off) return;
            }
            if (t < best.t || (t == best.t && found && p < best.prim)) {
                best = {t, p, b1, b2, det};
                found = true;
            }
        }
    };
    if (g.brute) {
        for (uint32_t p = 0; p < g.tris.size(); ++p) {
            test(p);
            if (ANY && found) return true;
        }
        return found;
    }
    const f3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const BNode& n = g.nodes[stack[--sp]];
        if (STATS) g.node_visits++;
        if (!box_hit(n, o, inv, tmin, best.t * 1.0001f + 1e-30f)) continue;
        if (n.left < 0) {
            for (int i = n.first; i < n.first + n.count; ++i) {
                test(g.order[i]);
                if (ANY && found) return true;
            }
        } else {
            stack[sp++] = n.left;
            stack[sp++] = n.right;
        }
    }
    return found;
}

// A scene = one geometry, optionally behind one or more instance transforms (IAS of the glTF path,
// SDK/sutil/Scene.cpp:1134-1212).  Rays are carried into object space with the inverse 3x4 and `t`
// is shared between spaces (OptiX semantics).
struct Instance { float m[12]; float inv[12]; int geom; uint32_t flags = 0u, mask = 1u; };  // OptixInstance flags / visibilityMask

struct Scene {
    std::vector<Geometry> geoms;
    std::vector<Instance> insts;  // empty => geoms[0] used directly (GAS handle launched directly)
};

// inverse of an affine 3x4 (row major), fixed op order: adjugate / det, then -Ainv*t
static void invert34(const float* m, float* inv)
{
    const float a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const float c00 = fm(e, i, -(f * h)), c01 = fm(f, g, -(d * i)), c02 = fm(d, h, -(e * g));
    const float det = fm(c, c02, fm(b, c01, a * c00));
    const float r = 1.0f / det;
    inv[0] = c00 * r;                     inv[1] = fm(c, h, -(b * i)) * r;   inv[2] = fm(b, f, -(c * e)) * r;
    inv[4] = c01 * r;                     inv[5] = fm(a, i, -(c * g)) * r;   inv[6] = fm(c, d, -(a * f)) * r;
    inv[8] = c02 * r;                     inv[9] = fm(b, g, -(a * h)) * r;   inv[10] = fm(a, e, -(b * d)) * r;
    const float tx = m[3], ty = m[7], tz = m[11];
    inv[3]  = -fm(inv[2], tz, fm(inv[1], ty, inv[0] * tx));
    inv[7]  = -fm(inv[6], tz, fm(inv[5], ty, inv[4] * tx));
    inv[11] = -fm(inv[10], tz, fm(inv[9], ty, inv[8] * tx));
}
static inline f3 xform_point(const float* m, f3 p)
{
    return mk(fm(m[2], p.z, fm(m[1], p.y, m[0] * p.x)) + m[3], fm(m[6], p.z, fm(m[5], p.y, m[4] * p.x)) + m[7],
              fm(m[10], p.z, fm(m[9], p.y, m[8] * p.x)) + m[11]);
}
static inline f3 xform_vec(const float* m, f3 v)
{
    return mk(fm(m[2], v.z, fm(m[1], v.y, m[0] * v.x)), fm(m[6], v.z, fm(m[5], v.y, m[4] * v.x)),
              fm(m[10], v.z, fm(m[9], v.y, m[8] * v.x)));
}
// normal object->world = transpose(inverse) * n
static inline f3 xform_normal(const float* inv, f3 n)
{
    return mk(fm(inv[8], n.z, fm(inv[4], n.y, inv[0] * n.x)), fm(inv[9], n.z, fm(inv[5], n.y, inv[1] * n.x)),
              fm(inv[10], n.z, fm(inv[6], n.y, inv[2] * n.x)));
}

struct SceneHit { float t; uint32_t inst, prim; float b1, b2; bool hit; };

template <bool ANY, bool STATS = false>
static SceneHit trace_scene(Scene& s, f3 o, f3 d, float tmin, float tmax, uint32_t ray_flags = 0)
{
    SceneHit r{tmax, 0xffffffffu, 0xffffffffu, 0.f, 0.f, false};
    if (s.insts.empty()) {
        HitRec b{tmax, 0, 0, 0, 0};
        if (trace_geom<ANY, STATS>(s.geoms[0], o, d, tmin, b, cull_word(ray_flags, 0u))) r = {b.t, 0, b.prim, b.b1, b.b2, true};
        return r;
    }
    for (uint32_t k = 0; k < s.insts.size(); ++k) {
        const Instance& in = s.insts[k];
        if (!(in.mask & ((ray_flags >> 16 ^ 1u) & 0xffu))) continue;  // OptixVisibilityMask of the ray in bits 16-23 of the flags word, stored XOR 1 (absent = mask 1)
        HitRec b{r.t, 0, 0, 0, 0};
        // a later instance only wins with strictly smaller t (lower instance index wins ties)
        if (trace_geom<ANY, STATS>(s.geoms[in.geom], xform_point(in.inv, o), xform_vec(in.inv, d), tmin, b, cull_word(ray_flags, in.flags))) {
            if (b.t < r.t) { r = {b.t, k, b.prim, b.b1, b.b2, true}; if (ANY) return r; }
        }
    }
    return r;
}

// ------------------------------------------------------------------------------------------
// sRGB quantisation — SDK/cuda/helpers.h:36-64 (powf: libm here, fast-math on device => ±1 LSB)
// ------------------------------------------------------------------------------------------
static inline float to_srgb1(float c)
{
    float powed = powf(c, 1.0f / 2.4f);
    return c < 0.0031308f ? 12.92f * c : fm(1.055f, powed, -0.055f);
}
static inline uint8_t quant8(float x)
{
    x = clampf(x, 0.0f, 1.0f);
    uint32_t q = (uint32_t)(x * 256.0f);
    return (uint8_t)(q < 255u ? q : 255u);
}
static inline void make_color(f3 c, uint8_t* out)
{
    out[0] = quant8(to_srgb1(clampf(c.x, 0.f, 1.f)));
    out[1] = quant8(to_srgb1(clampf(c.y, 0.f, 1.f)));
    out[2] = quant8(to_srgb1(clampf(c.z, 0.f, 1.f)));
    out[3] = 255;
}

// ------------------------------------------------------------------------------------------
// Cornell path tracer — SDK/optixPathTracer/optixPathTracer.cu:249-413 (mode 0) and
// SDK/optixMultiGPU/optixMultiGPU.cu:214-385 (mode 1: depth cap 3, no RR, sticky emitted/radiance)
// ------------------------------------------------------------------------------------------
struct PTParams {
    uint32_t subframe_index;
    int32_t width, height, spl;
    float eye[3], U[3], V[3], W[3];
    float light_corner[3], light_v1[3], light_v2[3], light_normal[3], light_emission[3];
    float bg[3];
    int32_t nmat;
    int32_t groups;         // sample groups (b200rt_pt_options.sample_groups): 0 / 1 = the reference's flat summation order
    const float* emission;  // nmat*3
    const float* diffuse;   // nmat*3
};

struct Onb {
    f3 t, b, n;
    explicit Onb(f3 normal)  // optixPathTracer.cu:47-78
    {
        n = normal;
        if (fabsf(n.x) > fabsf(n.z)) b = mk(-n.y, n.x, 0.0f);
        else b = mk(0.0f, -n.z, n.y);
        b = normalize(b);
        t = cross(b, n);
    }
    f3 inverse_transform(f3 p) const
    {
        return mk(fm(p.z, n.x, fm(p.y, b.x, p.x * t.x)), fm(p.z, n.y, fm(p.y, b.y, p.x * t.y)),
                  fm(p.z, n.z, fm(p.y, b.z, p.x * t.z)));
    }
};

static inline f3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }

// One pixel of one launch.  Returns the per-launch mean radiance (before the running-mean lerp);
// counts traced segments (radiance + shadow) into *segs.
static f3 pt_pixel(Scene& sc, const PTParams& P, int px, int py, int mode, uint64_t* segs)
{
    const int w = P.width, h = P.height;
    const f3 eye = ld3(P.eye), U = ld3(P.U), V = ld3(P.V), W = ld3(P.W);
    const f3 lc = ld3(P.light_corner), lv1 = ld3(P.light_v1), lv2 = ld3(P.light_v2), ln = ld3(P.light_normal),
             le = ld3(P.light_emission);
    uint32_t seed = tea4((uint32_t)(py * w + px), P.subframe_index);
    f3 result = mk(0, 0, 0), total = mk(0, 0, 0);
    uint64_t nseg = 0;
    Geometry& g0 = sc.geoms[0];
    // sample groups: group g sums samples [g*spl/G, (g+1)*spl/G) from zero; the pixel is ((s_0 + s_1) + ...) + s_{G-1}
    const int G = P.groups > 1 ? P.groups : 1;
    int group = 0;
    for (int i = 0; i < P.spl; ++i) {
        while (group + 1 < G && i >= (int)(((int64_t)(group + 1) * P.spl) / G)) {
            total = group == 0 ? result : total + result;
            result = mk(0, 0, 0);
            ++group;
        }
        const float jx = rnd(seed), jy = rnd(seed);
        const float dx = fm(2.0f, ((float)px + jx) / (float)w, -1.0f);
        const float dy = fm(2.0f, ((float)py + jy) / (float)h, -1.0f);
        f3 dir = normalize(mk(fm(dy, V.x, dx * U.x) + W.x, fm(dy, V.y, dx * U.y) + W.y, fm(dy, V.z, dx * U.z) + W.z));
        f3 org = eye;
        f3 att = mk(1, 1, 1), emitted = mk(0, 0, 0), radiance = mk(0, 0, 0);
        uint32_t pseed = seed;
        int depth = 0;
        bool count_emitted = true;
        for (;;) {
            SceneHit hit = trace_scene<false>(sc, org, dir, 0.01f, 1e16f);
            ++nseg;
            bool done;
            if (!hit.hit) {  // __miss__radiance
                radiance = ld3(P.bg);
                if (mode == 0) emitted = mk(0, 0, 0);  // MultiGPU miss leaves prd.emitted untouched
                done = true;
            } else {  // __closesthit__radiance
                const Tri& tr = g0.tris[hit.prim];
                const uint32_t mat = g0.sbt.empty() ? 0u : g0.sbt[hit.prim];
                const f3 N0 = normalize(cross(tr.v1 - tr.v0, tr.v2 - tr.v0));
                // faceforward(N0, -dir, N0) = N0 * copysignf(1, dot(-dir, N0))
                const f3 N = N0 * copysignf(1.0f, dot(neg(dir), N0));
                const f3 Pp = mk(fm(hit.t, dir.x, org.x), fm(hit.t, dir.y, org.y), fm(hit.t, dir.z, org.z));
                const bool emit_now = (mode == 0) ? (depth == 0) : count_emitted;
                emitted = emit_now ? ld3(P.emission + 3 * mat) : mk(0, 0, 0);
                const float z1 = rnd(pseed), z2 = rnd(pseed);
                float s, c;
                det_sincos(6.2831855f * z2, s, c);
                const float r = sqrtf(z1);
                f3 w_in = mk(r * c, r * s, 0.0f);
                w_in.z = sqrtf(fmaxf(0.0f, fm(-w_in.y, w_in.y, fm(-w_in.x, w_in.x, 1.0f))));
                Onb onb(N);
                const f3 ndir = onb.inverse_transform(w_in);
                att = att * ld3(P.diffuse + 3 * mat);
                count_emitted = false;
                const float l1 = rnd(pseed), l2 = rnd(pseed);
                const f3 lp = mk(fm(lv2.x, l2, fm(lv1.x, l1, lc.x)), fm(lv2.y, l2, fm(lv1.y, l1, lc.y)),
                                 fm(lv2.z, l2, fm(lv1.z, l1, lc.z)));
                const f3 Ld = lp - Pp;
                const float Ldist = length(Ld);
                const f3 L = normalize(Ld);
                const float nDl = dot(N, L);
                const float LnDl = -dot(ln, L);
                float weight = 0.0f;
                if (nDl > 0.0f && LnDl > 0.0f) {
                    ++nseg;
                    SceneHit sh = trace_scene<true>(sc, Pp, L, 0.01f, Ldist - 0.01f);
                    if (!sh.hit) {
                        const float A = length(cross(lv1, lv2));
                        weight = ((nDl * LnDl) * A) / ((3.14159265358979323846f * Ldist) * Ldist);
                    }
                }
                if (mode == 0) radiance = le * weight;
                else radiance = mk(fm(le.x, weight, radiance.x), fm(le.y, weight, radiance.y), fm(le.z, weight, radiance.z));
                org = Pp;
                dir = ndir;
                done = false;
            }
            result = result + emitted;
            result = mk(fm(radiance.x, att.x, result.x), fm(radiance.y, att.y, result.y), fm(radiance.z, att.z, result.z));
            if (mode == 0) {
                const float p = dot(att, mk(0.30f, 0.59f, 0.11f));
                if (done || rnd(pseed) > p) break;
                att = mk(att.x / p, att.y / p, att.z / p);
            } else {
                if (done || depth >= 3) break;
            }
            ++depth;
        }
    }
    if (segs) *segs += nseg;
    // close the open group and add the (empty, zero) groups that follow it
    for (; group < G; ++group) {
        total = group == 0 ? result : total + result;
        result = mk(0, 0, 0);
    }
    result = total;
    const float spl = (float)P.spl;
    return mk(result.x / spl, result.y / spl, result.z / spl);
}

}  // namespace

// ============================================================================================
// C ABI (ctypes)
// ============================================================================================
extern "C" {

uint32_t orc_tea4(uint32_t v0, uint32_t v1) { return tea4(v0, v1); }
uint32_t orc_lcg(uint32_t* state) { return lcg(*state); }
float orc_rnd(uint32_t* state) { return rnd(*state); }
void orc_sincos(float phi, float* s, float* c) { det_sincos(phi, *s, *c); }

// SDK/sutil/WorkDistribution.h:50-81
int orc_wd_num_samples(int w, int h, int ngpu)
{
    const int sw = 8 * ngpu, sh = 4;
    const int cols = w / sw + (w % sw ? 1 : 0), rows = h / sh + (h % sh ? 1 : 0);
    return rows * cols * 32;
}
void orc_wd_sample_pixel(int w, int h, int ngpu, int gpu, int sample, int* xy)
{
    (void)h;
    const int sw = 8 * ngpu;
    const int cols = w / sw + (w % sw ? 1 : 0);
    const int strip = sample / 32;
    const int sy = strip / cols, sx = strip - sy * cols;
    const int in_tile = sample - strip * 32;
    const int ty = in_tile / 8, tx = in_tile - ty * 8;
    const int xoff = ((gpu + sy % ngpu) % ngpu) * 8;
    xy[0] = sx * sw + tx + xoff;
    xy[1] = sy * 4 + ty;
}

// SDK/sutil/Camera.cpp:34-46 (m_fod == 1)
void orc_camera_uvw(const float* eye, const float* lookat, const float* up, float fovy, float aspect, float* uvw)
{
    // the reference's host vec_math is compiled by the host compiler without contraction
    auto hdot = [](f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
    auto hcross = [](f3 a, f3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); };
    auto hnorm = [&](f3 v) { float inv = 1.0f / sqrtf(hdot(v, v)); return v * inv; };
    f3 W = (ld3(lookat) - ld3(eye)) * 1.0f;
    float wlen = sqrtf(hdot(W, W));
    f3 U = hnorm(hcross(W, ld3(up)));
    f3 V = hnorm(hcross(U, W));
    float vlen = wlen * tanf(0.5f * fovy * 3.14159265358979323846f / 180.0f);
    V = V * vlen;
    float ulen = vlen * aspect;
    U = U * ulen;
    uvw[0] = U.x; uvw[1] = U.y; uvw[2] = U.z; uvw[3] = V.x; uvw[4] = V.y; uvw[5] = V.z; uvw[6] = W.x; uvw[7] = W.y; uvw[8] = W.z;
}

void orc_make_color(const float* rgb, int n, uint8_t* out)
{
    for (int i = 0; i < n; ++i) make_color(ld3(rgb + 3 * i), out + 4 * i);
}

// ---- scenes ---------------------------------------------------------------------------------
// verts: ntri*9 floats (object space).  sbt: per-triangle SBT offset or NULL.
void* orc_scene_create(const float* verts, int64_t ntri, const uint32_t* sbt)
{
    Scene* s = new Scene();
    s->geoms.emplace_back();
    Geometry& g = s->geoms[0];
    g.tris.resize((size_t)ntri);
    for (int64_t i = 0; i < ntri; ++i) {
        const float* v = verts + 9 * i;
        g.tris[(size_t)i] = {mk(v[0], v[1], v[2]), mk(v[3], v[4], v[5]), mk(v[6], v[7], v[8])};
    }
    if (sbt) g.sbt.assign(sbt, sbt + ntri);
    build_bvh(g, (int)std::max(1u, std::thread::hardware_concurrency()));
    return s;
}
// add one instance of geometry 0 with a row-major 3x4 object->world transform
uint32_t orc_cull_word(uint32_t ray_flags, uint32_t inst_flags) { return cull_word(ray_flags, inst_flags); }
void orc_scene_add_instance_ex(void* scene, const float* m34, uint32_t flags, uint32_t mask)
{
    Scene* s = (Scene*)scene;
    Instance in;
    memcpy(in.m, m34, sizeof(in.m));
    invert34(in.m, in.inv);
    in.geom = 0;
    in.flags = flags;
    in.mask = mask;
    s->insts.push_back(in);
}
void orc_scene_add_instance(void* scene, const float* m34) { orc_scene_add_instance_ex(scene, m34, 0u, 1u); }
void orc_scene_set_brute(void* scene, int brute)
{
    Scene* s = (Scene*)scene;
    for (auto& g : s->geoms) g.brute = brute != 0 || g.nodes.empty();
}
void orc_scene_destroy(void* scene) { delete (Scene*)scene; }
void orc_scene_set_geometry_flags(void* scene, uint32_t gflags) { for (auto& g : ((Scene*)scene)->geoms) { g.gflags = gflags; g.tri_gflags.clear(); } }
// OptixGeometryFlags per triangle (the flags of the SBT record each triangle belongs to), one byte each
void orc_scene_set_triangle_flags(void* scene, const uint8_t* flags)
{
    Geometry& g = ((Scene*)scene)->geoms[0];
    g.tri_gflags.assign(flags, flags + g.tris.size());
}
void orc_invert34(const float* m, float* inv) { invert34(m, inv); }

// rays: n * 8 floats {ox,oy,oz,tmin,dx,dy,dz,tmax} (the reference Ray struct,
// SDK/optixRaycasting/optixRaycastingKernels.h:35-41).  out: n * {t, prim, inst, b1, b2} as 5 x u32 bit patterns
// (miss: t = -1, prim = inst = 0xffffffff).  any_hit: out[i*5] = 1/0 only.
void orc_trace(void* scene, const float* rays, int64_t n, uint32_t* out, int any_hit, uint32_t ray_flags, int threads,
               uint64_t* stats /* [2] node visits, tri tests (threads==1 only) or NULL */)
{
    Scene* s = (Scene*)scene;
    const int chunk = 4096;
    const int rows = (int)((n + chunk - 1) / chunk);
    if (stats) for (auto& g : s->geoms) g.node_visits = g.tri_tests = 0;
    parallel_rows(rows, stats ? 1 : threads, [&](int row, int) {
        const int64_t b = (int64_t)row * chunk, e = std::min<int64_t>(n, b + chunk);
        for (int64_t i = b; i < e; ++i) {
            const float* r = rays + 8 * i;
            uint32_t* o = out + 5 * i;
            f3 org = mk(r[0], r[1], r[2]), dir = mk(r[4], r[5], r[6]);
            if (any_hit) {
                SceneHit h = stats ? trace_scene<true, true>(*s, org, dir, r[3], r[7], ray_flags)
                                   : trace_scene<true, false>(*s, org, dir, r[3], r[7], ray_flags);
                o[0] = h.hit ? 1u : 0u; o[1] = o[2] = o[3] = o[4] = 0;
            } else {
                SceneHit h = stats ? trace_scene<false, true>(*s, org, dir, r[3], r[7], ray_flags)
                                   : trace_scene<false, false>(*s, org, dir, r[3], r[7], ray_flags);
                float t = h.hit ? h.t : -1.0f;
                memcpy(o, &t, 4);
                o[1] = h.prim; o[2] = h.inst;
                memcpy(o + 3, &h.b1, 4); memcpy(o + 4, &h.b2, 4);
            }
        }
    });
    if (stats) { stats[0] = stats[1] = 0; for (auto& g : s->geoms) { stats[0] += g.node_visits; stats[1] += g.tri_tests; } }
}

// ---- optixRaycasting helpers (SDK/optixRaycasting/optixRaycastingKernels.cu:42-115) ------------
// host-side scalars exactly as createRaysOrthoOnDevice computes them (host compiler, no contraction)
void orc_raycast_ortho_scalars(const float* bbmin, const float* bbmax, int width, int height, float padding, float* out5)
{
    const float sx = bbmax[0] - bbmin[0], sy = bbmax[1] - bbmin[1], sz = bbmax[2] - bbmin[2];
    float dx = sx * (1 + 2 * padding) / width;
    float dy = sy * (1 + 2 * padding) / height;
    float x0 = bbmin[0] - sx * padding + dx / 2;
    float y0 = bbmin[1] - sy * padding + dy / 2;
    float z = bbmin[2] - fmaxf(sz, 1.0f) * .001f;
    out5[0] = x0; out5[1] = y0; out5[2] = z; out5[3] = dx; out5[4] = dy;
}
// device kernel restated: origin = (x0 + ix*dx, y0 + iy*dy, z); nvcc contracts to fma(ix, dx, x0)
void orc_raycast_create_rays(float* rays, int width, int height, float x0, float y0, float z, float dx, float dy)
{
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            float* r = rays + 8 * ((size_t)y * width + x);
            r[0] = fm((float)x, dx, x0); r[1] = fm((float)y, dy, y0); r[2] = z; r[3] = 0.0f;
            r[4] = 0.0f; r[5] = 0.0f; r[6] = 1.0f; r[7] = 1e34f;
        }
}
void orc_raycast_translate(float* rays, int64_t n, const float* off)
{
    for (int64_t i = 0; i < n; ++i) { rays[8 * i] += off[0]; rays[8 * i + 1] += off[1]; rays[8 * i + 2] += off[2]; }
}
// Closest hit + the reference's closest-hit program (SDK/optixRaycasting/optixRaycasting.cu:45-86):
// Hit = {float(unsigned(t)) , N} with N the interpolated shading normal taken object->world and
// normalised (SDK/cuda/LocalGeometry.h:94-131); miss = {-1,(1,0,0)}.  normals: per-triangle 9 floats
// (already de-indexed by the caller) or NULL => geometric normal.
void orc_raycast_hits(void* scene, const float* rays, int64_t n, const float* normals, float* hits /* n*4 */,
                      uint32_t* ext /* n*5 or NULL */, int threads)
{
    Scene* s = (Scene*)scene;
    std::vector<uint32_t> tmp;
    if (!ext) { tmp.resize((size_t)n * 5); ext = tmp.data(); }
    orc_trace(scene, rays, n, ext, 0, 0, threads, nullptr);
    Geometry& g = s->geoms[0];
    for (int64_t i = 0; i < n; ++i) {
        float t; memcpy(&t, ext + 5 * i, 4);
        float* h = hits + 4 * i;
        if (t < 0.0f) { h[0] = -1.0f; h[1] = 1.0f; h[2] = 0.0f; h[3] = 0.0f; continue; }
        const uint32_t prim = ext[5 * i + 1], inst = ext[5 * i + 2];
        float b1, b2; memcpy(&b1, ext + 5 * i + 3, 4); memcpy(&b2, ext + 5 * i + 4, 4);
        f3 N;
        if (normals) {
            const float* nn = normals + 9 * (size_t)prim;
            const float b0 = (1.0f - b1) - b2;
            // (1-b1-b2)*N0 + b1*N1 + b2*N2 with nvcc's contraction: fma(b2,N2, fma(b1,N1, b0*N0))
            N = mk(fm(b2, nn[6], fm(b1, nn[3], b0 * nn[0])), fm(b2, nn[7], fm(b1, nn[4], b0 * nn[1])),
                   fm(b2, nn[8], fm(b1, nn[5], b0 * nn[2])));
        } else {
            const Tri& tr = g.tris[prim];
            N = cross(tr.v1 - tr.v0, tr.v2 - tr.v0);
        }
        if (!s->insts.empty()) N = xform_normal(s->insts[inst].inv, N);
        N = normalize(N);
        h[0] = (float)(unsigned int)t;  // the `const unsigned int t = optixGetRayTmax()` quirk
        h[1] = N.x; h[2] = N.y; h[3] = N.z;
    }
}
// shadeHitsKernel: 0.5*N + 0.5 (contracted: fma(0.5,N,0.5)), background 0.2
void orc_raycast_shade(const float* hits, int64_t n, float* image /* n*3 */)
{
    for (int64_t i = 0; i < n; ++i) {
        const float* h = hits + 4 * i;
        float* o = image + 3 * i;
        if (h[0] < 0.0f) { o[0] = o[1] = o[2] = 0.2f; }
        else { o[0] = fm(0.5f, h[1], 0.5f); o[1] = fm(0.5f, h[2], 0.5f); o[2] = fm(0.5f, h[3], 0.5f); }
    }
}

// ---- Cornell path tracer -----------------------------------------------------------------------
struct orc_pt_params {
    uint32_t subframe_index;
    int32_t width, height, samples_per_launch;
    float eye[3], U[3], V[3], W[3];
    float light_corner[3], light_v1[3], light_v2[3], light_normal[3], light_emission[3];
    float bg[3];
    int32_t nmat;
    int32_t mode;  // 0 = optixPathTracer (Russian roulette), 1 = optixMultiGPU (depth cap 3)
    int32_t groups;  // sample groups, 0 / 1 = reference summation order
};
// Renders rows [y0,y1) x columns [x0,x1).  accum: width*height*4 floats, read for the running mean
// when subframe_index > 0 and written (optixPathTracer.cu:310-319); frame: width*height*4 bytes.
// Returns the number of traced segments.
uint64_t orc_pathtrace(void* scene, const orc_pt_params* p, const float* emission, const float* diffuse, float* accum,
                       uint8_t* frame, int x0, int y0, int x1, int y1, int threads)
{
    Scene* s = (Scene*)scene;
    PTParams P;
    P.subframe_index = p->subframe_index; P.width = p->width; P.height = p->height; P.spl = p->samples_per_launch;
    memcpy(P.eye, p->eye, 12); memcpy(P.U, p->U, 12); memcpy(P.V, p->V, 12); memcpy(P.W, p->W, 12);
    memcpy(P.light_corner, p->light_corner, 12); memcpy(P.light_v1, p->light_v1, 12); memcpy(P.light_v2, p->light_v2, 12);
    memcpy(P.light_normal, p->light_normal, 12); memcpy(P.light_emission, p->light_emission, 12); memcpy(P.bg, p->bg, 12);
    P.nmat = p->nmat; P.groups = p->groups; P.emission = emission; P.diffuse = diffuse;
    const int T = std::max(1, threads);
    std::vector<uint64_t> segs((size_t)T, 0);
    parallel_rows(y1 - y0, T, [&](int row, int tid) {
        const int y = y0 + row;
        for (int x = x0; x < x1; ++x) {
            f3 c = pt_pixel(*s, P, x, y, p->mode, &segs[(size_t)tid]);
            const size_t idx = (size_t)y * P.width + x;
            if (P.subframe_index > 0) {
                const float a = 1.0f / (float)(P.subframe_index + 1);
                const f3 prev = ld3(accum + 4 * idx);
                // lerp(a,b,t) = a + t*(b-a)  -> fma(t, b-a, a)
                c = mk(fm(a, c.x - prev.x, prev.x), fm(a, c.y - prev.y, prev.y), fm(a, c.z - prev.z, prev.z));
            }
            accum[4 * idx] = c.x; accum[4 * idx + 1] = c.y; accum[4 * idx + 2] = c.z; accum[4 * idx + 3] = 1.0f;
            if (frame) make_color(c, frame + 4 * idx);
        }
    });
    uint64_t tot = 0;
    for (auto v : segs) tot += v;
    return tot;
}

}  // extern "C"

// ---- synthetic tessellated scene (BASELINE.json configs[4]) ------------------------------------------
// Independent restatement of the product's procedural mesh (optix_raytracer_b200/csrc/pathtracer.cu,
// synth_*): same closed-form vertex functions in the contract's named operations, so the vertices are
// bit-identical to the GPU generator (checked by tests/test_gpu_parity.py).  The reference has no such
// asset; SURVEY.md §8(d) C5 defines it as a procedural stand-in.
namespace {
struct SynthLayout { uint64_t total, blob_each, wall_each, light; uint32_t rows, cols, grid; };
static SynthLayout synth_layout(uint64_t total)
{
    SynthLayout s;
    s.total = total;
    const uint64_t t = total > 64 ? total - 2 : 0;
    double per_blob = (double)t * 0.9 / 27.0 / 4.0;
    uint32_t rows = (uint32_t)floor(sqrt(per_blob > 0 ? per_blob : 0));
    if (rows < 2) rows = t >= 27 * 16 ? 2 : 0;
    s.rows = rows;
    s.cols = 2 * rows;
    s.blob_each = 2ull * rows * s.cols;
    const uint64_t used = 27 * s.blob_each;
    s.grid = (uint32_t)floor(sqrt((double)(t > used ? t - used : 0) / 10.0));
    s.wall_each = 2ull * s.grid * s.grid;
    s.light = total - used - 5 * s.wall_each;
    return s;
}
static f3 synth_blob_vertex(int b, uint32_t i, uint32_t j, uint32_t rows, uint32_t cols, uint32_t seed)
{
    const int bx = b % 3, by = (b / 3) % 3, bz = b / 9;
    const f3 ctr = mk(139.0f + 139.0f * (float)bx, 110.0f + 160.0f * (float)by, 140.0f + 140.0f * (float)bz);
    const float rad = 48.0f;
    if (i == 0) return mk(ctr.x, ctr.y + rad, ctr.z);
    if (i == rows) return mk(ctr.x, ctr.y - rad, ctr.z);
    j = j % cols;
    const float theta = (3.14159265358979f * (float)i) / (float)rows;
    const float phi = (6.28318530717959f * (float)j) / (float)cols;
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
    return mk(fm(rr * st, cp, ctr.x), fm(rr, ct, ctr.y), fm(rr * st, sp, ctr.z));
}
static f3 synth_wall_vertex(int wall, uint32_t i, uint32_t j, uint32_t grid)
{
    const float X = 556.0f, Y = 548.8f, Z = 559.2f;
    const float u = (float)i / (float)grid, v = (float)j / (float)grid;
    switch (wall) {
        case 0: return mk(X * u, 0.0f, Z * v);
        case 1: return mk(X * u, Y, Z * v);
        case 2: return mk(X * u, Y * v, Z);
        case 3: return mk(0.0f, Y * u, Z * v);
        default: return mk(X, Y * u, Z * v);
    }
}
}  // namespace

extern "C" void orc_synth_mesh(uint64_t total, uint32_t seed, float* verts /* total*9 */, uint32_t* mats, int threads)
{
    const SynthLayout s = synth_layout(total);
    const uint64_t blob_total = 27 * s.blob_each, wall_total = 5 * s.wall_each;
    const uint64_t chunk = 1 << 16;
    parallel_rows((int)((total + chunk - 1) / chunk), threads, [&](int row, int) {
        const uint64_t b0 = (uint64_t)row * chunk, e0 = std::min<uint64_t>(total, b0 + chunk);
        for (uint64_t t = b0; t < e0; ++t) {
            f3 a, b, c;
            uint32_t mat = 0;
            if (t < blob_total) {
                const int bi = (int)(t / s.blob_each);
                const uint64_t r = t - (uint64_t)bi * s.blob_each, quad = r >> 1;
                const uint32_t i = (uint32_t)(quad / s.cols), j = (uint32_t)(quad % s.cols);
                const f3 p00 = synth_blob_vertex(bi, i, j, s.rows, s.cols, seed), p01 = synth_blob_vertex(bi, i, j + 1, s.rows, s.cols, seed);
                const f3 p10 = synth_blob_vertex(bi, i + 1, j, s.rows, s.cols, seed), p11 = synth_blob_vertex(bi, i + 1, j + 1, s.rows, s.cols, seed);
                if (r & 1) { a = p00; b = p11; c = p01; } else { a = p00; b = p10; c = p11; }
                mat = (bi % 5 == 1) ? 1u : ((bi % 5 == 3) ? 2u : 0u);
            } else if (t < blob_total + wall_total) {
                const uint64_t r0 = t - blob_total;
                const int w = (int)(r0 / s.wall_each);
                const uint64_t r = r0 - (uint64_t)w * s.wall_each, quad = r >> 1;
                const uint32_t i = (uint32_t)(quad / s.grid), j = (uint32_t)(quad % s.grid);
                const f3 p00 = synth_wall_vertex(w, i, j, s.grid), p01 = synth_wall_vertex(w, i, j + 1, s.grid);
                const f3 p10 = synth_wall_vertex(w, i + 1, j, s.grid), p11 = synth_wall_vertex(w, i + 1, j + 1, s.grid);
                if (r & 1) { a = p00; b = p11; c = p01; } else { a = p00; b = p10; c = p11; }
                mat = w == 3 ? 1u : (w == 4 ? 2u : 0u);
            } else {
                const uint64_t r = t - blob_total - wall_total, K = s.light / 2;
                mat = 3u;
                if (K == 0 || (r >> 1) >= K) {
                    a = b = c = mk(343.0f, 548.6f, 227.0f);
                } else {
                    const uint64_t q = r >> 1;
                    const float x0 = fm(-130.0f, (float)q / (float)K, 343.0f);
                    const float x1 = (q + 1 == K) ? 213.0f : fm(-130.0f, (float)(q + 1) / (float)K, 343.0f);
                    const f3 p00 = mk(x0, 548.6f, 227.0f), p01 = mk(x0, 548.6f, 332.0f), p10 = mk(x1, 548.6f, 227.0f), p11 = mk(x1, 548.6f, 332.0f);
                    if (r & 1) { a = p00; b = p11; c = p01; } else { a = p00; b = p10; c = p11; }
                }
            }
            float* o = verts + 9 * t;
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = c.x; o[7] = c.y; o[8] = c.z;
            mats[t] = mat;
        }
    });
}

// ------------------------------------------------------------------------------------------
// imgui_test ("playground") — SDK/imgui_test/optixTriangle.cu:103-268, camera.h:127-144, light.h:19-40,
// volumetric_light.h:22-27, directional_light.h:19-24, point_light.h, diffuse.h:8-12.  Scalar restatement in
// the arithmetic contract (same op order as optix_raytracer_b200/csrc/playground.cu).
// ------------------------------------------------------------------------------------------
namespace {
struct PGCam { float eye[3], lookat[3], up[3]; uint8_t ortho; uint8_t pad[3]; float fov, fd, aperture, speed; float u[3], v[3], w[3]; };
static_assert(sizeof(PGCam) == 92, "imgui_test Camera");
struct PGLightO { float a[3], b[3], c[4]; int32_t tag; };
static_assert(sizeof(PGLightO) == 44, "imgui_test LightVariant");

static inline f3 pg_wi(const PGLightO& l, f3 p, uint32_t& seed)
{
    if (l.tag == 0) return mk(l.a[0] - p.x, l.a[1] - p.y, l.a[2] - p.z);
    if (l.tag == 1) {
        const float j = l.c[0];
        const float r0 = rnd(seed), r1 = rnd(seed), r2 = rnd(seed);
        return mk(fm(j, r0, l.a[0]), fm(j, r1, l.a[1]), fm(j, r2, l.a[2]));
    }
    const float r = l.b[0];
    const float r0 = rnd(seed), r1 = rnd(seed), r2 = rnd(seed);
    return mk(fm(r, r0, l.a[0]) - p.x, fm(r, r1, l.a[1]) - p.y, fm(r, r2, l.a[2]) - p.z);
}
static inline f3 pg_lumi(const PGLightO& l) { return l.tag == 2 ? mk(l.b[1], l.b[2], l.c[0]) : mk(l.b[0], l.b[1], l.b[2]); }
}  // namespace

extern "C" {

// film: width*height*3 floats (read when !dirty, written); image: width*height*4 bytes or NULL.  normals: ntri*9 floats,
// mat_indices: ntri ints, materials: nmat*3 floats.  Rows [y0,y1).  Returns rays traced (primary + probes).
uint64_t orc_playground(void* scene, const void* camera92, const void* lights44, int nlights, const float* materials, const float* normals,
                        const int32_t* mat_indices, int width, int height, uint32_t spf, uint32_t dt, int dirty, float* film, uint8_t* image,
                        int y0, int y1, int threads)
{
    Scene* s = (Scene*)scene;
    PGCam cam;
    memcpy(&cam, camera92, sizeof cam);
    const PGLightO* L = (const PGLightO*)lights44;
    const f3 cu = ld3(cam.u), cv = ld3(cam.v), cw = ld3(cam.w), eye = ld3(cam.eye);
    const int T = std::max(1, threads);
    std::vector<uint64_t> nrays((size_t)T, 0);
    parallel_rows(y1 - y0, T, [&](int row, int tid) {
        const int iy = y0 + row;
        for (int ix = 0; ix < width; ++ix) {
            uint32_t seed = tea4((uint32_t)ix + (uint32_t)width * (uint32_t)iy, dt);
            f3 result = mk(0.f, 0.f, 0.f);
            for (uint32_t smp = 0; smp < spf; ++smp) {
                float dx = fm(2.0f, (float)ix / (float)width, -1.0f), dy = fm(2.0f, (float)iy / (float)height, -1.0f);
                f3 org, dir;
                if (cam.ortho) {
                    dir = normalize(mk(fm(dy, cv.x, dx * cu.x) + cw.x, fm(dy, cv.y, dx * cu.y) + cw.y, fm(dy, cv.z, dx * cu.z) + cw.z));
                    org = mk(fm(dy, cv.x, fm(dx, cu.x, eye.x)), fm(dy, cv.y, fm(dx, cu.y, eye.y)), fm(dy, cv.z, fm(dx, cu.z, eye.z)));
                } else {
                    const float lx = (rnd(seed) - 0.5f) * cam.aperture, ly = (rnd(seed) - 0.5f) * cam.aperture;
                    dx = dx - lx; dy = dy - ly;
                    dir = normalize(mk(fm(dy, cv.x, dx * cu.x) + cw.x, fm(dy, cv.y, dx * cu.y) + cw.y, fm(dy, cv.z, dx * cu.z) + cw.z));
                    org = mk(fm(ly, cv.x, fm(lx, cu.x, eye.x)), fm(ly, cv.y, fm(lx, cu.y, eye.y)), fm(ly, cv.z, fm(lx, cu.z, eye.z)));
                }
                ++nrays[(size_t)tid];
                const SceneHit h = trace_scene<false>(*s, org, dir, 0.0f, 1e16f, (255u ^ 1u) << 16 /* OptixVisibilityMask(255), optixTriangle.cu:130 */);
                if (!h.hit) {
                    result = result + mk(fm(dir.x, 0.5f, 0.5f), fm(dir.y, 0.5f, 0.5f), fm(dir.z, 0.5f, 0.5f));
                    continue;
                }
                const float* N = normals + 9 * (size_t)h.prim;
                const float b0 = (1.0f - h.b1) - h.b2;
                const f3 n = mk(fm(b0, N[0], fm(h.b2, N[6], h.b1 * N[3])), fm(b0, N[1], fm(h.b2, N[7], h.b1 * N[4])), fm(b0, N[2], fm(h.b2, N[8], h.b1 * N[5])));
                const f3 P = mk(fm(n.x, 0.0001f, fm(h.t, dir.x, org.x)), fm(n.y, 0.0001f, fm(h.t, dir.y, org.y)), fm(n.z, 0.0001f, fm(h.t, dir.z, org.z)));
                uint32_t cs = tea4((uint32_t)ix + (uint32_t)width * (uint32_t)iy, dt);
                const int m = mat_indices[h.prim];
                const f3 color = ld3(materials + 3 * m);
                f3 r = mk(0.f, 0.f, 0.f);
                for (int li = 0; li < nlights; ++li) {
                    const f3 wi = pg_wi(L[li], P, cs);
                    const float nd = dot(n, wi);
                    ++nrays[(size_t)tid];
                    // TERMINATE_ON_FIRST_HIT | CULL_DISABLED_ANYHIT (1<<6)
                    const bool occ = trace_scene<true>(*s, P, wi, 0.01f, 1.0f, 64u).hit;
                    const f3 lumi = pg_lumi(L[li]);
                    const f3 term = mk((color.x * lumi.x) * nd, (color.y * lumi.y) * nd, (color.z * lumi.z) * nd);
                    r = r + ((occ || nd < 0.0f) ? mk(0.f, 0.f, 0.f) : term);
                }
                const Onb onb(n);
                const float u1 = rnd(cs), u2 = rnd(cs);
                float sn, csn;
                det_sincos(6.2831855f * u2, sn, csn);
                const float rr = sqrtf(u1);
                f3 w_in = mk(rr * csn, rr * sn, 0.0f);
                w_in.z = sqrtf(fmaxf(0.0f, fm(-w_in.y, w_in.y, fm(-w_in.x, w_in.x, 1.0f))));
                const f3 out = onb.inverse_transform(w_in);
                ++nrays[(size_t)tid];
                const bool bh = trace_scene<true>(*s, P, out, 0.01f, 1e16f, 0u).hit;
                const f3 amb = mk(0.01f, 0.01f, 0.01f);
                r = r + (bh ? amb : amb * color);
                result = result + r;
            }
            const size_t idx = (size_t)iy * width + ix;
            f3 f = result;
            if (!dirty) f = mk(film[3 * idx] + f.x, film[3 * idx + 1] + f.y, film[3 * idx + 2] + f.z);
            film[3 * idx] = f.x; film[3 * idx + 1] = f.y; film[3 * idx + 2] = f.z;
            if (image) make_color(mk(f.x / (float)dt, f.y / (float)dt, f.z / (float)dt), image + 4 * idx);
        }
    });
    uint64_t tot = 0;
    for (auto v : nrays) tot += v;
    return tot;
}

// Camera::compute_ray / LightVariant::wi in the contract's arithmetic, exposed so the tests can hold them against the reference's own
// host evaluation of the same functions (tests/golden/kat.json "playground_cameras" / "playground_lights")
void orc_playground_ray(const void* camera92, uint32_t ix, uint32_t iy, uint32_t width, uint32_t height, uint32_t* seed, float* org3, float* dir3)
{
    PGCam cam;
    memcpy(&cam, camera92, sizeof cam);
    const f3 cu = ld3(cam.u), cv = ld3(cam.v), cw = ld3(cam.w), eye = ld3(cam.eye);
    float dx = fm(2.0f, (float)ix / (float)width, -1.0f), dy = fm(2.0f, (float)iy / (float)height, -1.0f);
    f3 org, dir;
    if (cam.ortho) {
        dir = normalize(mk(fm(dy, cv.x, dx * cu.x) + cw.x, fm(dy, cv.y, dx * cu.y) + cw.y, fm(dy, cv.z, dx * cu.z) + cw.z));
        org = mk(fm(dy, cv.x, fm(dx, cu.x, eye.x)), fm(dy, cv.y, fm(dx, cu.y, eye.y)), fm(dy, cv.z, fm(dx, cu.z, eye.z)));
    } else {
        const float lx = (rnd(*seed) - 0.5f) * cam.aperture, ly = (rnd(*seed) - 0.5f) * cam.aperture;
        dx = dx - lx; dy = dy - ly;
        dir = normalize(mk(fm(dy, cv.x, dx * cu.x) + cw.x, fm(dy, cv.y, dx * cu.y) + cw.y, fm(dy, cv.z, dx * cu.z) + cw.z));
        org = mk(fm(ly, cv.x, fm(lx, cu.x, eye.x)), fm(ly, cv.y, fm(lx, cu.y, eye.y)), fm(ly, cv.z, fm(lx, cu.z, eye.z)));
    }
    org3[0] = org.x; org3[1] = org.y; org3[2] = org.z; dir3[0] = dir.x; dir3[1] = dir.y; dir3[2] = dir.z;
}
void orc_playground_light(const void* light44, const float* p, uint32_t* seed, float* wi3, float* lumi3)
{
    PGLightO l;
    memcpy(&l, light44, sizeof l);
    const f3 w = pg_wi(l, ld3(p), *seed), m = pg_lumi(l);
    wi3[0] = w.x; wi3[1] = w.y; wi3[2] = w.z; lumi3[0] = m.x; lumi3[1] = m.y; lumi3[2] = m.z;
}

// stand-in scene of optix_raytracer_b200/csrc/playground.cu (pg_scene_kernel), same arithmetic
uint64_t orc_playground_scene(uint32_t rows, uint32_t seed, float* verts, float* normals, int32_t* mats, int threads)
{
    const uint32_t cols = 2 * rows;
    const uint64_t blob_each = 2ull * rows * cols, total = 25ull * blob_each + 800ull;
    if (!verts) return total;
    auto blob_vertex = [&](int b, uint32_t i, uint32_t j, f3& nrm) {
        const int gx = b % 5, gy = b / 5;
        const f3 ctr = mk(0.25f * (float)(gx - 2), 0.1f, 0.25f * (float)(gy - 2));
        const float rad = 0.08f;
        if (i == 0) { nrm = mk(0.f, 1.f, 0.f); return mk(ctr.x, ctr.y + rad, ctr.z); }
        if (i == rows) { nrm = mk(0.f, -1.f, 0.f); return mk(ctr.x, ctr.y - rad, ctr.z); }
        j = j % cols;
        const float theta = (3.14159265358979f * (float)i) / (float)rows;
        const float phi = (6.28318530717959f * (float)j) / (float)cols;
        float st, ct, sp, cp, s1, s2, unused;
        det_sincos(theta, st, ct);
        det_sincos(phi, sp, cp);
        const float ph = 0.37f * (float)((seed + 7u * (uint32_t)b) % 17u);
        det_sincos(fm(3.0f, theta, ph), s1, unused);
        det_sincos(fm(4.0f, phi, ph), s2, unused);
        const float disp = fm(st, (0.15f * s1) * s2, 1.0f);
        const float rr = rad * disp;
        const f3 d = mk((rr * st) * cp, rr * ct, (rr * st) * sp);
        nrm = normalize(d);
        return mk(d.x + ctr.x, d.y + ctr.y, d.z + ctr.z);
    };
    const int chunk = 8192;
    parallel_rows((int)((total + chunk - 1) / chunk), std::max(1, threads), [&](int row, int) {
        const uint64_t b0 = (uint64_t)row * chunk, e0 = std::min<uint64_t>(total, b0 + chunk);
        for (uint64_t t = b0; t < e0; ++t) {
            f3 a, b, c, na, nb, nc;
            int mat;
            if (t < 25 * blob_each) {
                const int bi = (int)(t / blob_each);
                const uint64_t r = t - (uint64_t)bi * blob_each, quad = r >> 1;
                const uint32_t i = (uint32_t)(quad / cols), j = (uint32_t)(quad % cols);
                f3 n00, n01, n10, n11;
                const f3 p00 = blob_vertex(bi, i, j, n00), p01 = blob_vertex(bi, i, j + 1, n01);
                const f3 p10 = blob_vertex(bi, i + 1, j, n10), p11 = blob_vertex(bi, i + 1, j + 1, n11);
                if (r & 1) { a = p00; b = p11; c = p01; na = n00; nb = n11; nc = n01; }
                else { a = p00; b = p10; c = p11; na = n00; nb = n10; nc = n11; }
                mat = bi;
            } else {
                const uint32_t r = (uint32_t)(t - 25 * blob_each);
                const int cell = (int)(r >> 1), i = cell / 20 - 10, j = cell % 20 - 10;
                const float x0 = (float)i * 0.1f, x1 = (float)(i + 1) * 0.1f, z0 = (float)j * 0.1f, z1 = (float)(j + 1) * 0.1f;
                if (r & 1) { a = mk(x1, 0.f, z0); b = mk(x0, 0.f, z1); c = mk(x1, 0.f, z1); }
                else { a = mk(x0, 0.f, z0); b = mk(x0, 0.f, z1); c = mk(x1, 0.f, z0); }
                na = nb = nc = mk(0.f, 1.f, 0.f);
                mat = 26;
            }
            float* v = verts + 9 * t;
            float* n = normals + 9 * t;
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = b.x; v[4] = b.y; v[5] = b.z; v[6] = c.x; v[7] = c.y; v[8] = c.z;
            n[0] = na.x; n[1] = na.y; n[2] = na.z; n[3] = nb.x; n[4] = nb.y; n[5] = nb.z; n[6] = nc.x; n[7] = nc.y; n[8] = nc.z;
            mats[t] = mat;
        }
    });
    return total;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// optixMeshViewer — SDK/cuda/whitted.cu:44-98,139-289 + getLocalGeometry (SDK/cuda/LocalGeometry.h:59-163) for UNTEXTURED
// materials (texture fetches are hardware tex2D in the reference and in the product; they are compared on the GPU against the
// reference programs running on OptiX, tests/test_gpu_optix_parity.py).  alpha_mode 2 = ALPHA_MODE_BLEND: result * alpha plus the
// continuation of the ray from the hit times (1 - alpha), depth < 8 (whitted.cu:266-286); without a texture the any-hit programs
// accept every hit (whitted.cu:100-137).  Same op order as csrc/whitted.cu.
// ------------------------------------------------------------------------------------------
extern "C" {

struct orc_whitted_params {
    uint32_t width, height, subframe_index;
    float eye[3], U[3], V[3], W[3];
    float miss_color[3];
    float base_color[4];
    float metallic, roughness;
    float emissive[3];
    int32_t nlights;
    int32_t alpha_mode;  // 0 OPAQUE, 1 MASK, 2 BLEND
};

// normals: ntri*9 floats or NULL (geometric normal); lights: nlights x 36-byte Light records; accum: width*height*4 floats
// (read when subframe_index > 0, written); frame: width*height*4 bytes or NULL.  Returns rays traced.
uint64_t orc_whitted(void* scene, const orc_whitted_params* p, const float* normals, const void* lights36, float* accum, uint8_t* frame, int y0,
                     int y1, int threads)
{
    Scene* s = (Scene*)scene;
    struct LightRec { int32_t type; float color[3]; float intensity; float position[3]; int32_t falloff; };
    static_assert(sizeof(LightRec) == 36, "Light");
    const LightRec* L = (const LightRec*)lights36;
    const f3 eye = ld3(p->eye), U = ld3(p->U), V = ld3(p->V), W = ld3(p->W);
    const int T = std::max(1, threads);
    std::vector<uint64_t> nrays((size_t)T, 0);
    const float PI = 3.14159265358979323846f;
    parallel_rows(y1 - y0, T, [&](int row, int tid) {
        const uint32_t iy = (uint32_t)(y0 + row);
        for (uint32_t ix = 0; ix < p->width; ++ix) {
            uint32_t seed = tea4(iy * p->width + ix, p->subframe_index);
            float jx = 0.5f, jy = 0.5f;
            if (p->subframe_index != 0) { jx = rnd(seed); jy = rnd(seed); }
            const float dx = fm(2.0f, ((float)ix + jx) / (float)p->width, -1.0f), dy = fm(2.0f, ((float)iy + jy) / (float)p->height, -1.0f);
            const f3 dir = normalize(mk(fm(dy, V.x, dx * U.x) + W.x, fm(dy, V.y, dx * U.y) + W.y, fm(dy, V.z, dx * U.z) + W.z));
            // traceRadiance + __miss__constant_radiance / __closesthit__radiance; payload_depth is what the caller put in the payload
            std::function<f3(float, uint32_t)> radiance = [&](float tmin, uint32_t payload_depth) -> f3 {
            ++nrays[(size_t)tid];
            const SceneHit h = trace_scene<false>(*s, eye, dir, tmin, 1e16f, 16u /* CULL_BACK_FACING_TRIANGLES */);
            f3 result;
            if (!h.hit) {
                result = ld3(p->miss_color);
            } else {
                const uint32_t depth = payload_depth + 1u;
                const Instance* in = s->insts.empty() ? nullptr : &s->insts[h.inst];
                const Geometry& g = s->geoms[in ? in->geom : 0];
                const Tri& tr = g.tris[h.prim];
                const float b1 = h.b1, b2 = h.b2, b0 = (1.0f - b1) - b2;
                auto bary = [&](f3 a, f3 b, f3 c) { return mk(fm(b2, c.x, fm(b1, b.x, b0 * a.x)), fm(b2, c.y, fm(b1, b.y, b0 * a.y)), fm(b2, c.z, fm(b1, b.z, b0 * a.z))); };
                f3 P = bary(tr.v0, tr.v1, tr.v2);
                if (in) P = xform_point(in->m, P);
                f3 Ng = cross(tr.v1 - tr.v0, tr.v2 - tr.v0);
                if (in) Ng = xform_normal(in->inv, Ng);
                Ng = normalize(Ng);
                f3 N = Ng;
                if (normals) {
                    const float* nn = normals + 9 * (size_t)h.prim;
                    N = bary(ld3(nn), ld3(nn + 3), ld3(nn + 6));
                    if (in) N = xform_normal(in->inv, N);
                    N = normalize(N);
                }
                const float bc[4] = {p->base_color[0] * 1.0f, p->base_color[1] * 1.0f, p->base_color[2] * 1.0f, p->base_color[3] * 1.0f};
                const float metallic = p->metallic, roughness = p->roughness;
                const float F0 = 0.04f, km = 1.0f - metallic;
                const f3 diff_color = mk((bc[0] * (1.0f - F0)) * km, (bc[1] * (1.0f - F0)) * km, (bc[2] * (1.0f - F0)) * km);
                const f3 spec_color = mk(fm(metallic, bc[0] - F0, F0), fm(metallic, bc[1] - F0, F0), fm(metallic, bc[2] - F0, F0));
                const float alpha = roughness * roughness;
                result = mk(fm(p->emissive[0], 1.0f, 0.0f), fm(p->emissive[1], 1.0f, 0.0f), fm(p->emissive[2], 1.0f, 0.0f));
                if (dot(N, dir) > 0.0f) N = neg(N);
                const f3 Vv = neg(normalize(dir));
                for (int li = 0; li < p->nlights; ++li) {
                    const LightRec& l = L[li];
                    if (l.type == 0) {
                        if (!(depth < 8u)) continue;  // MAX_TRACE_DEPTH
                        const f3 Lv = mk(l.position[0] - P.x, l.position[1] - P.y, l.position[2] - P.z);
                        const float L_dist = length(Lv);
                        const f3 Ld = mk(Lv.x / L_dist, Lv.y / L_dist, Lv.z / L_dist);
                        const f3 H = normalize(Ld + Vv);
                        const float N_dot_L = dot(N, Ld), N_dot_V = dot(N, Vv), N_dot_H = dot(N, H), V_dot_H = dot(Vv, H);
                        if (N_dot_L > 0.0f && N_dot_V > 0.0f) {
                            const float x1 = 1.0f - V_dot_H, x2 = x1 * x1, x5 = (x2 * x2) * x1;
                            const f3 F = mk(fm(1.0f - spec_color.x, x5, spec_color.x), fm(1.0f - spec_color.y, x5, spec_color.y), fm(1.0f - spec_color.z, x5, spec_color.z));
                            const float a2 = alpha * alpha;
                            const float ggx0 = N_dot_L * sqrtf(fm(N_dot_V * N_dot_V, 1.0f - a2, a2));
                            const float ggx1 = N_dot_V * sqrtf(fm(N_dot_L * N_dot_L, 1.0f - a2, a2));
                            const float G_vis = ((2.0f * N_dot_L) * N_dot_V) / (ggx0 + ggx1);
                            const float xx = fm(N_dot_H * N_dot_H, a2 - 1.0f, 1.0f);
                            const float D = a2 / ((PI * xx) * xx);
                            const f3 diff = mk(((1.0f - F.x) * diff_color.x) / PI, ((1.0f - F.y) * diff_color.y) / PI, ((1.0f - F.z) * diff_color.z) / PI);
                            const f3 spec = mk((F.x * G_vis) * D, (F.y * G_vis) * D, (F.z * G_vis) * D);
                            const float sx = (l.color[0] * l.intensity) * N_dot_L, sy = (l.color[1] * l.intensity) * N_dot_L, sz = (l.color[2] * l.intensity) * N_dot_L;
                            const f3 term = mk(sx * (diff.x + spec.x), sy * (diff.y + spec.y), sz * (diff.z + spec.z));
                            ++nrays[(size_t)tid];
                            if (!trace_scene<true>(*s, P, Ld, 0.001f, L_dist - 0.001f, 0u).hit) result = result + term;
                        }
                    } else if (l.type == 1) {
                        result = result + mk(l.color[0] * bc[0], l.color[1] * bc[1], l.color[2] * bc[2]);
                    }
                }
                if (p->alpha_mode == 2) {
                    result = mk(result.x * bc[3], result.y * bc[3], result.z * bc[3]);
                    if (depth < 8u) {
                        const f3 deeper = radiance(h.t, depth);  // tmin = optixGetRayTmax()
                        const float w = 1.0f - bc[3];
                        result = mk(fm(deeper.x, w, result.x), fm(deeper.y, w, result.y), fm(deeper.z, w, result.z));
                    }
                }
            }
            return result;
            };
            f3 result = radiance(0.0f, 0u);
            const size_t idx = (size_t)iy * p->width + ix;
            if (p->subframe_index > 0) {
                const float a = 1.0f / (float)(p->subframe_index + 1);
                const f3 prev = ld3(accum + 4 * idx);
                result = mk(fm(a, result.x - prev.x, prev.x), fm(a, result.y - prev.y, prev.y), fm(a, result.z - prev.z, prev.z));
            }
            accum[4 * idx] = result.x; accum[4 * idx + 1] = result.y; accum[4 * idx + 2] = result.z; accum[4 * idx + 3] = 1.0f;
            if (frame) make_color(result, frame + 4 * idx);
        }
    });
    uint64_t tot = 0;
    for (auto v : nrays) tot += v;
    return tot;
}

}  // extern "C"
