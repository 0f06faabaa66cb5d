// bvhlab.cpp — TEST / DESIGN INFRASTRUCTURE (CPU only, not part of the product): an offline laboratory for the hierarchy decisions of
// csrc/bvh_build.cu.  It restates the builder's pipeline on the host (Morton keys -> Karras radix tree -> 8-wide collapse with octant
// slots and 8-bit child boxes), offers the alternatives that were considered (top-down binned SAH as the quality yardstick, the
// Ylitie/Karras/Laine 2017 dynamic-programming collapse, different leaf sizes and cost constants), and counts what the GPU's
// instrumented kernel counts — wide nodes visited and triangles tested per path segment — on a Cornell-style path-traced ray set over
// the synthetic scene of BASELINE.json configs[4] (mesh from oracle/liboracle.so: orc_synth_mesh).  The visit counts of a variant
// decide whether it is worth a GPU run; the SAH yardstick is the "oracle-side SAH tree" the roofline's node counts are reported against.
//
//   g++ -O2 -std=c++17 -pthread -o tools/bvhlab tools/bvhlab.cpp -Loracle -loracle -Wl,-rpath,$PWD/oracle
//   tools/bvhlab --tris 2000000 --hier lbvh|sah|ploc --collapse greedy|dp [--cnode 1 --ctri 0.3 --leaf 3] [--res 320x180]
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

extern "C" void orc_synth_mesh(uint64_t total, uint32_t seed, float* verts, uint32_t* mats, int threads);

struct V3 { float x, y, z; };
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline V3 norm(V3 a) { float l = sqrtf(dot(a, a)); return a * (1.0f / l); }
static inline float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
    void grow(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], b.lo[a]); hi[a] = fmaxf(hi[a], b.hi[a]); } }
    void grow(V3 p) { lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z); hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z); }
    float area() const { float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx * dy + dy * dz + dz * dx; }
};

struct Mesh { std::vector<V3> v; uint64_t n; };  // 3 vertices per triangle
static Box tri_box(const Mesh& m, uint32_t t) { Box b; b.reset(); b.grow(m.v[3 * (size_t)t]); b.grow(m.v[3 * (size_t)t + 1]); b.grow(m.v[3 * (size_t)t + 2]); return b; }

// ---- binary hierarchy: internal ids [0, n-1), leaf at sorted position s is (n-1)+s (as bvh_build.cu) -----------------------------------
struct BinTree {
    int n = 0;                         // leaves
    std::vector<int> left, right;      // per internal node
    std::vector<Box> box;              // 2n-1
    std::vector<int> count;            // per internal node: leaves below
    std::vector<uint32_t> leaf_tri;    // sorted position -> triangle
    int root = 0;
};

static uint64_t spread3(uint64_t x)
{
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

static void refit(BinTree& t, int node)
{
    // iterative post-order
    std::vector<int> st{node}, order;
    const int ni = t.n - 1;
    while (!st.empty()) { int x = st.back(); st.pop_back(); if (x >= ni) continue; order.push_back(x); st.push_back(t.left[x]); st.push_back(t.right[x]); }
    for (size_t k = order.size(); k-- > 0;) {
        const int x = order[k];
        Box b = t.box[t.left[x]]; b.grow(t.box[t.right[x]]);
        t.box[x] = b;
        t.count[x] = (t.left[x] >= ni ? 1 : t.count[t.left[x]]) + (t.right[x] >= ni ? 1 : t.count[t.right[x]]);
    }
}

static void sorted_leaves(const Mesh& m, BinTree& t, std::vector<uint64_t>& keys)
{
    const int n = (int)m.n;
    t.n = n;
    Box sb; sb.reset();
    std::vector<Box> tb(n);
    for (int i = 0; i < n; ++i) { tb[i] = tri_box(m, i); sb.grow(tb[i]); }
    const int bits = n < (1 << 14) ? 10 : (n < (1 << 22) ? 16 : (n < (1 << 27) ? 18 : 21));
    const float cells = (float)(1u << bits);
    const uint32_t maxc = (1u << bits) - 1u;
    std::vector<std::pair<uint64_t, uint32_t>> kv(n);
    for (int i = 0; i < n; ++i) {
        uint64_t q[3];
        for (int a = 0; a < 3; ++a) {
            const float c = 0.5f * (tb[i].lo[a] + tb[i].hi[a]);
            const float ext = sb.hi[a] - sb.lo[a];
            float u = ext > 0 ? (c - sb.lo[a]) / ext : 0.f;
            u = fminf(fmaxf(u * cells, 0.f), (float)maxc);
            q[a] = std::min((uint32_t)u, maxc);
        }
        kv[i] = {(spread3(q[0]) << 2) | (spread3(q[1]) << 1) | spread3(q[2]), (uint32_t)i};
    }
    std::sort(kv.begin(), kv.end());
    keys.resize(n);
    t.leaf_tri.resize(n);
    t.box.assign(2 * (size_t)n - 1, Box());
    t.left.assign(n - 1, -1); t.right.assign(n - 1, -1); t.count.assign(n - 1, 0);
    for (int s = 0; s < n; ++s) { keys[s] = kv[s].first; t.leaf_tri[s] = kv[s].second; t.box[n - 1 + s] = tb[kv[s].second]; }
}

static void build_lbvh(const Mesh& m, BinTree& t)
{
    std::vector<uint64_t> keys;
    sorted_leaves(m, t, keys);
    const int n = t.n;
    auto delta = [&](int i, int j) -> int {
        if (j < 0 || j >= n) return -1;
        const uint64_t a = keys[i], b = keys[j];
        if (a == b) return 64 + __builtin_clz((uint32_t)i ^ (uint32_t)j);
        return __builtin_clzll(a ^ b);
    };
    for (int i = 0; i < n - 1; ++i) {  // karras_kernel, verbatim semantics
        const int d = (delta(i, i + 1) - delta(i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = delta(i, i - d);
        int lmax = 2;
        while (delta(i, i + lmax * d) > dmin) lmax <<= 1;
        int l = 0;
        for (int tt = lmax >> 1; tt >= 1; tt >>= 1) if (delta(i, i + (l + tt) * d) > dmin) l += tt;
        const int j = i + l * d;
        const int dnode = delta(i, j);
        int s = 0, tt = l;
        do { tt = (tt + 1) >> 1; if (delta(i, i + (s + tt) * d) > dnode) s += tt; } while (tt > 1);
        const int gamma = i + s * d + std::min(d, 0);
        const int first = std::min(i, j), last = std::max(i, j);
        t.left[i] = (first == gamma) ? (n - 1 + gamma) : gamma;
        t.right[i] = (last == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    }
    t.root = 0;
    refit(t, 0);
}

// top-down binned SAH (16 bins, all three axes): the quality yardstick
static void build_sah(const Mesh& m, BinTree& t)
{
    const int n = (int)m.n;
    t.n = n;
    std::vector<Box> tb(n);
    std::vector<V3> ctr(n);
    for (int i = 0; i < n; ++i) { tb[i] = tri_box(m, i); ctr[i] = {0.5f * (tb[i].lo[0] + tb[i].hi[0]), 0.5f * (tb[i].lo[1] + tb[i].hi[1]), 0.5f * (tb[i].lo[2] + tb[i].hi[2])}; }
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    t.left.assign(n - 1, -1); t.right.assign(n - 1, -1); t.count.assign(n - 1, 0);
    t.box.assign(2 * (size_t)n - 1, Box());
    t.leaf_tri.resize(n);
    std::atomic<int> next_node{1};
    struct Job { int node, first, last; };
    std::function<void(int, int, int, int)> rec = [&](int node, int first, int last, int depth) {
        // node covers idx[first..last], last-first+1 >= 2
        constexpr int NB = 16;
        Box cb; cb.reset();
        for (int i = first; i <= last; ++i) cb.grow(ctr[idx[i]]);
        int best_axis = -1, best_bin = -1;
        float best_cost = INFINITY;
        for (int a = 0; a < 3; ++a) {
            const float ext = cb.hi[a] - cb.lo[a];
            if (!(ext > 0)) continue;
            Box bb[NB]; int bc[NB];
            for (int b = 0; b < NB; ++b) { bb[b].reset(); bc[b] = 0; }
            const float k = NB * (1 - 1e-6f) / ext;
            for (int i = first; i <= last; ++i) { int b = (int)((comp(ctr[idx[i]], a) - cb.lo[a]) * k); b = std::min(std::max(b, 0), NB - 1); bb[b].grow(tb[idx[i]]); bc[b]++; }
            float ra[NB]; Box r; r.reset(); int rc = 0;
            for (int b = NB - 1; b > 0; --b) { r.grow(bb[b]); rc += bc[b]; ra[b] = rc ? r.area() * rc : 0.f; }
            Box l; l.reset(); int lc = 0;
            for (int b = 0; b < NB - 1; ++b) {
                l.grow(bb[b]); lc += bc[b];
                if (lc == 0 || lc == last - first + 1) continue;
                const float c = l.area() * lc + ra[b + 1];
                if (c < best_cost) { best_cost = c; best_axis = a; best_bin = b; }
            }
        }
        int mid;
        if (best_axis < 0) mid = (first + last) / 2;
        else {
            const int a = best_axis;
            const float ext = cb.hi[a] - cb.lo[a];
            const float k = NB * (1 - 1e-6f) / ext;
            auto it = std::partition(idx.begin() + first, idx.begin() + last + 1, [&](uint32_t id) { int b = (int)((comp(ctr[id], a) - cb.lo[a]) * k); b = std::min(std::max(b, 0), NB - 1); return b <= best_bin; });
            mid = (int)(it - idx.begin()) - 1;
            if (mid < first || mid >= last) mid = (first + last) / 2;
        }
        auto child = [&](int f, int l) -> int {
            if (f == l) return n - 1 + f;
            return next_node.fetch_add(1);
        };
        const int lc = child(first, mid), rc = child(mid + 1, last);
        t.left[node] = lc; t.right[node] = rc;
        std::thread th;
        const bool par = depth < 3 && last - first > 100000;
        if (lc < n - 1) { if (par) th = std::thread(rec, lc, first, mid, depth + 1); else rec(lc, first, mid, depth + 1); }
        if (rc < n - 1) rec(rc, mid + 1, last, depth + 1);
        if (th.joinable()) th.join();
    };
    rec(0, 0, n - 1, 0);
    for (int s = 0; s < n; ++s) { t.leaf_tri[s] = idx[s]; t.box[n - 1 + s] = tb[idx[s]]; }
    t.root = 0;
    refit(t, 0);
}

// PLOC over the Morton-ordered leaves (ploc_* kernels of bvh_build.cu), radius R
static void build_ploc(const Mesh& m, BinTree& t, int R)
{
    std::vector<uint64_t> keys;
    sorted_leaves(m, t, keys);
    const int n = t.n;
    std::vector<uint32_t> cl(n), nxt;
    for (int i = 0; i < n; ++i) cl[i] = n - 1 + i;
    uint32_t node_base = 0;
    std::vector<uint32_t> nearest;
    while (cl.size() > 1) {
        const int c = (int)cl.size();
        nearest.assign(c, 0);
        for (int i = 0; i < c; ++i) {
            float best = INFINITY; int bj = -1;
            const Box& a = t.box[cl[i]];
            for (int d = -R; d <= R; ++d) {
                const int j = i + d;
                if (d == 0 || j < 0 || j >= c) continue;
                Box u = a; u.grow(t.box[cl[j]]);
                const float ar = u.area();
                if (ar < best) { best = ar; bj = j; }
            }
            nearest[i] = bj;
        }
        nxt.clear();
        for (int i = 0; i < c; ++i) {
            const int j = nearest[i];
            const bool mutual = nearest[j] == (uint32_t)i;
            if (mutual && i > j) continue;
            if (!mutual) { nxt.push_back(cl[i]); continue; }
            const uint32_t id = node_base++;
            t.left[id] = cl[i]; t.right[id] = cl[j];
            Box b = t.box[cl[i]]; b.grow(t.box[cl[j]]);
            t.box[id] = b;
            t.count[id] = (cl[i] >= (uint32_t)(n - 1) ? 1 : t.count[cl[i]]) + (cl[j] >= (uint32_t)(n - 1) ? 1 : t.count[cl[j]]);
            nxt.push_back(id);
        }
        cl.swap(nxt);
    }
    t.root = cl[0];
}


// Bottom-up tree rotations (Kensler 2008) as a SAH refit of a finished binary hierarchy: in post-order, every internal node considers
// exchanging one child with a grandchild on the other side (four candidates) and takes the one that shrinks the changed child's surface
// area the most.  The thread that finishes a node owns the whole subtree below it, so the GPU form is the refit kernel's second-arrival
// walk with this step added (no further synchronisation).
static int rotate_pass(BinTree& t)
{
    const int ni = t.n - 1;
    std::vector<int> st{t.root}, order;
    while (!st.empty()) { int x = st.back(); st.pop_back(); if (x >= ni) continue; order.push_back(x); st.push_back(t.left[x]); st.push_back(t.right[x]); }
    auto cnt = [&](int id) { return id >= ni ? 1 : t.count[id]; };
    int done = 0;
    for (size_t k = order.size(); k-- > 0;) {
        const int x = order[k];
        int L = t.left[x], R = t.right[x];
        float best = 0.f; int which = -1;
        auto unite = [&](int a, int b) { Box u = t.box[a]; u.grow(t.box[b]); return u.area(); };
        if (L < ni) {
            const float aL = t.box[L].area();
            const float c0 = unite(R, t.right[L]) - aL;   // R <-> LL : L' = (R, LR)
            const float c1 = unite(t.left[L], R) - aL;    // R <-> LR : L' = (LL, R)
            if (c0 < best) { best = c0; which = 0; }
            if (c1 < best) { best = c1; which = 1; }
        }
        if (R < ni) {
            const float aR = t.box[R].area();
            const float c2 = unite(L, t.right[R]) - aR;   // L <-> RL : R' = (L, RR)
            const float c3 = unite(t.left[R], L) - aR;    // L <-> RR : R' = (RL, L)
            if (c2 < best) { best = c2; which = 2; }
            if (c3 < best) { best = c3; which = 3; }
        }
        if (which >= 0) {
            ++done;
            if (which == 0) { const int g = t.left[L]; t.left[L] = R; t.right[x] = g; }
            else if (which == 1) { const int g = t.right[L]; t.right[L] = R; t.right[x] = g; }
            else if (which == 2) { const int g = t.left[R]; t.left[R] = L; t.left[x] = g; }
            else { const int g = t.right[R]; t.right[R] = L; t.left[x] = g; }
            const int c = which < 2 ? L : R;
            Box b = t.box[t.left[c]]; b.grow(t.box[t.right[c]]);
            t.box[c] = b; t.count[c] = cnt(t.left[c]) + cnt(t.right[c]);
        }
        Box b = t.box[t.left[x]]; b.grow(t.box[t.right[x]]);
        t.box[x] = b; t.count[x] = cnt(t.left[x]) + cnt(t.right[x]);
    }
    return done;
}

// ---- 8-wide collapse ----------------------------------------------------------------------------------------------------------------
struct WNode {
    int nchild = 0;
    int slot_child[8];      // -1 empty; >= 0 wide node index; <= -2: leaf group (-2 - group index)
    float lo[8][3], hi[8][3];  // decoded (quantised) child boxes
};
struct LeafGroup { uint32_t tri[3]; int cnt; };
struct Wide {
    std::vector<WNode> nodes;
    std::vector<LeafGroup> groups;
    int depth = 0;
};

struct CollapseOpt { std::string mode = "greedy"; float cnode = 1.0f, ctri = 0.3f; int leaf = 3; bool quant = true; };

static void leaves_of(const BinTree& t, int node, std::vector<uint32_t>& out)
{
    const int ni = t.n - 1;
    std::vector<int> st{node};
    while (!st.empty()) { int x = st.back(); st.pop_back(); if (x >= ni) { out.push_back(t.leaf_tri[x - ni]); continue; } st.push_back(t.right[x]); st.push_back(t.left[x]); }
}

static void emit_node(const BinTree& t, const std::vector<int>& ids, const std::vector<char>& inner, Wide& w, int self, std::vector<std::pair<int, int>>& next /* (binary root, wide index) */,
                      const CollapseOpt& o)
{
    const int m = (int)ids.size();
    Box all; all.reset();
    for (int id : ids) all.grow(t.box[id]);
    const float maxext = fmaxf(fmaxf(all.hi[0] - all.lo[0], all.hi[1] - all.lo[1]), all.hi[2] - all.lo[2]);
    const float pad = fmaxf(maxext * 1.52587890625e-5f, 1e-30f);
    float P[3], scale[3];
    for (int a = 0; a < 3; ++a) {
        P[a] = all.lo[a] - pad;
        const float H = all.hi[a] + pad;
        int e; frexpf((H - P[a]) / 255.0f, &e);
        float s = ldexpf(1.0f, e);  // >= ext/255
        while (s * 0.5f * 255.0f >= (H - P[a]) && s > 1e-37f) s *= 0.5f;
        scale[a] = s;
    }
    const float ctr[3] = {0.5f * (all.lo[0] + all.hi[0]), 0.5f * (all.lo[1] + all.hi[1]), 0.5f * (all.lo[2] + all.hi[2])};
    int child_in_slot[8];
    for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
    for (int k = 0; k < m; ++k) {
        const Box& b = t.box[ids[k]];
        const float ox = 0.5f * (b.lo[0] + b.hi[0]) - ctr[0], oy = 0.5f * (b.lo[1] + b.hi[1]) - ctr[1], oz = 0.5f * (b.lo[2] + b.hi[2]) - ctr[2];
        int bs = -1; float bc = -INFINITY;
        for (int s = 0; s < 8; ++s) {
            if (child_in_slot[s] >= 0) continue;
            const float c = ((s & 4) ? ox : -ox) + ((s & 2) ? oy : -oy) + ((s & 1) ? oz : -oz);
            if (c > bc) { bc = c; bs = s; }
        }
        child_in_slot[bs] = k;
    }
    WNode& nd = w.nodes[self];
    nd.nchild = m;
    for (int s = 0; s < 8; ++s) {
        nd.slot_child[s] = -1;
        const int k = child_in_slot[s];
        if (k < 0) continue;
        const Box& b = t.box[ids[k]];
        for (int a = 0; a < 3; ++a) {
            if (o.quant) {
                const float ql = floorf((b.lo[a] - pad - P[a]) / scale[a]), qh = ceilf((b.hi[a] + pad - P[a]) / scale[a]);
                nd.lo[s][a] = P[a] + fminf(fmaxf(ql, 0.f), 255.f) * scale[a];
                nd.hi[s][a] = P[a] + fminf(fmaxf(qh, 0.f), 255.f) * scale[a];
            } else { nd.lo[s][a] = b.lo[a] - pad; nd.hi[s][a] = b.hi[a] + pad; }
        }
        if (inner[k]) {
            const int wi = (int)w.nodes.size() + (int)next.size();
            nd.slot_child[s] = wi;
            next.push_back({ids[k], wi});
        } else {
            std::vector<uint32_t> tr;
            leaves_of(t, ids[k], tr);
            LeafGroup g; g.cnt = (int)tr.size();
            for (int j = 0; j < g.cnt && j < 3; ++j) g.tri[j] = tr[j];
            nd.slot_child[s] = -2 - (int)w.groups.size();
            w.groups.push_back(g);
        }
    }
}

static void collapse(const BinTree& t, Wide& w, const CollapseOpt& o)
{
    const int ni = t.n - 1;
    auto cnt = [&](int id) { return id >= ni ? 1 : t.count[id]; };
    // DP tables (Ylitie, Karras, Laine 2017, sec. 3.1): C[n][i-1], i = 1..7: cheapest forest of at most i wide nodes / leaves for subtree n
    std::vector<float> C;
    const float root_area = t.box[t.root].area();
    auto A = [&](int id) { return t.box[id].area() / root_area; };
    auto Cget = [&](int id, int i) -> float { return id >= ni ? A(id) * o.ctri : C[(size_t)id * 7 + (i - 1)]; };
    auto distribute = [&](int n, int j, int* kbest) -> float {
        float best = INFINITY; int kb = 1;
        for (int k = 1; k < j; ++k) {
            const float c = Cget(t.left[n], std::min(k, 7)) + Cget(t.right[n], std::min(j - k, 7));
            if (c < best) { best = c; kb = k; }
        }
        if (kbest) *kbest = kb;
        return best;
    };
    if (o.mode == "dp") {
        C.assign((size_t)ni * 7, 0.f);
        std::vector<int> st{t.root}, order;
        while (!st.empty()) { int x = st.back(); st.pop_back(); if (x >= ni) continue; order.push_back(x); st.push_back(t.left[x]); st.push_back(t.right[x]); }
        for (size_t q = order.size(); q-- > 0;) {
            const int n = order[q];
            const float cleaf = cnt(n) <= o.leaf ? A(n) * cnt(n) * o.ctri : INFINITY;
            const float cint = distribute(n, 8, nullptr) + A(n) * o.cnode;
            C[(size_t)n * 7] = fminf(cleaf, cint);
            for (int i = 2; i <= 7; ++i) C[(size_t)n * 7 + i - 1] = fminf(distribute(n, i, nullptr), C[(size_t)n * 7 + i - 2]);
        }
    }
    w.nodes.clear(); w.groups.clear();
    w.nodes.push_back(WNode());
    std::vector<std::pair<int, int>> level{{t.root, 0}}, next;
    w.depth = 0;
    while (!level.empty()) {
        ++w.depth;
        next.clear();
        const size_t base = w.nodes.size();
        std::vector<std::vector<int>> all_ids(level.size());
        std::vector<std::vector<char>> all_inner(level.size());
        for (size_t q = 0; q < level.size(); ++q) {
            const int root = level[q].first;
            std::vector<int>& ids = all_ids[q];
            std::vector<char>& inner = all_inner[q];
            if (root >= ni) { ids.push_back(root); inner.push_back(0); continue; }
            if (o.mode == "greedy") {
                ids = {t.left[root], t.right[root]};
                for (int phase = 0; phase < 2; ++phase) {
                    while ((int)ids.size() < 8) {
                        int best = -1; float ba = -1.f;
                        for (int k = 0; k < (int)ids.size(); ++k) {
                            const bool in = ids[k] < ni;
                            const bool ok = phase == 0 ? (in && cnt(ids[k]) > o.leaf) : in;
                            const float a = t.box[ids[k]].area();
                            if (ok && a > ba) { ba = a; best = k; }
                        }
                        if (best < 0) break;
                        const int id = ids[best];
                        ids[best] = t.left[id];
                        ids.push_back(t.right[id]);
                    }
                }
                for (int id : ids) inner.push_back(id < ni && cnt(id) > o.leaf);
            } else {
                // reconstruct the DP's choice: children of `root` = distribute(root, 8)
                std::function<void(int, int)> place = [&](int n, int i) {
                    // subtree n gets a budget of i slots
                    if (n >= ni) { ids.push_back(n); inner.push_back(0); return; }
                    while (i > 1 && Cget(n, i) == Cget(n, i - 1)) --i;
                    if (i == 1) {
                        const float cleaf = cnt(n) <= o.leaf ? A(n) * cnt(n) * o.ctri : INFINITY;
                        const float cint = distribute(n, 8, nullptr) + A(n) * o.cnode;
                        ids.push_back(n); inner.push_back(cint < cleaf ? 1 : 0);
                        return;
                    }
                    int k; distribute(n, i, &k);
                    place(t.left[n], k); place(t.right[n], i - k);
                };
                int k; distribute(root, 8, &k);
                place(t.left[root], k); place(t.right[root], 8 - k);
            }
        }
        // emit (two passes so that child indices are level-contiguous like the GPU's)
        std::vector<std::pair<int, int>> nx;
        for (size_t q = 0; q < level.size(); ++q) {
            std::vector<std::pair<int, int>> mine;
            // emit_node numbers children as nodes.size() + next.size(): keep `next` global
            emit_node(t, all_ids[q], all_inner[q], w, level[q].second, nx, o);
            (void)mine;
            // allocate the nodes of the next level lazily below
            (void)base;
        }
        // nx holds (binary root, wide index) with indices assigned as w.nodes.size() + position
        w.nodes.resize(w.nodes.size() + nx.size());
        level.swap(nx);
    }
}

// ---- traversal with the GPU's visiting order ------------------------------------------------------------------------------------------
struct Hit { float t; uint32_t tri; };
struct Counters { uint64_t nodes = 0, tris = 0, rays = 0; };

static bool tri_isect(const Mesh& m, uint32_t ti, V3 o, V3 d, float tmin, float tmax, float& t)
{
    const V3 v0 = m.v[3 * (size_t)ti], e1 = m.v[3 * (size_t)ti + 1] - v0, e2 = m.v[3 * (size_t)ti + 2] - v0;
    const V3 p = cross(d, e2);
    const float det = dot(e1, p);
    if (det == 0.f) return false;
    const float inv = 1.0f / det;
    const V3 s = o - v0;
    const float u = dot(s, p) * inv;
    if (u < 0.f || u > 1.f) return false;
    const V3 q = cross(s, e1);
    const float v = dot(d, q) * inv;
    if (v < 0.f || u + v > 1.f) return false;
    t = dot(e2, q) * inv;
    return t > tmin && t < tmax;
}

static bool trace(const Mesh& m, const Wide& w, V3 o, V3 d, float tmin, float tmax, bool any, Hit& hit, Counters& c)
{
    const float idx = 1.0f / (fabsf(d.x) < 1e-24f ? copysignf(1e-24f, d.x) : d.x), idy = 1.0f / (fabsf(d.y) < 1e-24f ? copysignf(1e-24f, d.y) : d.y),
                idz = 1.0f / (fabsf(d.z) < 1e-24f ? copysignf(1e-24f, d.z) : d.z);
    const int oct = (d.x < 0 ? 4 : 0) | (d.y < 0 ? 2 : 0) | (d.z < 0 ? 1 : 0), octinv = 7 - oct;
    int stack[256]; int sp = 0;
    stack[sp++] = 0;
    bool found = false;
    float tfar = tmax;
    c.rays++;
    while (sp > 0) {
        const int ni = stack[--sp];
        const WNode& nd = w.nodes[ni];
        c.nodes++;
        int inner[8], ninner = 0;
        // slots in ascending priority so that pushing leaves the highest priority on top
        for (int pr = 0; pr < 8; ++pr) {
            const int s = pr ^ octinv;
            const int ch = nd.slot_child[s];
            if (ch == -1) continue;
            const float x0 = (nd.lo[s][0] - o.x) * idx, x1 = (nd.hi[s][0] - o.x) * idx, y0 = (nd.lo[s][1] - o.y) * idy, y1 = (nd.hi[s][1] - o.y) * idy,
                        z0 = (nd.lo[s][2] - o.z) * idz, z1 = (nd.hi[s][2] - o.z) * idz;
            const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
            const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tfar));
            if (!(tn <= tf * 1.0000038f)) continue;
            if (ch >= 0) inner[ninner++] = ch;
            else {
                const LeafGroup& g = w.groups[-2 - ch];
                for (int j = 0; j < g.cnt; ++j) {
                    c.tris++;
                    float t;
                    if (tri_isect(m, g.tri[j], o, d, tmin, tfar, t)) { tfar = t; hit.t = t; hit.tri = g.tri[j]; found = true; if (any) return true; }
                }
            }
        }
        for (int k = 0; k < ninner; ++k) stack[sp++] = inner[k];
    }
    return found;
}

// ---- a Cornell-style path-traced ray set (depth cap 3 + next-event estimation, like optixMultiGPU.cu) ---------------------------------------
static inline uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s & 0x00ffffffu; }
static inline float rnd(uint32_t& s) { return (float)lcg(s) / 16777216.0f; }

int main(int argc, char** argv)
{
    uint64_t T = 2000000;
    std::string hier = "lbvh";
    CollapseOpt co;
    int rotate = 0;
    int W = 320, H = 180, ploc_r = 16, threads = (int)std::thread::hardware_concurrency();
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return std::string(argv[++i]); };
        if (a == "--tris") T = strtoull(next().c_str(), 0, 10);
        else if (a == "--hier") hier = next();
        else if (a == "--collapse") co.mode = next();
        else if (a == "--cnode") co.cnode = (float)atof(next().c_str());
        else if (a == "--ctri") co.ctri = (float)atof(next().c_str());
        else if (a == "--leaf") co.leaf = atoi(next().c_str());
        else if (a == "--noquant") co.quant = false;
        else if (a == "--rotate") rotate = atoi(next().c_str());
        else if (a == "--ploc-radius") ploc_r = atoi(next().c_str());
        else if (a == "--res") { std::string r = next(); sscanf(r.c_str(), "%dx%d", &W, &H); }
        else if (a == "--threads") threads = atoi(next().c_str());
    }
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    Mesh m; m.n = T; m.v.resize(3 * T);
    std::vector<uint32_t> mats(T);
    auto t0 = now();
    orc_synth_mesh(T, 0, (float*)m.v.data(), mats.data(), threads);
    BinTree bt;
    auto t1 = now();
    if (hier == "lbvh") build_lbvh(m, bt);
    else if (hier == "sah") build_sah(m, bt);
    else if (hier == "ploc") build_ploc(m, bt, ploc_r);
    else { fprintf(stderr, "unknown --hier\n"); return 1; }
    for (int r = 0; r < rotate; ++r) { const int d = rotate_pass(bt); fprintf(stderr, "rotate pass %d: %d rotations\n", r, d); }
    auto t2 = now();
    Wide w;
    collapse(bt, w, co);
    auto t3 = now();
    // statistics of the wide tree
    uint64_t children = 0, inner_children = 0, leaf_tris = 0; uint64_t fill[9] = {0};
    double sah_nodes = 0, sah_tris = 0;
    const float ra = bt.box[bt.root].area();
    for (const WNode& nd : w.nodes) {
        children += nd.nchild; fill[nd.nchild]++;
        for (int s = 0; s < 8; ++s) {
            if (nd.slot_child[s] == -1) continue;
            const float dx = nd.hi[s][0] - nd.lo[s][0], dy = nd.hi[s][1] - nd.lo[s][1], dz = nd.hi[s][2] - nd.lo[s][2];
            const double a = (dx * dy + dy * dz + dz * dx) / ra;
            if (nd.slot_child[s] >= 0) { inner_children++; sah_nodes += a; }
            else { const int c = w.groups[-2 - nd.slot_child[s]].cnt; leaf_tris += c; sah_tris += a * c; }
        }
    }
    printf("{\"triangles\": %llu, \"hier\": \"%s\", \"collapse\": \"%s\", \"cnode\": %g, \"ctri\": %g, \"leaf\": %d, \"quant\": %d,\n", (unsigned long long)T, hier.c_str(),
           co.mode.c_str(), co.cnode, co.ctri, co.leaf, (int)co.quant);
    printf(" \"wide_nodes\": %zu, \"leaf_groups\": %zu, \"depth\": %d, \"children_per_node\": %.3f, \"tris_per_group\": %.3f, \"leaf_tris\": %llu,\n", w.nodes.size(),
           w.groups.size(), w.depth, (double)children / w.nodes.size(), (double)leaf_tris / std::max<size_t>(w.groups.size(), 1), (unsigned long long)leaf_tris);
    printf(" \"fill_histogram\": [%llu, %llu, %llu, %llu, %llu, %llu, %llu, %llu, %llu],\n", (unsigned long long)fill[0], (unsigned long long)fill[1], (unsigned long long)fill[2],
           (unsigned long long)fill[3], (unsigned long long)fill[4], (unsigned long long)fill[5], (unsigned long long)fill[6], (unsigned long long)fill[7], (unsigned long long)fill[8]);
    printf(" \"sah_node_term\": %.3f, \"sah_tri_term\": %.3f, \"seconds\": {\"mesh\": %.2f, \"hierarchy\": %.2f, \"collapse\": %.2f},\n", 1.0 + sah_nodes, sah_tris,
           secs(t0, t1), secs(t1, t2), secs(t2, t3));

    // ray set
    const V3 eye{278.f, 273.f, -900.f}, lookat{278.f, 273.f, 330.f}, up{0.f, 1.f, 0.f};
    const float fov = 35.f, aspect = (float)W / H;
    V3 Wv = lookat - eye; const float wlen = sqrtf(dot(Wv, Wv));
    V3 U = norm(cross(Wv, up)), V = norm(cross(U, Wv));
    const float vlen = wlen * tanf(0.5f * fov * 3.14159265f / 180.f);
    V = V * vlen; U = U * (vlen * aspect);
    const V3 lc{343.f, 548.6f, 227.f}, lv1{-130.f, 0.f, 0.f}, lv2{0.f, 0.f, 105.f};
    std::vector<Counters> rad(threads), shd(threads), cam(threads);
    std::vector<std::thread> pool;
    std::atomic<int> row{0};
    for (int th = 0; th < threads; ++th)
        pool.emplace_back([&, th] {
            for (;;) {
                const int y = row.fetch_add(1);
                if (y >= H) break;
                for (int x = 0; x < W; ++x) {
                    uint32_t seed = (uint32_t)(y * W + x) * 9781u + 12345u;
                    for (int k = 0; k < 4; ++k) lcg(seed);
                    const float dx = 2.f * ((x + rnd(seed)) / W) - 1.f, dy = 2.f * ((y + rnd(seed)) / H) - 1.f;
                    V3 o = eye, d = norm(U * dx + V * dy + Wv);
                    for (int depth = 0; depth <= 3; ++depth) {
                        Hit h{0, 0};
                        Counters& cc = depth == 0 ? cam[th] : rad[th];
                        if (!trace(m, w, o, d, 0.01f, 1e16f, false, h, cc)) break;
                        const V3 P = o + d * h.t;
                        const V3 v0 = m.v[3 * (size_t)h.tri], v1 = m.v[3 * (size_t)h.tri + 1], v2 = m.v[3 * (size_t)h.tri + 2];
                        V3 N = norm(cross(v1 - v0, v2 - v0));
                        if (dot(N, d) > 0) N = N * -1.f;
                        // shadow ray to the light
                        const float l1 = rnd(seed), l2 = rnd(seed);
                        const V3 lp = lc + lv1 * l1 + lv2 * l2;
                        V3 Ld = lp - P; const float dist = sqrtf(dot(Ld, Ld)); Ld = Ld * (1.f / dist);
                        if (dot(N, Ld) > 0 && Ld.y > 0) { Hit hs; trace(m, w, P, Ld, 0.01f, dist - 0.01f, true, hs, shd[th]); }
                        // cosine bounce
                        const float z1 = rnd(seed), z2 = rnd(seed), r = sqrtf(z1), phi = 6.2831853f * z2;
                        const float lx = r * cosf(phi), ly = r * sinf(phi), lz = sqrtf(fmaxf(0.f, 1.f - lx * lx - ly * ly));
                        V3 bn = fabsf(N.x) > fabsf(N.z) ? V3{-N.y, N.x, 0.f} : V3{0.f, -N.z, N.y};
                        bn = norm(bn);
                        const V3 tg = cross(bn, N);
                        o = P; d = norm(tg * lx + bn * ly + N * lz);
                    }
                }
            }
        });
    for (auto& p : pool) p.join();
    auto t4 = now();
    Counters R, S, Cm;
    for (int th = 0; th < threads; ++th) { R.nodes += rad[th].nodes; R.tris += rad[th].tris; R.rays += rad[th].rays; S.nodes += shd[th].nodes; S.tris += shd[th].tris; S.rays += shd[th].rays;
                                           Cm.nodes += cam[th].nodes; Cm.tris += cam[th].tris; Cm.rays += cam[th].rays; }
    const uint64_t rays = R.rays + S.rays + Cm.rays;
    printf(" \"rays\": {\"camera\": %llu, \"bounce\": %llu, \"shadow\": %llu},\n", (unsigned long long)Cm.rays, (unsigned long long)R.rays, (unsigned long long)S.rays);
    printf(" \"nodes_per_segment\": %.3f, \"tris_per_segment\": %.3f,\n", (double)(R.nodes + S.nodes + Cm.nodes) / rays, (double)(R.tris + S.tris + Cm.tris) / rays);
    printf(" \"nodes_per_ray\": {\"camera\": %.3f, \"bounce\": %.3f, \"shadow\": %.3f}, \"tris_per_ray\": {\"camera\": %.3f, \"bounce\": %.3f, \"shadow\": %.3f},\n",
           (double)Cm.nodes / std::max<uint64_t>(Cm.rays, 1), (double)R.nodes / std::max<uint64_t>(R.rays, 1), (double)S.nodes / std::max<uint64_t>(S.rays, 1),
           (double)Cm.tris / std::max<uint64_t>(Cm.rays, 1), (double)R.tris / std::max<uint64_t>(R.rays, 1), (double)S.tris / std::max<uint64_t>(S.rays, 1));
    printf(" \"trace_seconds\": %.2f}\n", secs(t3, t4));
    return 0;
}
