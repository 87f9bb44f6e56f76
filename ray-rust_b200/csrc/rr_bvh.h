// rr_bvh.h — host-side build of the exact culling structure (BVH over the spheres) that rr_trace.cuh traverses.
// Pure C++ (float4 / make_float4 only), shared by the library (rr_ffi.cu) and by the CPU build of the kernel logic in
// tests/hostsim, so that the no-GPU test-suite exercises the same builder + traversal pair as the device.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rr_device.cuh"

namespace rr {

// ---- BVH over the spheres (host build, depth-first layout with escape indices) --------------------
struct Bvh {
    std::vector<float4> a, b;  // node arrays, see DevScene
    std::vector<float4> w;     // wide inner nodes for the ordered traversal (4 x float4 each)
    std::vector<int> order;    // sphere list index in leaf order
    int depth = 0;
    float lo[3], hi[3], r_min;
};

// Split rule: surface-area heuristic over a full sweep of the centre-sorted order on each axis (cost = A_l n_l + A_r n_r),
// for subtrees of <= RR_BVH_SAH_MAX spheres and while the tree is shallow; plain median split otherwise (bounded build
// time and depth). Any partition is valid: the traversal is exact for every tree (rr_trace.cuh).
constexpr int RR_BVH_SAH_MAX = 1 << 16;
constexpr int RR_BVH_SAH_DEPTH = 20;
static const bool g_bvh_sah = [] { const char *e = getenv("RR_BVH_SAH"); return e ? atoi(e) != 0 : true; }();

inline void build_node(const std::vector<float4> &sph, std::vector<int> &idx, int begin, int end, Bvh &out, int depth = 0) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = begin; i < end; ++i) {
        const float4 &s = sph[idx[i]];
        const float c[3] = {s.x, s.y, s.z};
        const float r = std::fabs(s.w);
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::fmin(lo[k], c[k] - r); hi[k] = std::fmax(hi[k], c[k] + r);
            clo[k] = std::fmin(clo[k], c[k]); chi[k] = std::fmax(chi[k], c[k]);
        }
    }
    const size_t me = out.a.size();
    out.a.push_back(make_float4(lo[0], lo[1], lo[2], 0.0f));
    out.b.push_back(make_float4(hi[0], hi[1], hi[2], 0.0f));
    const int count = end - begin;
    int leaf = -1;
    if (count <= RR_BVH_LEAF) {
        leaf = ((int)out.order.size() << 3) | count;
        for (int i = begin; i < end; ++i) out.order.push_back(idx[i]);
    } else {
        auto centre = [&](int p, int axis) { return axis == 0 ? sph[p].x : axis == 1 ? sph[p].y : sph[p].z; };
        int axis = 0, mid = begin + count / 2;
        for (int k = 1; k < 3; ++k) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
        bool sah = g_bvh_sah && count <= RR_BVH_SAH_MAX && depth < RR_BVH_SAH_DEPTH;
        if (sah) {
            double best = INFINITY;
            int best_axis = -1, best_split = 0;
            std::vector<int> tmp(idx.begin() + begin, idx.begin() + end);
            std::vector<double> right_area(count + 1);
            for (int k = 0; k < 3; ++k) {
                std::sort(tmp.begin(), tmp.end(), [&](int p, int q) { const float cp = centre(p, k), cq = centre(q, k); return cp < cq || (cp == cq && p < q); });
                auto grow = [&](double *l, double *h, int p) {
                    const double c[3] = {sph[p].x, sph[p].y, sph[p].z}, r = std::fabs(sph[p].w);
                    for (int a = 0; a < 3; ++a) { l[a] = std::min(l[a], c[a] - r); h[a] = std::max(h[a], c[a] + r); }
                };
                auto area = [](const double *l, const double *h) {
                    const double x = h[0] - l[0], y = h[1] - l[1], z = h[2] - l[2];
                    return x * y + y * z + z * x;
                };
                double l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
                for (int i = count - 1; i > 0; --i) { grow(l, h, tmp[i]); right_area[i] = area(l, h); }
                for (int a = 0; a < 3; ++a) { l[a] = INFINITY; h[a] = -INFINITY; }
                for (int i = 1; i < count; ++i) {  // split: [0, i) | [i, count)
                    grow(l, h, tmp[i - 1]);
                    const double cost = area(l, h) * i + right_area[i] * (count - i);
                    if (cost < best) { best = cost; best_axis = k; best_split = i; }
                }
            }
            if (best_axis >= 0) { axis = best_axis; mid = begin + best_split; } else sah = false;
        }
        auto less = [&](int p, int q) { const float cp = centre(p, axis), cq = centre(q, axis); return cp < cq || (cp == cq && p < q); };
        std::nth_element(idx.begin() + begin, idx.begin() + mid, idx.begin() + end, less);
        build_node(sph, idx, begin, mid, out, depth + 1);
        build_node(sph, idx, mid, end, out, depth + 1);
    }
    const int escape = (int)out.a.size();  // first node after this subtree
    std::memcpy(&out.a[me].w, &escape, sizeof(int));
    std::memcpy(&out.b[me].w, &leaf, sizeof(int));
}

// sph_m: (cx, cy, cz, r). Returns false when no BVH should be used (few spheres, non-finite data).
inline bool build_bvh(const std::vector<float4> &sph_m, Bvh &out) {
    const int n = (int)sph_m.size();
    if (n < RR_BVH_MIN_SPHERES || n >= (1 << 24)) return false;  // (inner references are byte offsets in an int)
    float rmin = INFINITY;
    for (const float4 &s : sph_m) {
        if (!std::isfinite(s.x) || !std::isfinite(s.y) || !std::isfinite(s.z) || !std::isfinite(s.w)) return false;
        // (coordinates beyond 5e7: the slab arithmetic of the traversal could overflow, see raycast() in rr_trace.cuh)
        if (std::fabs(s.x) + std::fabs(s.w) > 5e7f || std::fabs(s.y) + std::fabs(s.w) > 5e7f || std::fabs(s.z) + std::fabs(s.w) > 5e7f) return false;
        rmin = std::fmin(rmin, std::fabs(s.w));
    }
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    build_node(sph_m, idx, 0, n, out);
    for (int k = 0; k < 3; ++k) {
        out.lo[k] = k == 0 ? out.a[0].x : k == 1 ? out.a[0].y : out.a[0].z;
        out.hi[k] = k == 0 ? out.b[0].x : k == 1 ? out.b[0].y : out.b[0].z;
    }
    out.r_min = rmin;
    // Wide nodes for the ordered (stack) traversal: one record per INNER node holding both child boxes and the
    // child references (>= 0: inner record index, < 0: ~leaf code). In the depth-first arrays the left child of
    // node i is i + 1 and the right child is the escape index of the left child.
    const int nn = (int)out.a.size();
    auto as_int = [](float f) { int i; std::memcpy(&i, &f, sizeof i); return i; };
    auto as_float = [](int i) { float f; std::memcpy(&f, &i, sizeof f); return f; };
    std::vector<int> inner_id(nn, -1);
    int n_inner = 0;
    for (int i = 0; i < nn; ++i) if (as_int(out.b[i].w) < 0) inner_id[i] = n_inner++;
    if (n_inner == 0) return false;  // a single leaf: the brute-force scan is the better kernel
    out.w.resize((size_t)4 * n_inner);
    // inner reference = BYTE offset of the 64-byte record (the traversal adds it to the array base), leaf = ~code
    auto ref = [&](int node) { return inner_id[node] >= 0 ? 64 * inner_id[node] : ~as_int(out.b[node].w); };
    for (int i = 0; i < nn; ++i) {
        if (inner_id[i] < 0) continue;
        const int l = i + 1, r = as_int(out.a[l].w);
        float4 *q = &out.w[(size_t)4 * inner_id[i]];
        // centre / half extent of both child boxes, (left, right) side by side per component: the traversal forms the slab
        // distances of both children with one packed fma per term (rr_trace.cuh). h is rounded UP until [c - h, c + h]
        // contains the box exactly.
        float c[2][3], h[2][3];
        const int ch[2] = {l, r};
        for (int k = 0; k < 2; ++k) {
            const float lo3[3] = {out.a[ch[k]].x, out.a[ch[k]].y, out.a[ch[k]].z}, hi3[3] = {out.b[ch[k]].x, out.b[ch[k]].y, out.b[ch[k]].z};
            for (int a = 0; a < 3; ++a) {
                c[k][a] = (float)(((double)lo3[a] + (double)hi3[a]) * 0.5);
                h[k][a] = (float)(((double)hi3[a] - (double)lo3[a]) * 0.5);
                while ((double)c[k][a] - (double)h[k][a] > (double)lo3[a] || (double)c[k][a] + (double)h[k][a] < (double)hi3[a])
                    h[k][a] = std::nextafter(h[k][a], INFINITY);
            }
        }
        q[0] = make_float4(c[0][0], c[1][0], c[0][1], c[1][1]);
        q[1] = make_float4(c[0][2], c[1][2], h[0][0], h[1][0]);
        q[2] = make_float4(h[0][1], h[1][1], h[0][2], h[1][2]);
        q[3] = make_float4(as_float(ref(l)), as_float(ref(r)), 0.0f, 0.0f);
    }
    // depth of the tree bounds the traversal stack (one pushed sibling per level)
    std::vector<int> dep(nn, 0);
    for (int i = 0; i < nn; ++i) {
        out.depth = std::max(out.depth, dep[i]);
        if (inner_id[i] >= 0) { dep[i + 1] = dep[i] + 1; dep[as_int(out.a[i + 1].w)] = dep[i] + 1; }
    }
    if (out.depth >= RR_BVH_STACK) return false;
    return true;
}

}  // namespace rr
