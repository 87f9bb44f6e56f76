// rr_trace.cuh — ray-trace mode of the per-pixel path (render.rs:993-1224) as device functions.
//
// Design (DESIGN.md "trace kernel"):
//   * scene-level raycast() is a brute-force scan like the reference's, but over two homogeneous
//     SoA lists (floors, then spheres), so the inner loop has no per-object kind dispatch. The first
//     RR_HEAD_FLOORS floors and RR_HEAD_SPHERES spheres travel in the kernel-parameter constant bank
//     and are tested in a fully unrolled sequence (their coordinates become c[0][..] operands of the
//     FADD/FMUL instructions: no loads, no registers); the rest of a larger scene is scanned from
//     shared memory. Ties are resolved to the lowest ORIGINAL object index, which is what the
//     reference's in-order strict `<` scan does (appendix A Q6).
//   * one raycast call site: a per-thread state machine alternates "trace ray" and "shadow ray"
//     phases, so the scan code exists once (I-cache) and lanes in different phases still share it.
//   * the refraction recursion shading()->raytrace() (render.rs:1093-1115) is unrolled into an
//     explicit per-thread stack of suspended parent frames; evaluation order of every float sum
//     is unchanged (the child colour is complete before the parent blends it).
#pragma once
#include "rr_device.cuh"

// tools/simt_model.cpp counts traversal steps per ray through these hooks (CPU build only); empty everywhere else
#ifndef RR_MODEL_RAY
#define RR_MODEL_RAY()
#define RR_MODEL_INNER()
#define RR_MODEL_LEAF(n)
#endif

namespace rr {

// pointers to the tails of the intersection lists (shared memory when staged)
struct SceneView {
    const float4 *sph;
    const int *sph_oi;
    const float4 *flo_o;
    const float4 *flo_n;
    const int *flo_oi;
    int n_spheres, n_floors;  // totals (head + tail)
    // BVH mode (scenes with many spheres): nodes and the leaf-ordered sphere copy
    const float4 *bvh_a, *bvh_b, *bvh_w, *bsph;
    const int *bsph_oi;
    int n_bvh_nodes;
    // ordered BVH traversal: the first RR_BVH_SMEM_STACK entries of every thread's stack live in shared memory
    // (entry e of thread t at stk[e * stk_stride + t]: conflict-free), deeper entries in local memory. nullptr: all local.
    uint2 *stk;
    int stk_stride;
    unsigned stk_s, bvh_w_s, bsph_s;  // shared-window addresses of stk / bvh_w / bsph when they are staged (SMEMBVH instances)
};

struct Hit {
    float t;
    int idx;
};

// RenderFloor::raycast, render.rs:557-569, folded into the running minimum of render.rs:1010-1014.
// Floors are scanned first and in index order, so strict `<` keeps the lowest index.
__device__ __forceinline__ void floor_test(const float4 &o, const float4 &n4, int oi, const V3 &vi, const V3 &eye,
                                           int ig, float &t, int &idx) {
    const V3 n = mk(n4.x, n4.y, n4.z);
    const V3 wpt = vi - mk(o.x, o.y, o.z);
    const float w = dot(n, eye);
    if (w <= 0.0f && oi != ig) {
        const float t0 = (-dot(n, wpt)) / w;
        if (t0 >= 0.0f && t0 < t) {
            t = t0;
            idx = oi;
        }
    }
}

// RenderSphere::raycast, render.rs:447-471.
// The reference forms b = 2*(eye.wpt), c = wpt.wpt - r*r, d2 = b*b - 4*c, d = sqrt(d2),
// t0 = (-b - d)/2, t1 = t0 + d. With D = eye.wpt and q = D*D - c this is, exactly (scalings by 2 and
// 4 commute with IEEE rounding): d2 = 4q, d = 2*sqrt(q), t0 = -D - sqrt(q), t1 = t0 + 2*sqrt(q), and
// `d2 >= EPSILON` <=> `q >= EPSILON/4`. Same bits, fewer multiplies.
template <typename OiFn>
__device__ __forceinline__ void sphere_finish(float D, float q, OiFn get_oi, int ig, bool near_ok, bool far_ok, float &t, int &idx) {
    if (q >= F32_EPS_QUARTER) {
        const float sq = sqrtf(q);
        const float t0 = -D - sq;
        float cand = RR_INF;
        if (near_ok && t0 >= 0.0f) {
            cand = t0;
        } else if (far_ok) {
            const float t1 = t0 + 2.0f * sq;
            if (0.0f < t1) cand = t1;
        }
        if (cand <= t) {
            // lowest original index wins exact ties (floors were scanned first, spheres are in index order)
            const int oi = get_oi();  // fetched only for candidates
            if (oi != ig && (cand < t || (cand < RR_INF && oi < idx))) {
                t = cand;
                idx = oi;
            }
        }
    }
}
template <typename OiFn>
__device__ __forceinline__ void sphere_test(const float4 &c4, OiFn get_oi, const V3 &vi, const V3 &eye, int ig, bool near_ok,
                                            bool far_ok, float &t, int &idx) {
    const V3 wpt = vi - mk(c4.x, c4.y, c4.z);
    const float D = dot(eye, wpt);
    const float c = dot(wpt, wpt) - c4.w;
    const float q = D * D - c;
    sphere_finish(D, q, get_oi, ig, near_ok, far_ok, t, idx);
}

// Two spheres per instruction (packed f32x2, rr_device.cuh F2): the same operations in the same order as sphere_test(),
// on the pair arrangement of SceneHead / FrameParams. 16 packed operations form (D, q) of both spheres (32 scalar ones).
#ifndef RR_PACKED_SCAN
#define RR_PACKED_SCAN 1  // 0: the same scan with scalar FMUL/FADD on constant-bank operands (A/B builds)
#endif
#ifndef RR_GROUP_FINISH
#define RR_GROUP_FINISH 1
#endif
#ifndef RR_INIT_SHADING_STATE
#define RR_INIT_SHADING_STATE 0  // 1: A/B build with start values for the shading state carried across the shadow ray
#endif
struct PairDQ { F2 D, q; };
__device__ __forceinline__ F2 dot2(const PackK &K, const F2 &ax, const F2 &ay, const F2 &az, const F2 &bx, const F2 &by, const F2 &bz) {
    return add2(K, add2(K, mul2(K, ax, bx), mul2(K, ay, by)), mul2(K, az, bz));  // (x*x' + y*y') + z*z'
}
__device__ __forceinline__ PairDQ sphere_pair_dq(const PackK &K, const float2 &cx, const float2 &cy, const float2 &cz, const float2 &rr,
                                                 const V3 &vi, const V3 &eye) {
    const F2 wx = sub2(K, f2b(vi.x), f2(cx.x, cx.y)), wy = sub2(K, f2b(vi.y), f2(cy.x, cy.y)), wz = sub2(K, f2b(vi.z), f2(cz.x, cz.y));
    PairDQ r;
    r.D = dot2(K, f2b(eye.x), f2b(eye.y), f2b(eye.z), wx, wy, wz);
    const F2 c = sub2(K, dot2(K, wx, wy, wz, wx, wy, wz), f2(rr.x, rr.y));
    r.q = sub2(K, mul2(K, r.D, r.D), c);
    return r;
}
// primary rays: wpt and c come from the frame constants (FrameParams::pw_*), 7 packed operations per pair
__device__ __forceinline__ PairDQ sphere_pair_dq_primary(const PackK &K, const float2 &wx, const float2 &wy, const float2 &wz,
                                                         const float2 &c, const V3 &eye) {
    PairDQ r;
    r.D = dot2(K, f2b(eye.x), f2b(eye.y), f2b(eye.z), f2(wx.x, wx.y), f2(wy.x, wy.y), f2(wz.x, wz.y));
    r.q = sub2(K, mul2(K, r.D, r.D), f2(c.x, c.y));
    return r;
}

// ---------------------------------------------------------------------------------------------
// BVH: an EXACT cull for scenes with many spheres. The reference scans every object; a sphere may
// be skipped only if its (f32, reference-order) test provably cannot change the result, i.e. it
// returns a miss or a candidate t strictly greater than the current best. The leaf test is the
// same bit-exact sphere_test(); only the decision to skip uses other arithmetic, with margins:
//   * rounding analysis of the reference's q = D*D - (wpt.wpt - r*r): |q_f32 - q_true| <= 21 u d^2
//     (u = 2^-24, d = |origin - centre|), so a sphere the f32 test can hit lies within
//     sqrt(r^2 + K d^2) of the ray with K = 1.25e-6; a direction of squared length 1 + delta
//     (|delta| <= 4e-6, checked per ray in raycast) adds delta*c <= 4e-6 d^2. Boxes are inflated per ray by
//     e = sqrt(r_min^2 + M D^2) - r_min + 1e-6 D, M = 1.2e-5 (2.3x the 5.25e-6 total), D = largest distance
//     from the ray origin to the scene bounds; the same inflation bounds |t_f32 - t_true| for the entry test.
//   * slab arithmetic itself gets 1e-6 relative slack on both interval ends.
// Ties keep the reference rule (lowest original index) because leaves apply the full lexicographic
// comparison whatever the visiting order. Verified bit-for-bit against the brute-force oracle in
// tests/test_parity_gpu.py (synthetic scenes, all depth limits) and against the brute-force
// kernel instance in tests/test_bvh_gpu.py.
// ---------------------------------------------------------------------------------------------
constexpr float RR_BVH_M = 1.2e-5f;

__device__ __forceinline__ void bvh_scan(const DevScene &G, const SceneView &S, const V3 &vi, const V3 &eye, int ig,
                                         bool near_ok, bool far_ok, float &t, int &idx) {
    // per-ray inflation (fast math is fine here: these values only gate which exact tests run)
    const float Dx = fmaxf(fabsf(vi.x - G.scene_lo[0]), fabsf(vi.x - G.scene_hi[0]));
    const float Dy = fmaxf(fabsf(vi.y - G.scene_lo[1]), fabsf(vi.y - G.scene_hi[1]));
    const float Dz = fmaxf(fabsf(vi.z - G.scene_lo[2]), fabsf(vi.z - G.scene_hi[2]));
    const float D2 = __fmaf_rn(Dx, Dx, __fmaf_rn(Dy, Dy, Dz * Dz));
    const float D = sqrtf(D2);
    const float e = sqrtf(__fmaf_rn(RR_BVH_M, D2, G.r_min * G.r_min)) - G.r_min + 1e-6f * D;
    // a zero direction component would give 0*inf = NaN in the slab products; nudge it (cull-only arithmetic)
    const float ex = fabsf(eye.x) < 1e-30f ? copysignf(1e-30f, eye.x) : eye.x;
    const float ey = fabsf(eye.y) < 1e-30f ? copysignf(1e-30f, eye.y) : eye.y;
    const float ez = fabsf(eye.z) < 1e-30f ? copysignf(1e-30f, eye.z) : eye.z;
    const float ix = __frcp_rn(ex), iy = __frcp_rn(ey), iz = __frcp_rn(ez);
    // origin shifted so each slab bound is one subtraction: (lo - e - o) = lo - (o + e)
    const float ox_lo = vi.x + e, oy_lo = vi.y + e, oz_lo = vi.z + e;
    const float ox_hi = vi.x - e, oy_hi = vi.y - e, oz_hi = vi.z - e;
    int node = 0;
    const int n = S.n_bvh_nodes;
    while (node < n) {
        const float4 a = S.bvh_a[node];
        const float4 b = S.bvh_b[node];
        const float t1x = (a.x - ox_lo) * ix, t2x = (b.x - ox_hi) * ix;
        const float t1y = (a.y - oy_lo) * iy, t2y = (b.y - oy_hi) * iy;
        const float t1z = (a.z - oz_lo) * iz, t2z = (b.z - oz_hi) * iz;
        const float tmin = fmaxf(fmaxf(fminf(t1x, t2x), fminf(t1y, t2y)), fminf(t1z, t2z));
        const float tmax = fminf(fminf(fmaxf(t1x, t2x), fmaxf(t1y, t2y)), fmaxf(t1z, t2z));
        const float tmin_lo = tmin - fabsf(tmin) * 1e-6f;
        const float tmax_hi = tmax + fabsf(tmax) * 1e-6f;
        const bool reject = (tmax_hi < tmin_lo) || (tmax_hi < 0.0f) || (tmin_lo > t);
        const int leaf = __float_as_int(b.w);
        if (reject) {
            node = __float_as_int(a.w);  // escape: skip the subtree
        } else if (leaf < 0) {
            node = node + 1;             // descend (depth-first order: the left child is next)
        } else {
            const int first = leaf >> 3, count = leaf & 7;
            for (int k = 0; k < count; ++k)
                sphere_test(S.bsph[first + k], [&] { return S.bsph_oi[first + k]; }, vi, eye, ig, near_ok, far_ok, t, idx);
            node = __float_as_int(a.w);
        }
    }
}

// Ordered variant (the default): front-to-back traversal with a short per-thread stack. An inner node holds BOTH
// child boxes (4 x float4, DevScene::bvh_w); both are tested when the node is visited, the nearer child is entered
// first and the farther one is pushed together with its entry distance, so that after a hit has shortened t the
// stacked subtrees beyond it are dropped without touching their nodes. The reject rule, its margins and the leaf test
// are those of bvh_scan(); the visiting order cannot change the result because sphere_test() applies the full
// (t, original index) comparison. Child reference: >= 0 inner node, < 0 leaf ~((first << 3) | count).

// Slab arithmetic of the ordered traversal (cull-only arithmetic, so contraction is allowed). A child box is stored as centre
// c and half extent h (h rounded up by the builder so that [c - h, c + h] contains the exact box), and per axis
//   t_near = (c - o) i - (h + e) |i|,   t_far = (c - o) i + (h + e) |i|        (i = 1/d, e = the per-ray inflation)
// which needs no min/max to sort the two planes: the sign of the direction is in |i|. Each is two fused multiply-adds,
// fma(h, -+|i|, fma(c, i, k)) with the per-ray constants k_near = -(o i) - e |i|, k_far = -(o i) + e |i|, and the left and the
// right child of a node share every instruction (FFMA2: the node record holds (c_L, c_R) and (h_L, h_R) side by side).
// Per node that is 12 FFMA2 + 8 min/max, against 6 FFMA2 + 20 min/max for lo/hi planes: FMNMX runs on the half-rate ALU
// pipe, which was the busiest unit of the BVH instance (ncu, profiles/r2j_trace_synthetic1024_4k.md: ALU 70 %, math-pipe
// throttle the top stall).
// Rounding (position space = t error / |i|): o i and k are rounded on their own (2 x 2^-24 |o|), fma(c, i, k) and the
// final fma once each (2 x 2^-24 of a distance <= D); e |i| adds 2^-24 e. These are covered by the 4e-7 max|o_k| term of e
// (3.3x) and, together with the reciprocal's 2.4e-7 D, by its 1e-6 D term (2.7x).
// reject <=> min(t_far's, t) < max(t_near's, 0)   (t >= 0 always; a NaN slab distance is ignored by fminf/fmaxf = no cull).
struct BoxT { float lo, hi; };

// cull-only helpers: approximate reciprocal / square root (MUFU, ~2^-22 relative error, no IEEE fix-up branches). Their
// error is inside the margins: every quantity built from them is padded upwards by >= 1e-6 relative below.
__device__ __forceinline__ float cull_rcp(float x) {
#ifdef RR_HOSTSIM
    return 1.0f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float cull_sqrt_up(float x) {  // >= sqrt(x) for finite x >= 0
#ifdef RR_HOSTSIM
    return sqrtf(x) * 1.000001f;
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 1.000001f;
#endif
}

// Node / leaf / stack accessors of the ordered traversal. SMEM: the tree and the hot part of the stack are in shared
// memory and are addressed with 32-bit shared-window addresses (LDS/STS with immediate offsets: one address add per node
// instead of a 64-bit multiply-add, and no generic-address resolution). Otherwise generic loads through L1.
struct NodeW { float4 n0, n1, n2; int l, r; };
template <bool SMEM>
__device__ __forceinline__ NodeW load_node(const SceneView &S, int cur /* byte offset of the 64-byte record */) {
    NodeW w;
#ifndef RR_HOSTSIM
    if constexpr (SMEM) {
        const unsigned a = S.bvh_w_s + (unsigned)cur;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.n0.x), "=f"(w.n0.y), "=f"(w.n0.z), "=f"(w.n0.w) : "r"(a));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(w.n1.x), "=f"(w.n1.y), "=f"(w.n1.z), "=f"(w.n1.w) : "r"(a));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+32];" : "=f"(w.n2.x), "=f"(w.n2.y), "=f"(w.n2.z), "=f"(w.n2.w) : "r"(a));
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2+48];" : "=r"(w.l), "=r"(w.r) : "r"(a));
        return w;
    }
#endif
    const float4 *q = reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(S.bvh_w) + cur);
    w.n0 = q[0]; w.n1 = q[1]; w.n2 = q[2];
    const float4 n3 = q[3];
    w.l = __float_as_int(n3.x); w.r = __float_as_int(n3.y);
    return w;
}
template <bool SMEM>
__device__ __forceinline__ float4 load_leaf_sphere(const SceneView &S, int k) {
#ifndef RR_HOSTSIM
    if constexpr (SMEM) {
        float4 c;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "r"(S.bsph_s + 16u * (unsigned)k));
        return c;
    }
#endif
    return S.bsph[k];
}

template <bool SMEM>
__device__ __forceinline__ void bvh_scan_ordered(const DevScene &G, const SceneView &S, const V3 &vi, const V3 &eye, int ig,
                                                 bool near_ok, bool far_ok, float &t, int &idx) {
    const float Dx = fmaxf(fabsf(vi.x - G.scene_lo[0]), fabsf(vi.x - G.scene_hi[0]));
    const float Dy = fmaxf(fabsf(vi.y - G.scene_lo[1]), fabsf(vi.y - G.scene_hi[1]));
    const float Dz = fmaxf(fabsf(vi.z - G.scene_lo[2]), fabsf(vi.z - G.scene_hi[2]));
    const float D2 = __fmaf_rn(Dx, Dx, __fmaf_rn(Dy, Dy, Dz * Dz));
    const float D = cull_sqrt_up(D2);
    // e >= sqrt(r_min^2 + M D^2) - r_min + 1e-6 D + 4e-7 max|o_k|  (the approximate square roots are rounded up by 1e-6)
    const float e = cull_sqrt_up(__fmaf_rn(RR_BVH_M, D2, G.r_min * G.r_min)) - G.r_min + 1e-6f * D +
                    4e-7f * fmaxf(fmaxf(fabsf(vi.x), fabsf(vi.y)), fabsf(vi.z));
    const float ex = fabsf(eye.x) < 1e-30f ? copysignf(1e-30f, eye.x) : eye.x;
    const float ey = fabsf(eye.y) < 1e-30f ? copysignf(1e-30f, eye.y) : eye.y;
    const float ez = fabsf(eye.z) < 1e-30f ? copysignf(1e-30f, eye.z) : eye.z;
    // 1/d to 2^-22: a relative error of t, i.e. <= 2.4e-7 D in position, inside the 1e-6 D term of e
    const float ix = cull_rcp(ex), iy = cull_rcp(ey), iz = cull_rcp(ez);
    const float aix = fabsf(ix), aiy = fabsf(iy), aiz = fabsf(iz);
    const float ox = vi.x * ix, oy = vi.y * iy, oz = vi.z * iz, eix = e * aix, eiy = e * aiy, eiz = e * aiz;
    const F2 bx = f2b(ix), by = f2b(iy), bz = f2b(iz);
    const F2 nax = f2b(-aix), nay = f2b(-aiy), naz = f2b(-aiz), pax = f2b(aix), pay = f2b(aiy), paz = f2b(aiz);
    const F2 knx = f2b(-ox - eix), kny = f2b(-oy - eiy), knz = f2b(-oz - eiz);
    const F2 kfx = f2b(-ox + eix), kfy = f2b(-oy + eiy), kfz = f2b(-oz + eiz);
    constexpr int DONE = 0x7fffffff;
    // Stack of postponed (farther) children with their entry distances. Its hot part is in shared memory: as two local
    // arrays it was the largest source of local-memory traffic of the BVH instance (profiles/r1_s2_trace_synthetic1024_4k.md:
    // 14 M local loads + 14 M local stores per 4K frame, 65 % L1 hit rate, 2.17x the framebuffer bytes in DRAM traffic).
    // Entry e of thread t sits at stk + 8 * (e * RR_BVH_BLOCK + t) (conflict-free); `sp` is that byte offset, so with a
    // power-of-two block the level is sp >> log2(8 * RR_BVH_BLOCK) and "fits in shared memory" is one compare with a
    // constant. Entry 0 holds a sentinel (DONE, -inf): popping it ends the traversal, so pop() has no empty-stack test.
    constexpr int SD = RR_BVH_SMEM_STACK;
    int ovf_ref[RR_BVH_STACK + 1 - SD];
    float ovf_t[RR_BVH_STACK + 1 - SD];
#ifdef RR_HOSTSIM
    uint2 hs_stk[SD];
    constexpr unsigned step = 8u;
    unsigned sp = 0;
#else
    constexpr unsigned step = 8u * (unsigned)RR_BVH_BLOCK;
    static_assert((RR_BVH_BLOCK & (RR_BVH_BLOCK - 1)) == 0, "block size of the BVH instances must be a power of two");
    unsigned sp = 8u * threadIdx.x;
#endif
    constexpr unsigned lim = (unsigned)SD * step;
    int cur = 0;
    auto push = [&](int ref, float tm) {
        if (sp < lim) {
#ifdef RR_HOSTSIM
            hs_stk[sp / step] = make_uint2((unsigned)ref, (unsigned)__float_as_int(tm));
#else
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(S.stk_s + sp), "r"(ref), "f"(tm) : "memory");
#endif
        } else {
            ovf_ref[(sp - lim) / step] = ref; ovf_t[(sp - lim) / step] = tm;
        }
        sp += step;
    };
    // pop the next stacked subtree that can still hold a winner (entered at or before the current best hit)
    auto pop = [&]() {
        float tm;
        do {
            sp -= step;
            if (sp < lim) {
#ifdef RR_HOSTSIM
                const uint2 en = hs_stk[sp / step];
                cur = (int)en.x;
                tm = __int_as_float((int)en.y);
#else
                asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(cur), "=f"(tm) : "r"(S.stk_s + sp) : "memory");
#endif
            } else {
                cur = ovf_ref[(sp - lim) / step];
                tm = ovf_t[(sp - lim) / step];
            }
        } while (tm > t);  // NaN: visit
    };
    push(DONE, -RR_INF);  // sentinel
    while (cur != DONE) {
        // inner nodes until this lane holds a leaf (the warp leaves the loop when every lane does: leaf tests then
        // run with more lanes active than in an if/else per step)
        while ((unsigned)cur < (unsigned)DONE) {
            RR_MODEL_INNER();
            const NodeW w = load_node<SMEM>(S, cur);
            // (left, right) pairs: n0 = (c.x, c.x, c.y, c.y), n1 = (c.z, c.z, h.x, h.x), n2 = (h.y, h.y, h.z, h.z)
            const F2 cx = f2(w.n0.x, w.n0.y), cy = f2(w.n0.z, w.n0.w), cz = f2(w.n1.x, w.n1.y);
            const F2 hx = f2(w.n1.z, w.n1.w), hy = f2(w.n2.x, w.n2.y), hz = f2(w.n2.z, w.n2.w);
            const F2 nx = fma2(hx, nax, fma2(cx, bx, knx)), fx = fma2(hx, pax, fma2(cx, bx, kfx));
            const F2 ny = fma2(hy, nay, fma2(cy, by, kny)), fy = fma2(hy, pay, fma2(cy, by, kfy));
            const F2 nz = fma2(hz, naz, fma2(cz, bz, knz)), fz = fma2(hz, paz, fma2(cz, bz, kfz));
            BoxT L, R;
            L.lo = fmaxf(fmaxf(f2lo(nx), f2lo(ny)), fmaxf(f2lo(nz), 0.0f));
            L.hi = fminf(fminf(f2lo(fx), f2lo(fy)), fminf(f2lo(fz), t));
            R.lo = fmaxf(fmaxf(f2hi(nx), f2hi(ny)), fmaxf(f2hi(nz), 0.0f));
            R.hi = fminf(fminf(f2hi(fx), f2hi(fy)), fminf(f2hi(fz), t));
            const bool rej_l = L.hi < L.lo, rej_r = R.hi < R.lo;
            const int l = w.l, r = w.r;
            if (!rej_l && !rej_r) {
                const bool left_first = L.lo <= R.lo;
                push(left_first ? r : l, left_first ? R.lo : L.lo);
                cur = left_first ? l : r;
            } else if (!rej_l) {
                cur = l;
            } else if (!rej_r) {
                cur = r;
            } else {
                pop();
            }
        }
        if (cur != DONE) {
            const int leaf = ~cur;
            const int first = leaf >> 3, count = leaf & 7;
            RR_MODEL_LEAF(count);
            for (int k = 0; k < count; ++k)
                sphere_test(load_leaf_sphere<SMEM>(S, first + k), [&] { return S.bsph_oi[first + k]; }, vi, eye, ig, near_ok, far_ok, t, idx);
            pop();
        }
    }
}

// scene-level raycast, render.rs:993-1018
// HEADONLY: the whole scene is in the SceneHead (<= RR_HEAD_FLOORS floors, <= RR_HEAD_SPHERES spheres; unused slots hold
// never-hit objects, fill_head_pairs): no count checks, no tail loops. PRIMARY: the ray starts at the camera (ig = -1,
// flags = 0) and the origin-dependent terms come from the frame constants.
template <bool BVH, bool HEADONLY, bool PRIMARY, bool SMEMBVH = false>
__device__ __forceinline__ Hit raycast(const DevScene &G, const SceneHead &H, const SceneView &S, const FrameParams &P, const V3 &vi,
                                       const V3 &eye, int ig, unsigned flags) {
    float t = RR_INF;
    int idx = 0;
    RR_MODEL_RAY();
#pragma unroll
    for (int f = 0; f < RR_HEAD_FLOORS; ++f) {
        if (HEADONLY || f < S.n_floors) {
            if (PRIMARY && !BVH) {
                // floor_test() with -(n.wpt) taken from the frame constants
                const float w = dot(mk(H.flo_n[f].x, H.flo_n[f].y, H.flo_n[f].z), eye);
                if (w <= 0.0f) {
                    const float t0 = P.pf_nd[f] / w;
                    if (t0 >= 0.0f && t0 < t) { t = t0; idx = H.flo_oi[f]; }
                }
            } else {
                floor_test(H.flo_o[f], H.flo_n[f], H.flo_oi[f], vi, eye, ig, t, idx);
            }
        }
    }
    if (!HEADONLY)
        for (int f = RR_HEAD_FLOORS; f < S.n_floors; ++f) floor_test(S.flo_o[f], S.flo_n[f], S.flo_oi[f], vi, eye, ig, t, idx);
    const bool near_ok = PRIMARY ? true : (flags & OUTONLY) == 0;
    const bool far_ok = PRIMARY ? true : (flags & INONLY) == 0;
    if constexpr (BVH) {
        // The cull is geometric; the reference's test is geometric only for unit directions (it drops the
        // |eye|^2 factor of the quadratic). Directions are unit to a few ulp everywhere except after a bounce
        // off an un-normalised floor normal (face_normal is used as given, render.rs:553-563, appendix A Q22),
        // where eye += n*(-2 eye.n) changes its length. Such rays (and NaN directions) scan every sphere.
        // |eye|^2 = 1 + delta perturbs the discriminant by delta*c <= 4e-6 d^2, inside the margin M (rr_trace.cuh).
        // Origins beyond 5e7 also scan every sphere (as do scenes with such coordinates: the builder refuses them): with
        // |1/d| capped at 1e30 (axis-parallel rays) the slab terms (c - o)/d stay below the f32 range, so no infinity can
        // turn a box the ray is inside of into a reject.
        const float e2 = dot(eye, eye);
        if (fabsf(e2 - 1.0f) <= 4e-6f && fmaxf(fmaxf(fabsf(vi.x), fabsf(vi.y)), fabsf(vi.z)) < 5e7f) {
#if RR_BVH_ORDERED
            bvh_scan_ordered<SMEMBVH>(G, S, vi, eye, ig, near_ok, far_ok, t, idx);
#else
            bvh_scan(G, S, vi, eye, ig, near_ok, far_ok, t, idx);
#endif
        } else {
            for (int s = 0; s < S.n_spheres; ++s)
                sphere_test(S.bsph[s], [&] { return S.bsph_oi[s]; }, vi, eye, ig, near_ok, far_ok, t, idx);
        }
    } else {
        // (D, q) of every head sphere first (slots beyond the scene's spheres hold never-hit padding, so there are no count
        // checks in either instance), then ONE branch: most rays of most warps (sky, floor) miss every sphere, and the
        // candidate logic of all of them is skipped together when no lane has q >= EPSILON/4 for any sphere (a NaN q is
        // ignored by fmaxf and fails the per-sphere test anyway).
        float Ds[2 * RR_HEAD_PAIRS], qs[2 * RR_HEAD_PAIRS];
#if RR_PACKED_SCAN
        const PackK K{f2(P.pk_one.x, P.pk_one.y), f2(P.pk_nz.x, P.pk_nz.y), f2(P.pk_neg1.x, P.pk_neg1.y)};
#pragma unroll
        for (int p = 0; p < RR_HEAD_PAIRS; ++p) {
            const PairDQ dq = PRIMARY ? sphere_pair_dq_primary(K, P.pw_x[p], P.pw_y[p], P.pw_z[p], P.pw_c[p], eye)
                                      : sphere_pair_dq(K, H.pcx[p], H.pcy[p], H.pcz[p], H.prr[p], vi, eye);
            Ds[2 * p] = f2lo(dq.D); qs[2 * p] = f2lo(dq.q);
            Ds[2 * p + 1] = f2hi(dq.D); qs[2 * p + 1] = f2hi(dq.q);
        }
#else
#pragma unroll
        for (int s = 0; s < 2 * RR_HEAD_PAIRS; ++s) {
            const int p = s >> 1, k = s & 1;  // compile-time after unrolling: the constants are c[][] operands of the FMUL/FADDs
            if (PRIMARY) {
                const V3 wpt = mk(k ? P.pw_x[p].y : P.pw_x[p].x, k ? P.pw_y[p].y : P.pw_y[p].x, k ? P.pw_z[p].y : P.pw_z[p].x);
                Ds[s] = dot(eye, wpt);
                qs[s] = Ds[s] * Ds[s] - (k ? P.pw_c[p].y : P.pw_c[p].x);
            } else {
                const V3 wpt = vi - mk(k ? H.pcx[p].y : H.pcx[p].x, k ? H.pcy[p].y : H.pcy[p].x, k ? H.pcz[p].y : H.pcz[p].x);
                Ds[s] = dot(eye, wpt);
                const float c = dot(wpt, wpt) - (k ? H.prr[p].y : H.prr[p].x);
                qs[s] = Ds[s] * Ds[s] - c;
            }
        }
#endif
        float qmax = qs[0];
#pragma unroll
        for (int s = 1; s < 2 * RR_HEAD_PAIRS; ++s) qmax = fmaxf(qmax, qs[s]);
        if (!RR_GROUP_FINISH || qmax >= F32_EPS_QUARTER) {
#pragma unroll
            for (int s = 0; s < RR_HEAD_SPHERES; ++s)  // (the padded high half of an odd head can never hit; its index is not stored)
                sphere_finish(Ds[s], qs[s], [&] { return H.sph_oi[s]; }, ig, near_ok, far_ok, t, idx);
        }
        if (!HEADONLY) {
#pragma unroll 4
            for (int s = RR_HEAD_SPHERES; s < S.n_spheres; ++s)
                sphere_test(S.sph[s], [&] { return S.sph_oi[s]; }, vi, eye, ig, near_ok, far_ok, t, idx);
        }
    }
    return Hit{t, idx};
}

// A suspended raytrace() frame waiting for its refraction child: 68 bytes, the last 24 (continuation ray) written and
// read only when the parent's bounce loop goes on after this hit.
struct TraceFrame {
    float ret[3];
    float fcs[3];   // fcs before this hit was accumulated
    float A[3];     // (kd*k1 + k2) * (1 - f)
    float f;
    unsigned meta;  // object that was hit (== ig of the continuation) | lev << 22 | continuation flags << 28 | cont << 31
    float vi[3];    // continuation ray (valid when cont)
    float eye[3];
};
// ig < 2^22 (rr_scene_create rejects larger scenes for the trace stack), lev <= 32 (6 bits), flags in {0, OUTONLY, INONLY}
__device__ __forceinline__ unsigned pack_meta(int ig, int lev, bool cont, unsigned flags) {
    return (unsigned)ig | ((unsigned)lev << 22) | (flags << 28) | (cont ? 0x80000000u : 0u);
}
__device__ __forceinline__ int meta_ig(unsigned m) { return (int)(m & 0x3fffffu); }
__device__ __forceinline__ int meta_lev(unsigned m) { return (int)((m >> 22) & 0x3fu); }
__device__ __forceinline__ unsigned meta_flags(unsigned m) { return (m >> 28) & 3u; }
__device__ __forceinline__ bool meta_cont(unsigned m) { return (m & 0x80000000u) != 0u; }

constexpr int RR_MAX_STACK = 32;  // >= max_refractions (checked on the host)

// One pixel. The loop is rotated: its body is "consume the hit of the last scan, set up the next ray, scan", and the
// first scan (the primary ray, whose origin-dependent terms are frame constants) is peeled off in front of it. There is
// still ONE general scan site, shared by trace rays and shadow rays of lanes in different phases.
template <bool COUNT, bool BVH, bool HEADONLY = false, bool SMEMBVH = false>
__device__ __forceinline__ V3 trace_pixel(const DevScene &G, const SceneHead &H, const SceneView &S, const FrameParams &P,
                                          const V3 &eye0 /* primary direction, render.rs:808-815 */, Counters &cnt) {
    const V3 light = mk(P.light[0], P.light[1], P.light[2]);
    // current trace ray of the running raytrace() frame
    V3 vi = mk(P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]);
    V3 eye = eye0;
    int lev = 1 /* render.rs:1157, first iteration */, ig = -1, depth = 0;
    unsigned flags = 0;
    V3 ret = mk(0.0f, 0.0f, 0.0f), fcs = mk(1.0f, 1.0f, 1.0f);
    TraceFrame stack[RR_MAX_STACK];
    int ray_class = 0;  // 0 primary, 1 refract child's first ray, 2 reflect continuation
    // shading() state carried across the shadow ray
    // (hidx, pt, n and the two intensities are written by the first half of shading() before its second half reads them;
    // giving them start values would cost a register move each per pixel at the loop entry)
    bool shadow_phase = false;  // which kind of ray produced `h`
#if RR_INIT_SHADING_STATE
    int hidx = 0;
    V3 pt = vi, n = vi;
    float diffuse_intensity = 0.0f, reflection_intensity = 0.0f;
#else
    int hidx;
    V3 pt, n;
    float diffuse_intensity, reflection_intensity;
#endif
    if (COUNT) {
        cnt.pixels++;
        cnt.primary++;
        cnt.object_tests += (unsigned long long)G.n_objects;
        cnt.sphere_tests += (unsigned long long)spheres_tested(G, -1);
    }
    // BVH instances do not peel the primary scan: a second copy of the traversal costs more in instruction-cache misses
    // than the peeled scan saves (measured: 2.05 -> 2.17 ms on the 1 024-sphere scene); they enter the loop at the scan.
    Hit h{RR_INF, 0};
    bool enter_at_scan = BVH;
    if (!BVH) h = raycast<BVH, HEADONLY, true, SMEMBVH>(G, H, S, P, vi, eye, -1, 0u);

    for (;;) {
        V3 ro = vi, rd = eye;
        int rig = -1;
        unsigned rfl = 0u;
        if (!enter_at_scan) {
        bool frame_done = false;
        if (!shadow_phase) {
            if (h.t < RR_INF) {
                // hit: first half of shading(), render.rs:1020-1046, then go cast the shadow ray
                hidx = h.idx;
                pt = (eye * h.t) + vi;  // render.rs:1164
                const float4 oa = __ldg(&G.obj_a[hidx]);
                const int4 ob = __ldg(&G.obj_b[hidx]);
                if (ob.x == 0) {
                    n = normalized(pt - mk(oa.x, oa.y, oa.z));  // render.rs:443-445
                } else {
                    const float4 n4 = __ldg(&G.obj_n[hidx]);     // render.rs:553-555
                    n = mk(n4.x, n4.y, n4.z);
                }
                const float light_incidence = dot(light, n);
                const float ln2 = 2.0f * light_incidence;
                const V3 rr_light = (n * ln2) - light;
                const int pn = G.mat[ob.z].pn;
                diffuse_intensity = fmaxf(light_incidence, 0.0f);
                reflection_intensity = 0.0f;
                if (pn != 0) {
                    const float ri = -dot(rr_light, eye);
                    if (ri > 0.0f) reflection_intensity = rs_powi(ri, pn);
                }
                shadow_phase = true;
            } else {
                if (COUNT) cnt.bg_evals++;
                const V3 bg = bgcolor(P, eye);  // render.rs:1213-1216
                ret = mk(ret.x + bg.x * fcs.x, ret.y + bg.y * fcs.y, ret.z + bg.z * fcs.z);
                frame_done = true;
            }
        } else {
            // ---- second half of shading(), render.rs:1048-1139 ----
            shadow_phase = false;
            const int idx = hidx;
            const float4 oa = __ldg(&G.obj_a[idx]);
            const int4 ob = __ldg(&G.obj_b[idx]);
            const DevMaterial &m = G.mat[ob.z];
            float k1 = 0.2f, k2 = 0.0f;
            bool lit = h.t >= RR_INF;
            if (!lit) lit = 0.0f < G.mat[__ldg(&G.obj_b[h.idx]).z].t;
            if (lit) {
                k1 = fminf(k1 + diffuse_intensity, 1.0f);
                k2 = reflection_intensity;
            }
            const V3 kd = get_diffuse(G, m, pt - mk(oa.x, oa.y, oa.z), ob.y);
            const V3 face = mk(kd.x * k1 + k2, kd.y * k1 + k2, kd.z * k1 + k2);
            const V3 ks = mk(m.specular[0], m.specular[1], m.specular[2]);

            if (lev < P.max_refractions && 0.0f < m.t) {
                // refraction child, render.rs:1093-1115: suspend this frame
                const float sp = dot(eye, n);
                const float f = m.t;
                const float frac = m.n;
                const float reference = sp * ((sp > 0.0f ? frac : 1.0f / frac) - 1.0f);
                const V3 ray = normalized(eye + (n * reference));
                const V3 pt3 = pt + (ray * F32_EPSILON);
                const float omf = 1.0f - f;
                TraceFrame &F = stack[depth];
                F.ret[0] = ret.x; F.ret[1] = ret.y; F.ret[2] = ret.z;
                F.fcs[0] = fcs.x; F.fcs[1] = fcs.y; F.fcs[2] = fcs.z;
                F.A[0] = face.x * omf; F.A[1] = face.y * omf; F.A[2] = face.z * omf;
                F.f = f;
                // what the parent does after `ret += face*fcs; fcs *= ks` (render.rs:1175-1211)
                const V3 nf = mk(fcs.x * ks.x, fcs.y * ks.y, fcs.z * ks.z);
                const bool cont = !(idx == 0) && !((nf.x + nf.y + nf.z) <= 0.1f) && !(lev >= P.max_reflections);
                F.meta = pack_meta(idx, lev, cont, 0u);
                if (cont) {
                    const float en2 = -2.0f * dot(eye, n);
                    const V3 e2 = eye + n * en2;
                    F.vi[0] = pt.x; F.vi[1] = pt.y; F.vi[2] = pt.z;
                    F.eye[0] = e2.x; F.eye[1] = e2.y; F.eye[2] = e2.z;
                    F.meta = pack_meta(idx, lev, true, dot(n, e2) < 0.0f ? OUTONLY : INONLY);
                }
                depth += 1;
                vi = pt3;
                eye = ray;
                ig = idx;
                flags = sp < 0.0f ? OUTONLY : INONLY;
                ret = mk(0.0f, 0.0f, 0.0f);
                fcs = mk(1.0f, 1.0f, 1.0f);
                ray_class = 1;
                // the child starts with lev = nest; the increment below makes it nest + 1
            } else {
                // ---- back in raytrace(), render.rs:1173-1211 ----
                ret = mk(ret.x + face.x * fcs.x, ret.y + face.y * fcs.y, ret.z + face.z * fcs.z);
                fcs = mk(fcs.x * ks.x, fcs.y * ks.y, fcs.z * ks.z);
                if (idx == 0 || (fcs.x + fcs.y + fcs.z) <= 0.1f || lev >= P.max_reflections) {
                    frame_done = true;
                } else {
                    vi = pt;
                    const float en2 = -2.0f * dot(eye, n);
                    eye = eye + n * en2;
                    flags = dot(n, eye) < 0.0f ? OUTONLY : INONLY;
                    ig = idx;
                    ray_class = 2;
                }
            }
        }

        // return from finished frames into their suspended parents (render.rs:1128-1132 then :1175-1211)
        while (frame_done) {
            if (depth == 0) return ret;
            depth -= 1;
            const TraceFrame &F = stack[depth];
            const float f = F.f;
            const V3 face = mk(F.A[0] + ret.x * f, F.A[1] + ret.y * f, F.A[2] + ret.z * f);
            ret = mk(F.ret[0] + face.x * F.fcs[0], F.ret[1] + face.y * F.fcs[1], F.ret[2] + face.z * F.fcs[2]);
            const unsigned meta = F.meta;
            if (meta_cont(meta)) {
                ig = meta_ig(meta);
                const DevMaterial &pm = G.mat[__ldg(&G.obj_b[ig]).z];
                fcs = mk(F.fcs[0] * pm.specular[0], F.fcs[1] * pm.specular[1], F.fcs[2] * pm.specular[2]);
                vi = mk(F.vi[0], F.vi[1], F.vi[2]);
                eye = mk(F.eye[0], F.eye[1], F.eye[2]);
                flags = meta_flags(meta);
                lev = meta_lev(meta);
                ray_class = 2;
                frame_done = false;
            }
        }

        // ---- the one general scene scan: the frame's next trace ray or the shadow ray of the hit ----
        if (!shadow_phase) {
            lev += 1;  // render.rs:1157
            ro = vi; rd = eye; rig = ig; rfl = flags;
            if (COUNT) {
                if (ray_class == 1) cnt.refract++;
                else cnt.reflect++;
            }
        } else {
            ro = pt + mk(P.light_eps[0], P.light_eps[1], P.light_eps[2]);  // pt + light * EPSILON, render.rs:1034 (the product is a frame constant)
            rd = light; rig = hidx; rfl = 0u;
            if (COUNT) {
                cnt.shadow++;
                if (__ldg(&G.obj_b[hidx]).x == 0) cnt.sphere_hits++;
            }
        }
        if (COUNT) {
            cnt.object_tests += (unsigned long long)(G.n_objects - (rig >= 0 ? 1 : 0));
            cnt.sphere_tests += (unsigned long long)spheres_tested(G, rig);
        }
        }  // !enter_at_scan
        enter_at_scan = false;
        h = raycast<BVH, HEADONLY, false, SMEMBVH>(G, H, S, P, ro, rd, rig, rfl);
    }
}

template <bool COUNT, bool BVH, bool HEADONLY = false, bool SMEMBVH = false>
__device__ __forceinline__ V3 trace_pixel(const DevScene &G, const SceneHead &H, const SceneView &S, const FrameParams &P,
                                          int ix, int iy, Counters &cnt) {
    return trace_pixel<COUNT, BVH, HEADONLY, SMEMBVH>(G, H, S, P, primary_dir_tab(P, __ldg(&P.ptab[ix]), __ldg(&P.ptab[P.xres + iy])), cnt);
}

}  // namespace rr
