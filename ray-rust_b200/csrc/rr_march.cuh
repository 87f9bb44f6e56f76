// rr_march.cuh — ray-march mode of the per-pixel path (render.rs:1226-1411) as device functions.
//
// raymarch_single()'s sphere-tracing loop is the hot loop (up to 10 001 dependent iterations, each
// an O(N) distance scan). Design:
//   * the scan runs over the SceneHead (first floors/spheres as constant-bank operands, fully
//     unrolled) and then over shared-memory tails, like the trace kernel;
//   * EXACT sqrt skipping: distance_estimate only needs the minimum of max(|c - p| - r, 0). A sphere
//     whose squared centre distance exceeds ((best + r) * (1 + 2^-19))^2 provably has an f32 distance
//     strictly above the running minimum (proof in sphere_dist), so its sqrt/sub/max/compare are
//     skipped without changing a bit of the result. Spheres with a glowing material are always
//     evaluated when glow is tracked, because the glow minimum needs their distance;
//   * the same idea one level up: the head spheres carry a bounding sphere (C, R) (SceneHead::grp, built by the host);
//     when |C - p|^2 > ((best + R) * (1 + 1e-5))^2 none of them can reach the running minimum and the whole unrolled
//     scan is skipped (proof in distance_estimate). The glow pass skips the sqrt of a glowing sphere whose glow value
//     provably cannot go below the minimum accumulated so far (proof there);
//   * creeping step: rays that crawl along the one floor far from every sphere (74 % of all steps) form the next
//     position and its floor distance before the branch on the "spheres far" test, which shortens the loop-carried
//     dependency chain that bounds the frame time (raymarch_single);
//   * large scenes: the sphere scan is pruned through the trace kernel's BVH by point-to-box distance (MBVH instances,
//     proof in distance_estimate);
//   * one march call site: a per-thread state machine alternates "trace march" and "shadow march",
//     so the loop body exists once; the glow distance (render.rs:1244-1247) is tracked only for trace
//     marches and only when --gloweffect is set and a glowing material exists (nothing else reads it);
//   * the march-mode miss quirk (render.rs:1385-1393) is executed as "march once, replay the
//     accumulation" (the repeated marches are bit-identical).
#pragma once
#include "rr_device.cuh"

namespace rr {

struct MarchView {  // list tails (shared memory when staged), indexed with the global list index
    const float4 *sph;     // (cx, cy, cz, r)
    const float *sph_glow; // glow_dist per sphere
    const int *sph_oi;
    const float4 *flo_o;   // (ox, oy, oz, glow_dist)
    const float4 *flo_n;
    const int *flo_oi;
    int n_spheres, n_floors;
    // large scenes (MBVH instances): the trace kernel's depth-first BVH arrays and the leaf-ordered (cx, cy, cz, r) copy
    const float4 *bvh_a, *bvh_b, *bsph;
    const int *bsph_oi;
    int n_bvh_nodes;
    float scene_abs;  // largest |coordinate| of the scene bounds
};

struct MarchResult {  // render.rs:1257-1264
    float final_dist;
    int idx;
    V3 pos;
    int iter;
    float travel_dist;
    float min_dist;
};

// RenderFloor::distance, render.rs:571-573, folded into the running minimum of render.rs:1238-1247
template <int GLOW>
__device__ __forceinline__ void floor_dist(const float4 &o, const float4 &n, int oi, const V3 &vi, int ig, bool track,
                                           float &best, int &idx, float &gl) {
    if (oi == ig) return;
    const float dist = fmaxf(dot(vi - mk(o.x, o.y, o.z), mk(n.x, n.y, n.z)), 0.0f);
    if (dist < best) {  // floors are scanned first and in index order: strict <
        best = dist;
        idx = oi;
    }
    if (GLOW == 2 && track) {
        const float g = dist * o.w;
        if (0.0f < g && g < gl) gl = g;
    }
}

// RenderSphere::distance, render.rs:473-475: max(len(org - vi) - r, 0).
//
// Skip rule. Let T = fl(best + r) with r >= 0 and sq = the f32 squared length. If
// sq > fl(fl(T*T) * 1.000004) then, with u = 2^-24: sq > (best+r)^2 (1 + 3.6e-6), so the correctly
// rounded sqrt is >= (best+r)(1 + 1.7e-6), and fl(sqrt - r) >= (best + 1.7e-6 (best+r)) (1-u) > best.
// Hence dist > best strictly: the sphere can neither lower the minimum nor tie with it, and
// skipping it leaves (best, idx) exactly as the reference's scan would. best = inf never skips.
template <int GLOW>
__device__ __forceinline__ void sphere_dist(const float4 &c, float glow, int oi, const V3 &vi, int ig, bool track,
                                            float &best, int &idx, float &gl) {
    if (oi == ig) return;
    const V3 d = mk(c.x, c.y, c.z) - vi;
    const float sq = d.x * d.x + d.y * d.y + d.z * d.z;
    const bool glows = GLOW == 2 && track && glow != 0.0f;
    if (!glows && c.w >= 0.0f) {
        const float T = best + c.w;
        if (sq > T * T * 1.000004f) return;
    }
    const float dist = fmaxf(sqrtf(sq) - c.w, 0.0f);
    if (dist < best || (dist == best && oi < idx && dist < RR_INF)) {  // lowest original index wins ties
        best = dist;
        idx = oi;
    }
    if (glows) {
        const float g = dist * glow;
        if (0.0f < g && g < gl) gl = g;
    }
}

// glow pass (render.rs:1244-1247) over the few glowing objects only (GLOW == 1); see distance_estimate
__device__ __forceinline__ void glow_pass(const SceneHead &H, const V3 &vi, int ig, float &gl) {
    // glow pass (render.rs:1244-1247) over the few glowing objects only; their distance is formed
    // with the same operations as in the scan, so the bits are the same whether or not the scan
    // above skipped the object's sqrt.
    // Sqrt skip: a glowing sphere (r >= 0, k = glow_dist > 0; the host stores glow_ik = fl(1/k), NaN otherwise) can
    // only lower gl if fl(dist * k) < gl. With T = fl(fl(gl * ik) + r) >= (gl/k + r)(1 - 3u): sq > fl(fl(T*T) * 1.00002)
    // implies sqrt(sq) >= (gl/k + r)(1 + 9.3e-6), dist >= (gl/k)(1 + 9.2e-6), fl(dist * k) > gl: no update. gl = inf or a
    // NaN anywhere makes the comparison false and the value is computed.
#pragma unroll 1
    for (int g = 0; g < H.n_glow_head; ++g) {  // rolled: keeps the march loop inside the L0 I-cache
        if (H.glow_oi[g] != ig) {
            const float4 a = H.glow_a[g];
            float dist;
            if (H.glow_kind[g] == 0) {
                const V3 d = mk(a.x, a.y, a.z) - vi;
                const float sq = d.x * d.x + d.y * d.y + d.z * d.z;
                const float T = gl * H.glow_ik[g] + a.w;
                if (sq > T * T * 1.00002f) continue;
                dist = fmaxf(sqrtf(sq) - a.w, 0.0f);
            } else {
                const float4 nn = H.glow_b[g];
                dist = fmaxf(dot(vi - mk(a.x, a.y, a.z), mk(nn.x, nn.y, nn.z)), 0.0f);
            }
            const float gv = dist * H.glow_k[g];
            if (0.0f < gv && gv < gl) gl = gv;
        }
    }
}

// distance_estimate, render.rs:1226-1251. `glowing` is in/out: the smallest glow value seen so far by the caller
// (render.rs:1281 takes the minimum over the march, :1331 over the marches of a frame; only that minimum is used).
//
// Group skip. The host stores in H.grp a sphere (C, R) with |C - c_s| + r_s <= R for every head sphere s (all r_s >= 0,
// else R = -1 = disabled). For any p: |c_s - p| - r_s >= |C - p| - R. If the f32 squared distance sqC (same operation
// order as sphere_dist: relative error <= 5u, u = 2^-24) satisfies sqC > fl(fl(T*T) * 1.00001), T = fl(best + R), then
// |C - p| =: x > (best + R)(1 + 4.4e-6), and the distance the reference computes for s, fl(fl(sqrt(sq_s)) - r_s), is
// >= (x - R) - 3.5u (x + R) - u x >= best (1 + 4.1e-6) + 3.9e-6 R > best: no head sphere can lower or tie the running
// minimum, exactly the situation in which sphere_dist() returns early for each of them. best = inf or NaN never skips.
//
// Large scenes (MBVH): the spheres are visited through the BVH the trace kernel uses (stackless depth-first order with
// escape indices; leaves hold <= 4 spheres), pruned by point-to-box distance. A node box holds the balls of its spheres
// up to the f32 rounding of its corners (<= u S, S = scene_abs). With m = 3e-6 (|p|_inf + S): d2 > fl(fl(T*T) * 1.00001),
// T = fl(best + m), d2 the f32 squared distance from p to the box, implies dist(p, box) > (best + m)(1 + 4.5e-6), hence for
// every sphere below the node |c_s - p| - r_s > best + m - 1.1e-7 S, and the distance the reference computes for it,
// fl(fl(sqrt(sq_s)) - r_s) >= that - 3.5u |c_s - p| - u rho_s >= best + m - 7e-7 (|p|_inf + S) > best: it can neither lower
// nor tie the running minimum. Leaves run the unchanged sphere_dist() (with its own exact sqrt skip), so the visiting
// order does not matter either (ties go to the lowest original index there). Inline glow (GLOW == 2) is not combined
// with this instance: the launcher keeps the linear scan for it.
template <int GLOW, bool MBVH = false>
__device__ __forceinline__ void distance_estimate(const SceneHead &H, const MarchView &S, const V3 &vi, int ig, bool track,
                                                  float &closest, int &idx_out, float &glowing, bool &all_far) {
    float best = RR_INF, gl = glowing;
    int idx = 0;
#pragma unroll
    for (int f = 0; f < RR_HEAD_FLOORS; ++f)
        if (f < S.n_floors) floor_dist<GLOW>(H.flo_o[f], H.flo_n[f], H.flo_oi[f], vi, ig, track, best, idx, gl);
    for (int f = RR_HEAD_FLOORS; f < S.n_floors; ++f)
        floor_dist<GLOW>(S.flo_o[f], S.flo_n[f], S.flo_oi[f], vi, ig, track, best, idx, gl);
    if constexpr (MBVH) {
        static_assert(GLOW != 2, "inline glow needs every glowing object's distance: linear scan only");
        const float m = 3e-6f * (fmaxf(fmaxf(fabsf(vi.x), fabsf(vi.y)), fabsf(vi.z)) + S.scene_abs);
        int node = 0;
        while (node < S.n_bvh_nodes) {
            const float4 a = S.bvh_a[node], b = S.bvh_b[node];
            const float dx = fmaxf(fmaxf(a.x - vi.x, vi.x - b.x), 0.0f);
            const float dy = fmaxf(fmaxf(a.y - vi.y, vi.y - b.y), 0.0f);
            const float dz = fmaxf(fmaxf(a.z - vi.z, vi.z - b.z), 0.0f);
            const float d2 = dx * dx + dy * dy + dz * dz;
            const float T = best + m;
            const int leaf = __float_as_int(b.w);
            if (d2 > T * T * 1.00001f) {  // best = inf or a NaN anywhere: false, the subtree is visited
                node = __float_as_int(a.w);
            } else if (leaf < 0) {
                node = node + 1;
            } else {
                const int first = leaf >> 3, count = leaf & 7;
                for (int k = 0; k < count; ++k)
                    sphere_dist<GLOW>(S.bsph[first + k], 0.0f, S.bsph_oi[first + k], vi, ig, track, best, idx, gl);
                node = __float_as_int(a.w);
            }
        }
        if (GLOW == 1 && track) glow_pass(H, vi, ig, gl);
        closest = best;
        idx_out = idx;
        glowing = gl;
        all_far = false;
        return;
    }
    // The same sphere also bounds the glowing spheres of the glow pass when the host says so (H.grp_ik = fl(1 / k_min),
    // k_min the smallest glow_dist, NaN otherwise): with x = |C - p| > (gl / k_min + R)(1 + 9.3e-6) every glowing sphere g
    // has fl(dist_g * k_g) >= (x - R - 3.5u (x + R) - u x) k_min (1 - u) > gl, so the whole pass cannot lower gl.
    bool head_far = false, glow_far = false;
    if (H.grp.w >= 0.0f && !(GLOW == 2 && track)) {
        const V3 d = mk(H.grp.x, H.grp.y, H.grp.z) - vi;
        const float sq = d.x * d.x + d.y * d.y + d.z * d.z;
        const float T = best + H.grp.w;
        head_far = sq > T * T * 1.00001f;
        if (GLOW == 1 && track) {
            const float Tg = gl * H.grp_ik + H.grp.w;
            glow_far = sq > Tg * Tg * 1.00002f;
        }
    }
    if (!head_far) {
#pragma unroll
        for (int s = 0; s < RR_HEAD_SPHERES; ++s)
            if (s < S.n_spheres) sphere_dist<GLOW>(H.sph_m[s], H.sph_glow[s], H.sph_oi[s], vi, ig, track, best, idx, gl);
    }
#pragma unroll 2
    for (int s = RR_HEAD_SPHERES; s < S.n_spheres; ++s)
        sphere_dist<GLOW>(S.sph[s], GLOW == 2 ? S.sph_glow[s] : 0.0f, S.sph_oi[s], vi, ig, track, best, idx, gl);
    if (GLOW == 1 && track && !glow_far) glow_pass(H, vi, ig, gl);
    closest = best;
    idx_out = idx;
    glowing = gl;
    all_far = head_far && (glow_far || !(GLOW == 1 && track));  // nothing but the floors decided this step
}

// The two bounding-sphere tests of distance_estimate() for a given running minimum `best` (= the floor distance).
template <int GLOW>
__device__ __forceinline__ bool spheres_far(const SceneHead &H, const V3 &vi, bool track, float best, float gl) {
    const V3 d = mk(H.grp.x, H.grp.y, H.grp.z) - vi;
    const float sq = d.x * d.x + d.y * d.y + d.z * d.z;
    const float T = best + H.grp.w;
    bool far = sq > T * T * 1.00001f;
    if (GLOW == 1 && track) {
        const float Tg = gl * H.grp_ik + H.grp.w;
        far = far && sq > Tg * Tg * 1.00002f;
    }
    return far;
}

// raymarch_single, render.rs:1266-1297. `glow_bound`: the smallest glow value the calling frame has seen so far (inf
// for a fresh frame); the returned min_dist is min(glow_bound, this march's minimum), which is all the caller uses.
//
// Creeping loop. 74 % of all steps belong to rays that crawl along the one floor far from every sphere; such a pixel is
// a chain of up to ~6 marches of 10 001 DEPENDENT steps, and the frame cannot end before its slowest chain does (8 GPUs
// render this frame in 5.0 ms against 7.7 ms on one: the chain, not throughput, is the bound). In that regime a step is:
// floor distance df, "spheres far" test, pos += eye * df. Written naively the far test sits on the loop-carried path
// (df -> T -> T*T -> compare -> branch -> next df, ~70 cycles). The loop below forms the NEXT position and its floor
// distance before the branch on the far test, so the carried chain is df -> pos' -> df' (~30 cycles) and the test
// resolves beside it. It is entered only after a full step has reported all_far, and it performs exactly the
// arithmetic distance_estimate() + the step update would perform for such a step (same operations, same order, idx = the
// floor's index); anything else (a NaN or infinite distance, a failed test, an ignored floor, more floors or
// spheres than the head holds, inline glow) leaves it without having changed any state and takes the full step.
template <int GLOW, bool MBVH = false>
__device__ __forceinline__ MarchResult raymarch_single(const SceneHead &H, const MarchView &S, const PackK &K, const V3 &init_pos,
                                                       const V3 &eye, int ig, bool track, float glow_bound) {
    int iter = 0;
    float travel = 0.0f;
    V3 pos = init_pos;
    float min_dist = glow_bound;
    const bool creep_ok = !MBVH && RR_HEAD_FLOORS >= 1 && S.n_floors == 1 && S.n_spheres <= RR_HEAD_SPHERES && H.flo_oi[0] != ig &&
                          H.grp.w >= 0.0f && !(GLOW == 2 && track);
    // ONE loop: in every iteration a lane takes either the creeping step or the full step, so that the lanes of a warp
    // advance together (an inner creeping loop would make the lanes that need full steps wait for thousands of
    // iterations). Warps whose lanes all creep - the tiles next to the horizon - skip the full step warp-wide.
    const V3 fo = mk(H.flo_o[0].x, H.flo_o[0].y, H.flo_o[0].z), fn = mk(H.flo_n[0].x, H.flo_n[0].y, H.flo_n[0].z);
    bool creeping = false;  // the previous step reported all_far; df = floor distance at pos
    float df = 0.0f;
    for (;;) {
        // Warp-wide creeping: when EVERY lane of the warp that is still in this march creeps (the tiles next to the horizon,
        // for thousands of iterations), the step runs in a tight inner loop without the full step's code around it, and in
        // packed arithmetic: the x and y components of every vector operation share one instruction (F2, rr_device.cuh:
        // the same IEEE operations, unfused, in the same order), as do the two thresholds of the far test. The loop is
        // left, with nothing committed for the current step, as soon as any lane's step is not a plain creeping step
        // (failed far test, termination); the general iteration below then handles every lane individually.
        {
            const unsigned act = __activemask();
            if (__all_sync(act, creeping)) {
                const F2 exy = f2(eye.x, eye.y), foxy = f2(fo.x, fo.y), fnxy = f2(fn.x, fn.y), gxy = f2(H.grp.x, H.grp.y);
                const F2 rr2 = f2b(H.grp.w), margin = f2(1.00001f, 1.00002f);
                const bool two = GLOW == 1 && track;  // the glow threshold applies as well
                F2 pxy = f2(pos.x, pos.y);
                float pz = pos.z;
                for (;;) {
                    // npos = (eye * df) + pos; ndf = max(dot(npos - fo, fn), 0)
                    const F2 nxy = add2(K, mul2(K, exy, f2b(df)), pxy);
                    const float nz = eye.z * df + pz;
                    const F2 qxy = mul2(K, sub2(K, nxy, foxy), fnxy);
                    const float ndf = fmaxf((f2lo(qxy) + f2hi(qxy)) + (nz - fo.z) * fn.z, 0.0f);
                    // spheres_far(): d = grp - pos, sq = d.d, sq > (df + R)^2 * 1.00001 [and sq > (gl * ik + R)^2 * 1.00002]
                    const F2 dxy = sub2(K, gxy, pxy);
                    const float dz = H.grp.z - pz;
                    const F2 sxy = mul2(K, dxy, dxy);
                    const float sq = (f2lo(sxy) + f2hi(sxy)) + dz * dz;
                    const F2 tt = add2(K, f2(df, min_dist * H.grp_ik), rr2);
                    const F2 th = mul2(K, mul2(K, tt, tt), margin);
                    const bool far = (sq > f2lo(th)) & (!two | (sq > f2hi(th)));
                    const bool ok = (df < RR_INF) & far & !(df < RAYMARCH_EPS) & !(FAR_AWAY < df) & !(MAX_ITER < iter + 1);
                    if (!__all_sync(act, ok)) break;
                    travel += df;
                    iter += 1;
                    pxy = nxy;
                    pz = nz;
                    df = ndf;
                }
                pos = mk(f2lo(pxy), f2hi(pxy), pz);
            }
        }
        if (creeping) {
            const V3 npos = (eye * df) + pos;                      // speculative: the step if it is a creeping one
            const float ndf = fmaxf(dot(npos - fo, fn), 0.0f);     // ... and the floor distance after it
            if (df < RR_INF && spheres_far<GLOW>(H, pos, track, df, min_dist)) {
                travel += df;
                iter += 1;
                if (df < RAYMARCH_EPS || FAR_AWAY < df || MAX_ITER < iter) return MarchResult{df, H.flo_oi[0], npos, iter, travel, min_dist};
                pos = npos;
                df = ndf;
                continue;
            }
        }
        float dist;
        int idx;
        bool all_far;
        distance_estimate<GLOW, MBVH>(H, S, pos, ig, track, dist, idx, min_dist, all_far);
        pos = (eye * dist) + pos;
        travel += dist;
        iter += 1;
        if (dist < RAYMARCH_EPS || FAR_AWAY < dist || MAX_ITER < iter) return MarchResult{dist, idx, pos, iter, travel, min_dist};
        creeping = creep_ok && all_far;
        if (creeping) df = fmaxf(dot(pos - fo, fn), 0.0f);  // RenderFloor::distance at the new pos
    }
}

struct MarchFrame {  // a suspended raymarch() frame waiting for its refraction child
    float ret[3];
    float fcs[3];
    float A[3];
    float f;
    float pos[3];
    float eye[3];
    float mmd;  // min_min_dist of the suspended frame
    int ig;
    int lev;
    int cont;
};

constexpr int RR_MARCH_MAX_STACK = 32;

// glow factor, render.rs:1397-1410. powf is CUDA's (<= 2-4 ulp from glibc's).
__device__ __forceinline__ V3 apply_glow(const FrameParams &P, const V3 &c, float mmd) {
    if (!P.glow_enabled) return c;
    const float factor = (mmd == RR_INF) ? 1.0f : 1.0f + (0.0f + P.glow_effect * powf(0.99f, mmd));
    return mk(factor * c.x, factor * c.y, factor * c.z);
}

template <bool COUNT, int GLOW, bool MBVH = false>
__device__ __forceinline__ V3 march_pixel(const DevScene &G, const SceneHead &H, const MarchView &S, const FrameParams &P,
                                          int ix, int iy, Counters &cnt) {
    const V3 light = mk(P.light[0], P.light[1], P.light[2]);
    const PackK K{f2(P.pk_one.x, P.pk_one.y), f2(P.pk_nz.x, P.pk_nz.y), f2(P.pk_neg1.x, P.pk_neg1.y)};
    V3 pos = mk(P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]);
    V3 eye = primary_ray(P, ix, iy);
    int lev = 0, ig = -1, depth = 0;
    V3 ret = mk(0.0f, 0.0f, 0.0f), fcs = mk(1.0f, 1.0f, 1.0f);
    float mmd = RR_INF;
    MarchFrame stack[RR_MARCH_MAX_STACK];
    int ray_class = 0;
    bool shadow_phase = false;
    int hidx = 0;
    V3 pt = pos, n = pos;
    float diffuse_intensity = 0.0f, reflection_intensity = 0.0f;
    if (COUNT) cnt.pixels++;

    for (;;) {
        // ---- the one march: the frame's trace ray or the shadow ray of a hit ----
        V3 ro, rd;
        int rig;
        if (!shadow_phase) {
            lev += 1;  // render.rs:1317
            ro = pos; rd = eye; rig = ig;
        } else {
            ro = pt + mk(P.light_eps[0], P.light_eps[1], P.light_eps[2]);  // pt + light * EPSILON, render.rs:1034 (frame constant)
            rd = light; rig = hidx;
        }
        const MarchResult r = raymarch_single<GLOW, MBVH>(H, S, K, ro, rd, rig, !shadow_phase, shadow_phase ? RR_INF : mmd);
        if (COUNT) {
            if (shadow_phase) {
                cnt.shadow++;
                if (__ldg(&G.obj_b[hidx]).x == 0) cnt.sphere_hits++;
            } else if (ray_class == 0) cnt.primary++;
            else if (ray_class == 1) cnt.refract++;
            else cnt.reflect++;
            cnt.march_steps += (unsigned long long)r.iter;
            cnt.object_tests += (unsigned long long)r.iter * (unsigned long long)(G.n_objects - (rig >= 0 ? 1 : 0));
            cnt.sphere_tests += (unsigned long long)r.iter * (unsigned long long)spheres_tested(G, rig);
        }

        bool frame_done = false;
        if (!shadow_phase) {
            if (GLOW && r.min_dist < mmd) mmd = r.min_dist;
            if (r.final_dist < RAYMARCH_EPS) {
                // hit: first half of shading(), render.rs:1020-1046, then march the shadow ray
                hidx = r.idx;
                pt = r.pos;
                const float4 oa = __ldg(&G.obj_a[hidx]);
                const int4 ob = __ldg(&G.obj_b[hidx]);
                if (ob.x == 0) {
                    n = normalized(pt - mk(oa.x, oa.y, oa.z));
                } else {
                    const float4 n4 = __ldg(&G.obj_n[hidx]);
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
                continue;
            }
            // Miss (render.rs:1385-1393): bg is added and, because pos/eye/ig are unchanged, the very
            // same march repeats until lev reaches MAX_REFLECTIONS. The repeats are bit-identical, so
            // the march result is reused and only the accumulation is replayed (appendix A Q15).
            const V3 bg = bgcolor(P, eye);
            for (;;) {
                if (COUNT) cnt.bg_evals++;
                ret = mk(ret.x + bg.x * fcs.x, ret.y + bg.y * fcs.y, ret.z + bg.z * fcs.z);
                if (MAX_REFLECTIONS_CONST <= lev) break;
                lev += 1;
                if (COUNT) {  // the reference re-marches here
                    cnt.reflect++;
                    cnt.march_steps += (unsigned long long)r.iter;
                    cnt.object_tests += (unsigned long long)r.iter * (unsigned long long)(G.n_objects - (ig >= 0 ? 1 : 0));
                    cnt.sphere_tests += (unsigned long long)r.iter * (unsigned long long)spheres_tested(G, ig);
                }
            }
            frame_done = true;
        } else {
            // ---- second half of shading(), march-mode shadow rule render.rs:1052-1067 ----
            shadow_phase = false;
            const int idx = hidx;
            const float4 oa = __ldg(&G.obj_a[idx]);
            const int4 ob = __ldg(&G.obj_b[idx]);
            const DevMaterial &m = G.mat[ob.z];
            float k1 = 0.2f, k2 = 0.0f;
            const bool lit = FAR_AWAY <= r.travel_dist || MAX_ITER <= r.iter || 0.0f < m.t;
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
                MarchFrame &F = stack[depth];
                F.ret[0] = ret.x; F.ret[1] = ret.y; F.ret[2] = ret.z;
                F.fcs[0] = fcs.x; F.fcs[1] = fcs.y; F.fcs[2] = fcs.z;
                F.A[0] = face.x * omf; F.A[1] = face.y * omf; F.A[2] = face.z * omf;
                F.f = f;
                F.ig = idx;
                F.lev = lev;
                F.mmd = mmd;
                const V3 nf = mk(fcs.x * ks.x, fcs.y * ks.y, fcs.z * ks.z);
                const bool cont = !(idx == 0) && !((nf.x + nf.y + nf.z) <= 0.1f) && !(lev >= MAX_REFLECTIONS_CONST);
                F.cont = cont ? 1 : 0;
                if (cont) {
                    const float en2 = -2.0f * dot(eye, n);
                    const V3 e2 = eye + n * en2;
                    F.pos[0] = pt.x; F.pos[1] = pt.y; F.pos[2] = pt.z;
                    F.eye[0] = e2.x; F.eye[1] = e2.y; F.eye[2] = e2.z;
                }
                depth += 1;
                pos = pt3;
                eye = ray;
                ig = idx;
                ret = mk(0.0f, 0.0f, 0.0f);
                fcs = mk(1.0f, 1.0f, 1.0f);
                mmd = RR_INF;
                ray_class = 1;
                continue;
            }

            // ---- back in raymarch(), render.rs:1346-1384 ----
            ret = mk(ret.x + face.x * fcs.x, ret.y + face.y * fcs.y, ret.z + face.z * fcs.z);
            fcs = mk(fcs.x * ks.x, fcs.y * ks.y, fcs.z * ks.z);
            if (idx == 0 || (fcs.x + fcs.y + fcs.z) <= 0.1f || lev >= MAX_REFLECTIONS_CONST) {
                frame_done = true;
            } else {
                pos = pt;
                const float en2 = -2.0f * dot(eye, n);
                eye = eye + n * en2;  // (flags are recomputed by the reference here but distance() ignores them)
                ig = idx;
                ray_class = 2;
            }
        }

        while (frame_done) {
            const V3 val = apply_glow(P, ret, mmd);  // each raymarch() call applies its own factor
            if (depth == 0) return val;
            depth -= 1;
            const MarchFrame &F = stack[depth];
            const float f = F.f;
            const V3 face = mk(F.A[0] + val.x * f, F.A[1] + val.y * f, F.A[2] + val.z * f);
            ret = mk(F.ret[0] + face.x * F.fcs[0], F.ret[1] + face.y * F.fcs[1], F.ret[2] + face.z * F.fcs[2]);
            mmd = F.mmd;
            if (F.cont) {
                const DevMaterial &pm = G.mat[__ldg(&G.obj_b[F.ig]).z];
                fcs = mk(F.fcs[0] * pm.specular[0], F.fcs[1] * pm.specular[1], F.fcs[2] * pm.specular[2]);
                pos = mk(F.pos[0], F.pos[1], F.pos[2]);
                eye = mk(F.eye[0], F.eye[1], F.eye[2]);
                ig = F.ig;
                lev = F.lev;
                ray_class = 2;
                frame_done = false;
            }
        }
    }
}

}  // namespace rr
