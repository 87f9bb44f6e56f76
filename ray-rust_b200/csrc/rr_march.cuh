// rr_march.cuh — ray-march mode of the per-pixel path (render.rs:1226-1411) as device functions.
//
// raymarch_single()'s sphere-tracing loop is the hot loop (up to 10 001 dependent iterations, each
// an O(N) distance scan). The scan runs over the same shared-memory SoA lists as the trace kernel;
// the glow distance (render.rs:1244-1247) is only tracked when --gloweffect is set and the scene
// has a glowing material, because nothing else reads it.
#pragma once
#include "rr_device.cuh"

namespace rr {

struct MarchView {
    const float4 *sph;     // (cx, cy, cz, r)
    const float *sph_glow; // glow_dist per sphere
    const int *sph_oi;
    const float4 *flo_o;   // (ox, oy, oz, glow_dist)
    const float4 *flo_n;
    const int *flo_oi;
    int n_spheres, n_floors;
};

struct MarchResult {  // render.rs:1257-1264
    float final_dist;
    int idx;
    V3 pos;
    int iter;
    float travel_dist;
    float min_dist;
};

// distance_estimate, render.rs:1226-1251 (+ distance(): :473-475 sphere, :571-573 floor).
// Lowest original index wins ties, as the reference's in-order strict `<` scan does.
template <bool GLOW>
__device__ __forceinline__ void distance_estimate(const MarchView &S, const V3 &vi, int ig, float &closest, int &idx_out,
                                                  float &glowing) {
    float best = RR_INF, gl = RR_INF;
    int idx = 0;
    for (int f = 0; f < S.n_floors; ++f) {
        const int oi = S.flo_oi[f];
        if (oi == ig) continue;
        const float4 o = S.flo_o[f];
        const float4 n = S.flo_n[f];
        const float dist = fmaxf(dot(vi - mk(o.x, o.y, o.z), mk(n.x, n.y, n.z)), 0.0f);
        if (dist < best) {
            best = dist;
            idx = oi;
        }
        if (GLOW) {
            const float g = dist * o.w;
            if (0.0f < g && g < gl) gl = g;
        }
    }
    for (int s = 0; s < S.n_spheres; ++s) {
        const float4 c = S.sph[s];
        const V3 d = mk(c.x, c.y, c.z) - vi;
        const float dist = fmaxf(sqrtf(d.x * d.x + d.y * d.y + d.z * d.z) - c.w, 0.0f);
        const int oi = S.sph_oi[s];
        if (oi == ig) continue;
        if (dist < best || (dist == best && oi < idx && dist < RR_INF)) {
            best = dist;
            idx = oi;
        }
        if (GLOW) {
            const float g = dist * S.sph_glow[s];
            if (0.0f < g && g < gl) gl = g;
        }
    }
    closest = best;
    idx_out = idx;
    glowing = gl;
}

// raymarch_single, render.rs:1266-1297
template <bool GLOW>
__device__ __forceinline__ MarchResult raymarch_single(const MarchView &S, const V3 &init_pos, const V3 &eye, int ig) {
    int iter = 0;
    float travel = 0.0f;
    V3 pos = init_pos;
    float min_dist = RR_INF;
    for (;;) {
        float dist, gl;
        int idx;
        distance_estimate<GLOW>(S, pos, ig, dist, idx, gl);
        pos = (eye * dist) + pos;
        travel += dist;
        iter += 1;
        if (GLOW && gl < min_dist) min_dist = gl;
        if (dist < RAYMARCH_EPS || FAR_AWAY < dist || MAX_ITER < iter) return MarchResult{dist, idx, pos, iter, travel, min_dist};
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
    unsigned flags;
    int cont;
};

constexpr int RR_MARCH_MAX_STACK = 32;

// glow factor, render.rs:1397-1410. powf is CUDA's (<= 2-4 ulp from glibc's).
__device__ __forceinline__ V3 apply_glow(const FrameParams &P, const V3 &c, float mmd) {
    if (!P.glow_enabled) return c;
    const float factor = (mmd == RR_INF) ? 1.0f : 1.0f + (0.0f + P.glow_effect * powf(0.99f, mmd));
    return mk(factor * c.x, factor * c.y, factor * c.z);
}

template <bool COUNT, bool GLOW>
__device__ __forceinline__ V3 march_pixel(const DevScene &G, const MarchView &S, const FrameParams &P, int ix, int iy,
                                          Counters &cnt) {
    const V3 light = mk(P.light[0], P.light[1], P.light[2]);
    V3 pos = mk(P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]);
    V3 eye = primary_ray(P, ix, iy);
    int lev = 0, ig = -1, depth = 0;
    unsigned flags = 0;
    V3 ret = mk(0.0f, 0.0f, 0.0f), fcs = mk(1.0f, 1.0f, 1.0f);
    float mmd = RR_INF;
    MarchFrame stack[RR_MARCH_MAX_STACK];
    int ray_class = 0;
    if (COUNT) cnt.pixels++;

    for (;;) {
        lev += 1;  // render.rs:1317
        const MarchResult r = raymarch_single<GLOW>(S, pos, eye, ig);
        if (COUNT) {
            if (ray_class == 0) cnt.primary++;
            else if (ray_class == 1) cnt.refract++;
            else cnt.reflect++;
            cnt.march_steps += (unsigned long long)r.iter;
            cnt.object_tests += (unsigned long long)r.iter * (unsigned long long)(G.n_objects - (ig >= 0 ? 1 : 0));
            cnt.sphere_tests += (unsigned long long)r.iter * (unsigned long long)spheres_tested(G, ig);
        }
        if (r.min_dist < mmd) mmd = r.min_dist;
        bool frame_done;
        if (r.final_dist < RAYMARCH_EPS) {
            const int idx = r.idx;
            const V3 pt = r.pos;
            const float4 oa = __ldg(&G.obj_a[idx]);
            const int4 ob = __ldg(&G.obj_b[idx]);
            V3 n;
            if (ob.x == 0) n = normalized(pt - mk(oa.x, oa.y, oa.z));
            else {
                const float4 n4 = __ldg(&G.obj_n[idx]);
                n = mk(n4.x, n4.y, n4.z);
            }
            const DevMaterial &m = G.mat[ob.z];

            // ---- shading(), render.rs:1020-1140, march-mode shadow (:1052-1067) ----
            const float light_incidence = dot(light, n);
            const float ln2 = 2.0f * light_incidence;
            const V3 rr_light = (n * ln2) - light;
            const int pn = m.pn;
            const float diffuse_intensity = fmaxf(light_incidence, 0.0f);
            const V3 shadow_org = pt + (light * F32_EPSILON);
            float reflection_intensity = 0.0f;
            if (pn != 0) {
                const float ri = -dot(rr_light, eye);
                if (ri > 0.0f) reflection_intensity = rs_powi(ri, pn);
            }
            float k1 = 0.2f, k2 = 0.0f;
            {
                const MarchResult sh = raymarch_single<false>(S, shadow_org, light, idx);
                if (COUNT) {
                    cnt.shadow++;
                    cnt.march_steps += (unsigned long long)sh.iter;
                    cnt.object_tests += (unsigned long long)sh.iter * (unsigned long long)(G.n_objects - 1);
                    cnt.sphere_tests += (unsigned long long)sh.iter * (unsigned long long)spheres_tested(G, idx);
                    if (ob.x == 0) cnt.sphere_hits++;
                }
                const bool lit = FAR_AWAY <= sh.travel_dist || MAX_ITER <= sh.iter || 0.0f < m.t;
                if (lit) {
                    k1 = fminf(k1 + diffuse_intensity, 1.0f);
                    k2 = reflection_intensity;
                }
            }
            float u, v;
            get_uv(m, pt - mk(oa.x, oa.y, oa.z), ob.y, u, v);
            const V3 kd = lookup_texture(G, m, u, v);
            V3 face = mk(kd.x * k1 + k2, kd.y * k1 + k2, kd.z * k1 + k2);
            const V3 ks = mk(m.specular[0], m.specular[1], m.specular[2]);

            if (lev < P.max_refractions && 0.0f < m.t) {
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
                    F.flags = dot(n, e2) < 0.0f ? OUTONLY : INONLY;
                }
                depth += 1;
                pos = pt3;
                eye = ray;
                ig = idx;
                flags = sp < 0.0f ? OUTONLY : INONLY;
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
                eye = eye + n * en2;
                flags = dot(n, eye) < 0.0f ? OUTONLY : INONLY;
                ig = idx;
                ray_class = 2;
                frame_done = false;  // lev < MAX_REFLECTIONS here, so render.rs:1391 does not break
            }
        } else {
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
        }
        (void)flags;

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
                flags = F.flags;
                ig = F.ig;
                lev = F.lev;
                ray_class = 2;
                frame_done = false;
            }
        }
    }
}

}  // namespace rr
