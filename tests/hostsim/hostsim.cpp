// hostsim.cpp — the device's per-pixel logic (trace_pixel / march_pixel from ray-rust_b200/csrc/*.cuh) compiled
// for the CPU. TEST INFRASTRUCTURE: lets the no-GPU test-suite compare the kernels' source logic with the oracle,
// and lets a failing pixel be debugged with printf. It re-implements only the host-side flattening of rr_ffi.cu.
#include "cuda_stub.h"
#define cudaError_t int
#include "../../include/rr_ffi.h"
#include "../../ray-rust_b200/csrc/rr_trace.cuh"
#include "../../ray-rust_b200/csrc/rr_march.cuh"
#include "../../ray-rust_b200/csrc/rr_bvh.h"

#include <vector>

using namespace rr;

namespace {
struct Flat {
    std::vector<float4> sph, sph_m, flo_o, flo_n, obj_a, obj_n;
    std::vector<float> sph_glow;
    std::vector<int> sph_oi, flo_oi;
    std::vector<int4> obj_b;
    std::vector<DevMaterial> mats;
    std::vector<DevTexture> tex;
    Bvh bvh;                                  // same builder as the library (csrc/rr_bvh.h)
    std::vector<float4> bsph, bsph_m;
    std::vector<int> bsph_oi;
    DevScene G{};
    SceneHead H{};
};

void flatten(const rr_scene_desc *d, Flat &f) {
    int n_glow = 0;
    for (uint32_t i = 0; i < d->n_objects; ++i) {
        const rr_object &o = d->objects[i];
        const rr_material &m = d->materials[o.material];
        if (m.glow_dist != 0.0f) n_glow++;
        f.obj_a.push_back(make_float4(o.org[0], o.org[1], o.org[2], o.r));
        f.obj_n.push_back(make_float4(o.face_normal[0], o.face_normal[1], o.face_normal[2], 0.0f));
        f.obj_b.push_back(make_int4(o.kind, o.uvmap, o.material, 0));
        if (o.kind == RR_SPHERE) {
            f.sph.push_back(make_float4(o.org[0], o.org[1], o.org[2], o.r * o.r));
            f.sph_m.push_back(make_float4(o.org[0], o.org[1], o.org[2], o.r));
            f.sph_glow.push_back(m.glow_dist);
            f.sph_oi.push_back((int)i);
        } else {
            f.flo_o.push_back(make_float4(o.org[0], o.org[1], o.org[2], m.glow_dist));
            f.flo_n.push_back(make_float4(o.face_normal[0], o.face_normal[1], o.face_normal[2], 0.0f));
            f.flo_oi.push_back((int)i);
        }
    }
    for (uint32_t i = 0; i < d->n_textures; ++i) f.tex.push_back(DevTexture{d->textures[i].rgb8, d->textures[i].width, d->textures[i].height});
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const rr_material &m = d->materials[i];
        DevMaterial q{};
        for (int k = 0; k < 3; ++k) { q.diffuse[k] = m.diffuse[k]; q.specular[k] = m.specular[k]; }
        q.pn = m.pn; q.t = m.t; q.n = m.n; q.glow_dist = m.glow_dist; q.pattern = m.pattern;
        q.pattern_scale = m.pattern_scale; q.pattern_angle_scale = m.pattern_angle_scale; q.texture = m.texture; q.texture_filter = m.texture_filter;
        f.mats.push_back(q);
    }
    DevScene &G = f.G;
    G.n_spheres = (int)f.sph.size(); G.n_floors = (int)f.flo_o.size(); G.n_objects = (int)d->n_objects; G.n_materials = (int)d->n_materials;
    G.n_glow = n_glow; G.n_bvh_nodes = 0;
    G.sph = f.sph.data(); G.sph_oi = f.sph_oi.data(); G.sph_m = f.sph_m.data(); G.sph_glow = f.sph_glow.data();
    G.flo_o = f.flo_o.data(); G.flo_n = f.flo_n.data(); G.flo_oi = f.flo_oi.data();
    G.obj_a = f.obj_a.data(); G.obj_n = f.obj_n.data(); G.obj_b = f.obj_b.data(); G.mat = f.mats.data(); G.tex = f.tex.data();
    if (build_bvh(f.sph_m, f.bvh)) {          // mirrors rr_scene_create (rr_ffi.cu)
        for (int k : f.bvh.order) { f.bsph.push_back(f.sph[k]); f.bsph_m.push_back(f.sph_m[k]); f.bsph_oi.push_back(f.sph_oi[k]); }
        G.bvh_a = f.bvh.a.data(); G.bvh_b = f.bvh.b.data(); G.bvh_w = f.bvh.w.data();
        G.bsph = f.bsph.data(); G.bsph_m = f.bsph_m.data(); G.bsph_oi = f.bsph_oi.data();
        G.n_bvh_nodes = (int)f.bvh.a.size(); G.n_bvh_inner = (int)(f.bvh.w.size() / 4);
        for (int c = 0; c < 3; ++c) { G.scene_lo[c] = f.bvh.lo[c]; G.scene_hi[c] = f.bvh.hi[c]; }
        G.r_min = f.bvh.r_min;
    }
    SceneHead &H = f.H;
    for (int k = 0; k < RR_HEAD_SPHERES && k < (int)f.sph.size(); ++k) { H.sph[k] = f.sph[k]; H.sph_oi[k] = f.sph_oi[k]; H.sph_m[k] = f.sph_m[k]; H.sph_glow[k] = f.sph_glow[k]; }
    for (int k = 0; k < RR_HEAD_FLOORS && k < (int)f.flo_o.size(); ++k) { H.flo_o[k] = f.flo_o[k]; H.flo_n[k] = f.flo_n[k]; H.flo_oi[k] = f.flo_oi[k]; }
    int ng = 0;
    for (uint32_t i = 0; i < d->n_objects && ng >= 0; ++i) {
        const rr_object &o = d->objects[i];
        const rr_material &m = d->materials[o.material];
        if (m.glow_dist == 0.0f) continue;
        if (ng == RR_HEAD_GLOW) { ng = -1; break; }
        H.glow_a[ng] = make_float4(o.org[0], o.org[1], o.org[2], o.kind == RR_SPHERE ? o.r : 0.0f);
        H.glow_b[ng] = make_float4(o.face_normal[0], o.face_normal[1], o.face_normal[2], 0.0f);
        H.glow_k[ng] = m.glow_dist; H.glow_kind[ng] = o.kind == RR_SPHERE ? 0 : 1; H.glow_oi[ng] = (int)i;
        ++ng;
    }
    H.n_glow_head = ng;
    fill_march_bounds(H, (int)f.sph.size() < RR_HEAD_SPHERES ? (int)f.sph.size() : RR_HEAD_SPHERES);
    fill_head_pairs(H, (int)f.sph.size() < RR_HEAD_SPHERES ? (int)f.sph.size() : RR_HEAD_SPHERES,
                    (int)f.flo_o.size() < RR_HEAD_FLOORS ? (int)f.flo_o.size() : RR_HEAD_FLOORS);
}

std::vector<float4> g_ptab;  // primary-ray tables of the frame being rendered (what rr_ffi.cu's fill kernel writes)
FrameParams to_dev(const rr_frame_params *p, const SceneHead &H) {
    FrameParams d{};
    d.xres = p->xres; d.yres = p->yres; d.xfov = p->xfov; d.yfov = p->yfov;
    for (int k = 0; k < 3; ++k) { d.cam_pos[k] = p->cam_position[k]; d.light[k] = p->light[k]; }
    for (int k = 0; k < 4; ++k) d.cam_rot[k] = p->cam_rotation[k];
    d.use_raymarching = p->use_raymarching; d.glow_enabled = p->glow_enabled; d.glow_effect = p->glow_effect;
    d.max_reflections = p->max_reflections; d.max_refractions = p->max_refractions; d.bg_kind = p->bg_kind;
    d.band_count = 1; d.band_rows = 1; d.band_index = 0; d.band_span = 1; d.local_rows = p->yres; d.row0 = 0; d.placed = 0;
    finish_frame_params(d, H);
    g_ptab.clear();
    for (int ix = 0; ix < d.xres; ++ix) g_ptab.push_back(prim_col_entry(d, ix));
    for (int iy = 0; iy < d.yres; ++iy) g_ptab.push_back(prim_row_entry(d, iy));
    d.ptab = g_ptab.data();
    return d;
}
}  // namespace

// culling: 0 = brute-force scan, 1 = the BVH instance when the builder produces a tree (what the device does by default)
extern "C" int hostsim_render_f32_ex(const rr_scene_desc *desc, const rr_frame_params *params, float *out, rr_ray_counts *counts,
                                     int culling, int *used_bvh) {
    Flat f;
    flatten(desc, f);
    FrameParams P = to_dev(params, f.H);
    Counters cnt{};
    SceneView S{};
    S.sph = f.G.sph; S.sph_oi = f.G.sph_oi; S.flo_o = f.G.flo_o; S.flo_n = f.G.flo_n; S.flo_oi = f.G.flo_oi;
    S.n_spheres = f.G.n_spheres; S.n_floors = f.G.n_floors;
    const bool bvh = (culling & 1) && f.G.n_bvh_nodes > 0;
    S.bvh_a = f.G.bvh_a; S.bvh_b = f.G.bvh_b; S.bvh_w = f.G.bvh_w; S.bsph = f.G.bsph; S.bsph_oi = f.G.bsph_oi;
    S.n_bvh_nodes = bvh ? f.G.n_bvh_nodes : 0;
    if (used_bvh) *used_bvh = bvh ? 1 : 0;
    MarchView M{};
    M.sph = f.G.sph_m; M.sph_glow = f.G.sph_glow; M.sph_oi = f.G.sph_oi; M.flo_o = f.G.flo_o; M.flo_n = f.G.flo_n; M.flo_oi = f.G.flo_oi;
    M.n_spheres = f.G.n_spheres; M.n_floors = f.G.n_floors;
    const int glow = !(P.glow_enabled && f.G.n_glow > 0) ? 0 : (f.H.n_glow_head >= 0 ? 1 : 2);
    const bool mbvh = P.use_raymarching && (culling & 1) && f.G.n_bvh_nodes > 0 && glow != 2;  // as launch_two (rr_march.cu)
    M.bvh_a = f.G.bvh_a; M.bvh_b = f.G.bvh_b; M.bsph = f.G.bsph_m; M.bsph_oi = f.G.bsph_oi; M.n_bvh_nodes = f.G.n_bvh_nodes;
    M.scene_abs = 0.0f;
    for (int k = 0; k < 3; ++k) M.scene_abs = fmaxf(M.scene_abs, fmaxf(fabsf(f.G.scene_lo[k]), fabsf(f.G.scene_hi[k])));
    if (used_bvh && P.use_raymarching) *used_bvh = mbvh ? 1 : 0;
    for (int iy = 0; iy < P.yres; ++iy)
        for (int ix = 0; ix < P.xres; ++ix) {
            V3 c;
            // the device picks the head-only instance for scenes that fit the SceneHead (rr_trace.cu launch_tw); `culling & 2`
            // forces the general instance so that both are compared with the oracle
            const bool headonly = !bvh && !(culling & 2) && f.G.n_floors <= RR_HEAD_FLOORS && f.G.n_spheres <= RR_HEAD_SPHERES;
            if (!P.use_raymarching) c = bvh ? trace_pixel<true, true>(f.G, f.H, S, P, ix, iy, cnt)
                                      : headonly ? trace_pixel<true, false, true>(f.G, f.H, S, P, ix, iy, cnt)
                                                 : trace_pixel<true, false>(f.G, f.H, S, P, ix, iy, cnt);
            else if (mbvh && glow == 0) c = march_pixel<true, 0, true>(f.G, f.H, M, P, ix, iy, cnt);
            else if (mbvh) c = march_pixel<true, 1, true>(f.G, f.H, M, P, ix, iy, cnt);
            else if (glow == 0) c = march_pixel<true, 0>(f.G, f.H, M, P, ix, iy, cnt);
            else if (glow == 1) c = march_pixel<true, 1>(f.G, f.H, M, P, ix, iy, cnt);
            else c = march_pixel<true, 2>(f.G, f.H, M, P, ix, iy, cnt);
            float *o = out + ((size_t)iy * P.xres + ix) * 3;
            o[0] = c.x; o[1] = c.y; o[2] = c.z;
        }
    if (counts) {
        counts->pixels = cnt.pixels; counts->primary = cnt.primary; counts->reflect = cnt.reflect; counts->refract = cnt.refract;
        counts->shadow = cnt.shadow; counts->object_tests = cnt.object_tests; counts->march_steps = cnt.march_steps;
        counts->bg_evals = cnt.bg_evals; counts->sphere_tests = cnt.sphere_tests; counts->sphere_hits = cnt.sphere_hits;
    }
    return 0;
}

extern "C" int hostsim_render_f32(const rr_scene_desc *desc, const rr_frame_params *params, float *out, rr_ray_counts *counts) {
    return hostsim_render_f32_ex(desc, params, out, counts, 0, nullptr);
}

// fmod_2pi() (rr_device.cuh) against fmodf over every float in [lo, hi): returns the number of mismatching arguments
extern "C" long long hostsim_fmod_2pi_mismatches(float lo, float hi, float *first_bad) {
    long long bad = 0;
    const float y = 2.0f * PI_F;
    for (float x = lo; x < hi; x = nextafterf(x, INFINITY)) {
        const float a = fmod_2pi(x), b = fmodf(x, y);
        if (memcmp(&a, &b, 4) != 0) {
            if (bad == 0 && first_bad) *first_bad = x;
            ++bad;
        }
    }
    return bad;
}

// the ray-march kernel's tile-row rotation (FrameParams::march_tile_rot, a host-side scheduling hint)
extern "C" int hostsim_march_tile_rot(const rr_scene_desc *desc, const rr_frame_params *params) {
    Flat f;
    flatten(desc, f);
    return to_dev(params, f.H).march_tile_rot;
}
