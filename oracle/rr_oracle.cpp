// rr_oracle.cpp — CPU oracle for ray-rust's per-pixel tracing path.
//
// TEST INFRASTRUCTURE ONLY (checker for tests/, smoke(), and the CPU baseline of bench.py).
// The product path (ray-rust_b200/) never links or calls this file.
//
// What it is: a scalar IEEE-f32 restatement of the reference algorithm, function by function,
// in the reference's operation order (Rust never contracts a*b+c into an FMA and never
// re-associates, so this file must be compiled with -ffp-contract=off and without -ffast-math;
// see oracle/Makefile). Each function cites the reference lines it follows
// (paths relative to /root/reference/src).
//
// Pinning status (see DESIGN.md "Oracle"):
//   pinned by reference-authored vectors (tests/test_oracle_golden.py):
//     - fmod/imod/umod/fimod (modutil.rs:16-55) and add_pixel/scale_pixel (pixelutil.rs:15-46);
//     - images/example.png, the only rendered artefact the reference ships. It was produced by an
//       older scene (one mirror sphere, floor mapped u=x v=z). Rendering that scene with this
//       oracle reproduces the image outside the glass sphere: primary rays, Quat::transform,
//       bgcolor, the quantiser, sphere and floor intersection, Phong shading, shadow rays and
//       mirror reflection are pinned by it (>=99.6 % of those pixels bit-exact, >=99.9 % within
//       1 LSB; the rest are floor-pattern wrap lines / horizon speckle / the sun-glare core).
//   PARITY UNPINNED for the refraction recursion (the image predates the current refraction
//     code), the ray-marching mode and textures: the reference has no tests or vectors for them
//     and no Rust toolchain exists here to run it.
//
// Third-party arithmetic the reference reaches through Rust std (not vendored, version not pinned
// by Cargo.lock): sqrt/floor (SSE), and glibc's atan2f/asinf/powf/fmodf/sinf/cosf/acosf. This
// file calls the same glibc functions (2.39 in this image). powi follows compiler-builtins'
// __powisf2 (repeated squaring).

#include "rr_oracle.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr float F32_EPSILON = 1.1920929e-7f;  // std::f32::EPSILON
constexpr float F32_INF = std::numeric_limits<float>::infinity();
constexpr float PI_F = 3.14159265358979323846264338327950288f;  // std::f32::consts::PI

// render.rs:11-18
constexpr int32_t MAX_REFLECTIONS = 3;
constexpr uint32_t OUTONLY = 1;
constexpr uint32_t INONLY = 1 << 1;
constexpr uint32_t RIGNORE = 1 << 2;
constexpr uint32_t GIGNORE = 1 << 3;
constexpr uint32_t BIGNORE = 1 << 4;

// render.rs:1253-1255
constexpr float RAYMARCH_EPS = 1e-3f;
constexpr float FAR_AWAY = 1e4f;
constexpr size_t MAX_ITER = 10000;

// ---- Rust scalar semantics ------------------------------------------------------------------
// `x as i32` from f32: saturating, NaN -> 0.
inline int32_t f32_as_i32(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}
// `x as u32` from f32: saturating, NaN -> 0.
inline uint32_t f32_as_u32(float x) {
    if (x != x) return 0;
    if (x >= 4294967296.0f) return UINT32_MAX;
    if (x <= 0.0f) return 0;
    return (uint32_t)x;
}
// `x as u8` from f32: saturating, NaN -> 0.
inline uint8_t f32_as_u8(float x) {
    if (x != x) return 0;
    if (x >= 255.0f) return 255;
    if (x <= 0.0f) return 0;
    return (uint8_t)x;
}
// f32::min / f32::max: if one argument is NaN the other is returned (== fminf/fmaxf).
inline float rs_min(float a, float b) { return fminf(a, b); }
inline float rs_max(float a, float b) { return fmaxf(a, b); }

// f32::powi -> llvm.powi.f32 -> compiler-builtins __powisf2 (repeated squaring).
inline float rs_powi(float a, int32_t b) {
    const bool recip = b < 0;
    uint32_t pow = b < 0 ? (uint32_t)(-(int64_t)b) : (uint32_t)b;
    float mul = 1.0f;
    for (;;) {
        if (pow & 1) mul *= a;
        pow >>= 1;
        if (pow == 0) break;
        a *= a;
    }
    return recip ? 1.0f / mul : mul;
}

// ---- vec3.rs --------------------------------------------------------------------------------
struct Vec3 {
    float x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    explicit Vec3(const float *p) : x(p[0]), y(p[1]), z(p[2]) {}
    float dot(const Vec3 &b) const { return x * b.x + y * b.y + z * b.z; }      // vec3.rs:24-26
    float squared_len() const { return x * x + y * y + z * z; }                  // vec3.rs:28-30
    float len() const { return sqrtf(squared_len()); }                           // vec3.rs:32-34
    Vec3 normalized() const {                                                    // vec3.rs:36-39
        float l = len();
        return Vec3(x / l, y / l, z / l);
    }
};
inline Vec3 operator*(const Vec3 &a, float o) { return Vec3(a.x * o, a.y * o, a.z * o); }  // vec3.rs:60-75
inline Vec3 operator+(const Vec3 &a, const Vec3 &o) { return Vec3(a.x + o.x, a.y + o.y, a.z + o.z); }  // :79-85
inline Vec3 operator-(const Vec3 &a, const Vec3 &o) { return Vec3(a.x - o.x, a.y - o.y, a.z - o.z); }  // :95-108

// ---- quat.rs --------------------------------------------------------------------------------
struct Quat {
    float x, y, z, w;
    Quat() : x(0), y(0), z(0), w(0) {}
    Quat(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    Quat conjugated() const { return Quat(-x, -y, -z, w); }  // quat.rs:59-61
    Quat mul(const Quat &qb) const {                          // quat.rs:63-72
        const Quat &qa = *this;
        return Quat(qa.y * qb.z - qa.z * qb.y + qa.x * qb.w + qa.w * qb.x,
                    qa.z * qb.x - qa.x * qb.z + qa.y * qb.w + qa.w * qb.y,
                    qa.x * qb.y - qa.y * qb.x + qa.z * qb.w + qa.w * qb.z,
                    -qa.x * qb.x - qa.y * qb.y - qa.z * qb.z + qa.w * qb.w);
    }
    Vec3 transform(const Vec3 &v) const {                     // quat.rs:74-80
        Quat qc = conjugated();
        Quat q(v.x, v.y, v.z, 0.0f);                          // From<Vec3>, quat.rs:181-185
        Quat qr = mul(q);
        Quat qret = qr.mul(qc);
        return Vec3(qret.x, qret.y, qret.z);
    }
    float dot(const Quat &b) const { return x * b.x + y * b.y + z * b.z + w * b.w; }  // quat.rs:26-28
    static Quat rotation(float p, float sx, float sy, float sz) {  // quat.rs:92-95
        float len = sinf(p / 2.0f);
        return Quat(len * sx, len * sy, len * sz, cosf(p / 2.0f));
    }
    static Quat from_pyr(const Vec3 &pyr) {                   // quat.rs:129-134
        Quat mx = rotation(pyr.z, 1.0f, 0.0f, 0.0f);
        Quat my = rotation(pyr.y, 0.0f, 0.0f, 1.0f);
        Quat mp = rotation(pyr.x, 0.0f, 1.0f, 0.0f);
        return mx.mul(my).mul(mp);
    }
    bool equals(const Quat &o) const { return x == o.x && y == o.y && z == o.z && w == o.w; }
    Quat slerp(const Quat &o, float t) const {                // quat.rs:97-127
        float qr = dot(o);
        float ss = 1.0f - qr * qr;
        if (ss <= sqrtf(1e-10f) || equals(o)) return *this;
        float sp = sqrtf(ss);
        float ph = acosf(qr);
        float pt = ph * t;
        float t1 = sinf(pt) / sp;
        float t0 = sinf(ph - pt) / sp;
        if (qr < 0.0f) t1 *= -1.0f;
        return Quat(x * t0 + o.x * t1, y * t0 + o.y * t1, z * t0 + o.z * t1, w * t0 + o.w * t1);
    }
};

// ---- modutil.rs / pixelutil.rs --------------------------------------------------------------
inline float m_fmod(float f, float freq) { return f - floorf(f / freq) * freq; }  // modutil.rs:1-3
inline int32_t m_imod(int32_t f, int32_t freq) {                                  // modutil.rs:4-6
    // wrapping arithmetic (release-mode Rust)
    int32_t k = f32_as_i32(floorf((float)f / (float)freq));
    return (int32_t)((uint32_t)f - (uint32_t)k * (uint32_t)freq);
}
inline uint32_t m_umod(uint32_t f, uint32_t freq) {                               // modutil.rs:7-9
    uint32_t k = f32_as_u32(floorf((float)f / (float)freq));
    return f - k * freq;
}
inline void m_fimod(float f, float freq, float *frac, uint32_t *idx) {            // modutil.rs:10-14
    float fm = m_fmod(f, freq);
    float fi = floorf(fm);
    *frac = fm - fi;
    *idx = (uint32_t)m_imod(f32_as_i32(fm), f32_as_i32(freq));
}

struct Color {
    float r, g, b;
    Color() : r(0), g(0), b(0) {}
    Color(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {}
};

// ---- scene model (render.rs:82-97, 378-384, 487-493, 646-666) ---------------------------------
struct Env {
    const rr_scene_desc *d;
    const rr_frame_params *p;
    Vec3 cam_pos;
    Quat cam_rot;
    Vec3 light;
    bool use_raymarching;
    bool glow_enabled;
    float glow_effect;
    int32_t max_reflections, max_refractions;
    int32_t xres, yres;
    float xfov, yfov;
    int32_t bg_kind;
    const rr_object &obj(size_t i) const { return d->objects[i]; }
    const rr_material &mat_of(size_t i) const { return d->materials[d->objects[i].material]; }
};

struct Instr {  // per-thread instrumentation (only touched when COUNT)
    rr_ray_counts c;
    uint32_t tag;
    Instr() : tag(0) { std::memset(&c, 0, sizeof(c)); }
};

inline bool near_ulps(float a, float b, float ulps) {
    float m = fmaxf(fabsf(a), fabsf(b));
    return fabsf(a - b) <= ulps * 1.1920929e-7f * m;
}

// RenderMaterial::get_uv — render.rs:220-233
inline void get_uv(const rr_material &m, const Vec3 &pos, int32_t uvmap, float *u, float *v) {
    switch (uvmap) {
        case RR_UV_XY: *u = pos.x / m.pattern_scale; *v = pos.y / m.pattern_scale; break;
        case RR_UV_YZ: *u = pos.y / m.pattern_scale; *v = pos.z / m.pattern_scale; break;
        case RR_UV_ZX: *u = pos.z / m.pattern_scale; *v = pos.x / m.pattern_scale; break;
        // Test-only mapping (not in the reference's UVMap enum, rejected by the device library): the
        // floor mapping of the older revision that produced images/example.png. Lets the oracle be
        // checked against that image beyond its sky region (tests/test_oracle_golden.py).
        case ORACLE_UV_XZ_LEGACY: *u = pos.x / m.pattern_scale; *v = pos.z / m.pattern_scale; break;
        default: {  // LL
            float dx = pos.x, dz = pos.z;
            *u = atan2f(pos.z, pos.x) / m.pattern_angle_scale;
            *v = atan2f(sqrtf(dx * dx + dz * dz), pos.y) / m.pattern_angle_scale;
        }
    }
}

inline const uint8_t *tex_pixel(const rr_texture &t, uint32_t x, uint32_t y) {
    // image::ImageBuffer::get_pixel panics when out of bounds; the reference only gets there
    // through imod/umod round-off at |coordinate| > 2^24. The oracle clamps instead of aborting.
    if (x >= t.width) x = t.width - 1;
    if (y >= t.height) y = t.height - 1;
    return t.rgb8 + ((size_t)y * t.width + x) * 3;
}

// RenderMaterialInterface::lookup_texture — render.rs:249-317
template <bool COUNT>
inline Color lookup_texture(const rr_material &m, const rr_texture *tex, float u, float v, Instr *ins) {
    if (tex != nullptr && tex->rgb8 != nullptr && tex->width > 0 && tex->height > 0) {
        const float W = (float)tex->width, H = (float)tex->height;
        if (m.texture_filter == RR_NEAREST) {  // render.rs:253-266
            uint32_t px = (uint32_t)m_imod(f32_as_i32(u * W), (int32_t)tex->width);
            uint32_t py = (uint32_t)m_imod(f32_as_i32(v * H), (int32_t)tex->height);
            const uint8_t *p = tex_pixel(*tex, px, py);
            return Color((float)p[0] / 256.0f, (float)p[1] / 256.0f, (float)p[2] / 256.0f);
        } else {  // Bilinear, render.rs:267-296
            float fu, fv;
            uint32_t iu, iv;
            m_fimod(u * W, W, &fu, &iu);
            m_fimod(v * H, H, &fv, &iv);
            const float w[4] = {(1.0f - fu) * (1.0f - fv), (1.0f - fu) * fv, fu * (1.0f - fv), fu * fv};
            const uint8_t *p[4] = {
                tex_pixel(*tex, iu, iv),
                tex_pixel(*tex, iu, m_umod(iv + 1, tex->height)),
                tex_pixel(*tex, m_umod(iu + 1, tex->width), iv),
                tex_pixel(*tex, m_umod(iu + 1, tex->width), m_umod(iv + 1, tex->height)),
            };
            float acc[3] = {0.0f, 0.0f, 0.0f};  // fold(zero, add_pixel)
            for (int k = 0; k < 4; ++k)
                for (int c = 0; c < 3; ++c) acc[c] = acc[c] + w[k] * (float)p[k][c];  // scale_pixel then add_pixel
            return Color(acc[0] / 256.0f, acc[1] / 256.0f, acc[2] / 256.0f);
        }
    }
    switch (m.pattern) {  // render.rs:299-315
        case RR_SOLID: return Color(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
        case RR_CHECKERBOARD: {
            float fu = floorf(u), fv = floorf(v);
            if (COUNT && (near_ulps(u, fu, 8) || near_ulps(u, fu + 1.0f, 8) || near_ulps(v, fv, 8) ||
                          near_ulps(v, fv + 1.0f, 8)))
                ins->tag |= ORACLE_TAG_WRAP;
            int32_t ix = f32_as_i32(fu);
            int32_t iy = f32_as_i32(fv);
            int32_t s = (int32_t)((uint32_t)ix + (uint32_t)iy);
            if (s % 2 == 0) return Color(0.0f, 0.0f, 0.0f);
            return Color(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
        }
        default: {  // RepeatedGradation
            float mu = m_fmod(u, 1.0f), mv = m_fmod(v, 1.0f);
            if (COUNT && (mu < 1e-3f || mu > 0.999f || mv < 1e-3f || mv > 0.999f)) ins->tag |= ORACLE_TAG_WRAP;
            return Color(m.diffuse[0] * mu, m.diffuse[1] * mv, m.diffuse[2]);
        }
    }
}

inline const rr_texture *tex_of(const Env &ren, const rr_material &m) {
    if (m.texture < 0 || (uint32_t)m.texture >= ren.d->n_textures) return nullptr;
    return &ren.d->textures[m.texture];
}

// RenderSphere::raycast — render.rs:447-471
template <bool COUNT>
inline float sphere_raycast(const rr_object &o, const Vec3 &vi, const Vec3 &eye, float ray_length,
                            uint32_t flags, Instr *ins) {
    Vec3 wpt = vi - Vec3(o.org);
    float b = 2.0f * eye.dot(wpt);
    float c = wpt.dot(wpt) - o.r * o.r;
    float d2 = b * b - 4.0f * c;
    if (COUNT && fabsf(d2 - F32_EPSILON) <= 64.0f * F32_EPSILON * fmaxf(b * b, fabsf(4.0f * c)))
        ins->tag |= ORACLE_TAG_GRAZING;
    if (d2 >= F32_EPSILON) {
        float d = sqrtf(d2);
        float t0 = (-b - d) / 2.0f;
        if (0 == (flags & OUTONLY) && t0 >= 0.0f && t0 < ray_length) {
            return t0;
        } else if (0 == (flags & INONLY) && 0.0f < (t0 + d) && t0 + d < ray_length) {
            return t0 + d;
        }
    }
    return ray_length;
}

// RenderFloor::raycast — render.rs:557-569
inline float floor_raycast(const rr_object &o, const Vec3 &vi, const Vec3 &eye, float ray_length) {
    Vec3 wpt = vi - Vec3(o.org);
    Vec3 n(o.face_normal);
    float w = n.dot(eye);
    if (w <= 0.0f) {
        float t0 = (-n.dot(wpt)) / w;
        if (t0 >= 0.0f && t0 < ray_length) return t0;
    }
    return ray_length;
}

// distance — render.rs:473-475 (sphere), :571-573 (floor)
inline float sphere_distance(const rr_object &o, const Vec3 &vi) {
    return rs_max((Vec3(o.org) - vi).len() - o.r, 0.0f);
}
inline float floor_distance(const rr_object &o, const Vec3 &vi) {
    return rs_max((vi - Vec3(o.org)).dot(Vec3(o.face_normal)), 0.0f);
}

// get_normal — render.rs:443-445 (sphere), :553-555 (floor)
inline Vec3 get_normal(const rr_object &o, const Vec3 &pt) {
    if (o.kind == RR_SPHERE) return (pt - Vec3(o.org)).normalized();
    return Vec3(o.face_normal);
}

// get_diffuse — render.rs:434-437, :544-547
template <bool COUNT>
inline Color get_diffuse(const Env &ren, size_t idx, const Vec3 &pt, Instr *ins) {
    const rr_object &o = ren.obj(idx);
    const rr_material &m = ren.mat_of(idx);
    float u, v;
    get_uv(m, pt - Vec3(o.org), o.uvmap, &u, &v);
    return lookup_texture<COUNT>(m, tex_of(ren, m), u, v, ins);
}

// bgcolor — main.rs:231-260
template <bool COUNT>
inline Color bgcolor(const Env &ren, const Vec3 &direction, Instr *ins) {
    if (COUNT) { ins->c.bg_evals++; ins->tag |= ORACLE_TAG_BG; }
    if (ren.bg_kind == RR_BG_BLACK) return Color(0.0f, 0.0f, 0.0f);
    const float PI = PI_F;
    float phi = atan2f(direction.z, direction.x);
    float the = asinf(direction.y);
    float d = fmodf(50.0f * PI + phi * 10.0f * PI, 2.0f * PI) - PI;
    float dd = fmodf(50.0f * PI + the * 10.0f * PI, 2.0f * PI) - PI;
    Color ret(0.5f / (15.0f * (d * d * dd * dd) + 1.0f), 0.25f - direction.y / 4.0f,
              0.25f - direction.y / 4.0f);
    float dot = ren.light.dot(direction);
    if (dot > 0.9f) {
        if (COUNT) ins->tag |= ORACLE_TAG_SUN;
        if (0.9995f < dot) return Color(2.0f, 2.0f, 2.0f);
        Color ret2 = ret;
        if (0.995f < dot) {
            float dd2 = (dot - 0.995f) * 150.0f;
            ret2 = Color(ret.r + dd2, ret.g + dd2, ret.b + dd2);
        }
        float dot2 = dot - 0.9f;
        return Color(ret2.r + dot2 * 5.0f, ret2.g + dot2 * 5.0f, ret2.b);
    }
    return ret;
}

// scene-level raycast — render.rs:993-1018. `ig` is an object index or -1 (the reference compares
// element addresses inside ren.objects, i.e. indices).
template <bool COUNT>
inline float raycast(const Env &ren, const Vec3 &vi, const Vec3 &eye, int64_t ig, uint32_t flags,
                     size_t *ret_idx_out, Instr *ins) {
    float t = F32_INF;
    size_t ret_idx = 0;
    const size_t n = ren.d->n_objects;
    for (size_t idx = 0; idx < n; ++idx) {
        if (ig >= 0 && (size_t)ig == idx) continue;
        const rr_object &o = ren.obj(idx);
        if (COUNT) { ins->c.object_tests++; if (o.kind == RR_SPHERE) ins->c.sphere_tests++; }
        float obj_t = (o.kind == RR_SPHERE) ? sphere_raycast<COUNT>(o, vi, eye, t, flags, ins)
                                            : floor_raycast(o, vi, eye, t);
        if (obj_t < t) {
            t = obj_t;
            ret_idx = idx;
        }
    }
    *ret_idx_out = ret_idx;
    return t;
}

// distance_estimate — render.rs:1226-1251
template <bool COUNT>
inline void distance_estimate(const Env &ren, const Vec3 &vi, int64_t ig, float *closest, size_t *ret_idx,
                              float *glowing, Instr *ins) {
    float closest_dist = F32_INF;
    size_t idx_out = 0;
    float glowing_dist = F32_INF;
    const size_t n = ren.d->n_objects;
    for (size_t idx = 0; idx < n; ++idx) {
        if (ig >= 0 && (size_t)ig == idx) continue;
        const rr_object &o = ren.obj(idx);
        if (COUNT) { ins->c.object_tests++; if (o.kind == RR_SPHERE) ins->c.sphere_tests++; }
        float dist = (o.kind == RR_SPHERE) ? sphere_distance(o, vi) : floor_distance(o, vi);
        if (dist < closest_dist) {
            closest_dist = dist;
            idx_out = idx;
        }
        float glow = dist * ren.mat_of(idx).glow_dist;
        if (0.0f < glow && glow < glowing_dist) glowing_dist = glow;
    }
    *closest = closest_dist;
    *ret_idx = idx_out;
    *glowing = glowing_dist;
}

struct MarchResult {  // render.rs:1257-1264
    float final_dist;
    size_t idx;
    Vec3 pos;
    size_t iter;
    float travel_dist;
    float min_dist;
};

// raymarch_single — render.rs:1266-1297
template <bool COUNT>
inline MarchResult raymarch_single(const Env &ren, const Vec3 &init_pos, const Vec3 &eye, int64_t ig, Instr *ins) {
    size_t iter = 0;
    float travel_dist = 0.0f;
    Vec3 pos = init_pos;
    float min_dist = F32_INF;
    for (;;) {
        float dist, glowing_dist;
        size_t idx;
        distance_estimate<COUNT>(ren, pos, ig, &dist, &idx, &glowing_dist, ins);
        pos = (eye * dist) + pos;
        travel_dist += dist;
        iter += 1;
        if (COUNT) ins->c.march_steps++;
        if (glowing_dist < min_dist) min_dist = glowing_dist;
        if (dist < RAYMARCH_EPS || FAR_AWAY < dist || MAX_ITER < iter) {
            if (COUNT && MAX_ITER < iter) ins->tag |= ORACLE_TAG_MAXITER;
            return MarchResult{dist, idx, pos, iter, travel_dist, min_dist};
        }
    }
}

enum RayClass { RC_PRIMARY = 0, RC_REFRACT = 1 };

template <bool COUNT>
Color raytrace(const Env &ren, Vec3 vi, Vec3 eye, int32_t lev, int64_t init_ig, uint32_t flags, int rc, Instr *ins);
template <bool COUNT>
Color raymarch(const Env &ren, Vec3 vi, Vec3 eye, int32_t lev, int64_t init_ig, uint32_t flags, int rc, Instr *ins);

// shading — render.rs:1020-1140
template <bool COUNT>
Color shading(const Env &ren, size_t idx, const Vec3 &n, const Vec3 &pt, const Vec3 &eye, int32_t nest, Instr *ins) {
    const rr_material &mat = ren.mat_of(idx);

    // render.rs:1024-1046
    float light_incidence = ren.light.dot(n);
    float ln2 = 2.0f * light_incidence;
    Vec3 reflected_ray_to_light_source = (n * ln2) - ren.light;
    const float eps = F32_EPSILON;
    const int32_t pn = mat.pn;
    float diffuse_intensity = rs_max(light_incidence, 0.0f);
    Vec3 reflected_ray = pt + (ren.light * eps);
    float reflection_intensity;
    if (0 != pn) {
        float reflection_incidence = -reflected_ray_to_light_source.dot(eye);
        reflection_intensity = reflection_incidence > 0.0f ? rs_powi(reflection_incidence, pn) : 0.0f;
    } else {
        reflection_intensity = 0.0f;
    }

    // shadow trace, render.rs:1048-1082
    float k1, k2;
    {
        const Vec3 ray = ren.light;
        k1 = 0.2f;
        bool lit;
        if (COUNT) { ins->c.shadow++; if (ren.obj(idx).kind == RR_SPHERE) ins->c.sphere_hits++; }
        if (ren.use_raymarching) {
            MarchResult r = raymarch_single<COUNT>(ren, reflected_ray, ray, (int64_t)idx, ins);
            lit = FAR_AWAY <= r.travel_dist || MAX_ITER <= r.iter || 0.0f < ren.mat_of(idx).t;
        } else {
            size_t i;
            float t = raycast<COUNT>(ren, reflected_ray, ray, (int64_t)idx, 0, &i, ins);
            lit = t >= F32_INF || 0.0f < ren.mat_of(i).t;
        }
        if (lit) {
            k1 = rs_min(k1 + diffuse_intensity, 1.0f);
            k2 = reflection_intensity;
        } else {
            k2 = 0.0f;
        }
        if (COUNT) ins->tag |= lit ? ORACLE_TAG_LIT : ORACLE_TAG_SHADOWED;
    }

    // face texturing, render.rs:1084-1085
    Color kd = get_diffuse<COUNT>(ren, idx, pt, ins);

    // refraction, render.rs:1092-1139
    if (nest < ren.max_refractions && 0.0f < mat.t) {
        float sp = eye.dot(n);
        float f = mat.t;
        float frac = mat.n;
        float reference = sp * ((sp > 0.0f ? frac : 1.0f / frac) - 1.0f);
        Vec3 ray = (eye + (n * reference)).normalized();
        Vec3 pt3 = pt + (ray * eps);
        if (COUNT) ins->tag |= ORACLE_TAG_REFRACT;
        uint32_t fl = sp < 0.0f ? OUTONLY : INONLY;
        Color fc2 = ren.use_raymarching ? raymarch<COUNT>(ren, pt3, ray, nest, (int64_t)idx, fl, RC_REFRACT, ins)
                                        : raytrace<COUNT>(ren, pt3, ray, nest, (int64_t)idx, fl, RC_REFRACT, ins);
        return Color((kd.r * k1 + k2) * (1.0f - f) + fc2.r * f, (kd.g * k1 + k2) * (1.0f - f) + fc2.g * f,
                     (kd.b * k1 + k2) * (1.0f - f) + fc2.b * f);
    }
    return Color(kd.r * k1 + k2, kd.g * k1 + k2, kd.b * k1 + k2);
}

// raytrace — render.rs:1142-1224
template <bool COUNT>
Color raytrace(const Env &ren, Vec3 vi, Vec3 eye, int32_t lev, int64_t init_ig, uint32_t flags, int rc, Instr *ins) {
    Color fcs(1.0f, 1.0f, 1.0f);
    Color ret_color(0.0f, 0.0f, 0.0f);
    int64_t ig = init_ig;
    bool first = true;
    for (;;) {
        lev += 1;
        if (COUNT) {
            if (!first) ins->c.reflect++;
            else if (rc == RC_PRIMARY) ins->c.primary++;
            else ins->c.refract++;
            first = false;
        }
        size_t idx;
        float t = raycast<COUNT>(ren, vi, eye, ig, flags, &idx, ins);
        if (t < F32_INF) {
            Vec3 pt = (eye * t) + vi;
            const rr_object &o = ren.obj(idx);
            Vec3 n = get_normal(o, pt);
            Color face_color = shading<COUNT>(ren, idx, n, pt, eye, lev, ins);
            const rr_material &m = ren.mat_of(idx);
            Color ks(m.specular[0], m.specular[1], m.specular[2]);
            if (0 == (RIGNORE & flags)) { ret_color.r += face_color.r * fcs.r; fcs.r *= ks.r; }
            if (0 == (GIGNORE & flags)) { ret_color.g += face_color.g * fcs.g; fcs.g *= ks.g; }
            if (0 == (BIGNORE & flags)) { ret_color.b += face_color.b * fcs.b; fcs.b *= ks.b; }
            if (idx == 0) break;
            if ((fcs.r + fcs.g + fcs.b) <= 0.1f) break;
            if (lev >= ren.max_reflections) break;

            vi = pt;
            float en2 = -2.0f * eye.dot(n);
            eye = eye + n * en2;
            if (n.dot(eye) < 0.0f) {
                flags &= ~INONLY;
                flags |= OUTONLY;
            } else {
                flags &= ~OUTONLY;
                flags |= INONLY;
            }
            ig = (int64_t)idx;
            if (COUNT) ins->tag |= ORACLE_TAG_REFLECT;
        } else {
            Color fc2 = bgcolor<COUNT>(ren, eye, ins);
            ret_color.r += fc2.r * fcs.r;
            ret_color.g += fc2.g * fcs.g;
            ret_color.b += fc2.b * fcs.b;
        }
        if (!(t < F32_INF && lev < ren.max_reflections)) break;
    }
    return ret_color;
}

// raymarch — render.rs:1299-1411
template <bool COUNT>
Color raymarch(const Env &ren, Vec3 vi, Vec3 eye, int32_t lev, int64_t init_ig, uint32_t flags, int rc, Instr *ins) {
    Color fcs(1.0f, 1.0f, 1.0f);
    Vec3 pos = vi;
    Color ret_color(0.0f, 0.0f, 0.0f);
    float min_min_dist = F32_INF;
    int64_t ig = init_ig;
    bool first = true;
    for (;;) {
        lev += 1;
        if (COUNT) {
            if (!first) ins->c.reflect++;
            else if (rc == RC_PRIMARY) ins->c.primary++;
            else ins->c.refract++;
            first = false;
        }
        MarchResult r = raymarch_single<COUNT>(ren, pos, eye, ig, ins);
        if (r.min_dist < min_min_dist) min_min_dist = r.min_dist;
        if (r.final_dist < RAYMARCH_EPS) {
            const size_t idx = r.idx;
            const Vec3 pt = r.pos;
            const rr_object &o = ren.obj(idx);
            Vec3 n = get_normal(o, pt);
            Color face_color = shading<COUNT>(ren, idx, n, pt, eye, lev, ins);
            const rr_material &m = ren.mat_of(idx);
            Color ks(m.specular[0], m.specular[1], m.specular[2]);
            if (0 == (RIGNORE & flags)) { ret_color.r += face_color.r * fcs.r; fcs.r *= ks.r; }
            if (0 == (GIGNORE & flags)) { ret_color.g += face_color.g * fcs.g; fcs.g *= ks.g; }
            if (0 == (BIGNORE & flags)) { ret_color.b += face_color.b * fcs.b; fcs.b *= ks.b; }
            if (idx == 0) break;
            if ((fcs.r + fcs.g + fcs.b) <= 0.1f) break;
            if (lev >= MAX_REFLECTIONS) break;

            pos = pt;
            float en2 = -2.0f * eye.dot(n);
            eye = eye + n * en2;
            if (n.dot(eye) < 0.0f) {
                flags &= ~INONLY;
                flags |= OUTONLY;
            } else {
                flags &= ~OUTONLY;
                flags |= INONLY;
            }
            ig = (int64_t)idx;
            if (COUNT) ins->tag |= ORACLE_TAG_REFLECT;
        } else {
            Color fc2 = bgcolor<COUNT>(ren, eye, ins);
            ret_color.r += fc2.r * fcs.r;
            ret_color.g += fc2.g * fcs.g;
            ret_color.b += fc2.b * fcs.b;
            // no break: the same march repeats until lev reaches MAX_REFLECTIONS (render.rs:1385-1393)
        }
        if (MAX_REFLECTIONS <= lev) break;
    }
    if (ren.glow_enabled) {  // render.rs:1397-1410
        float factor = (min_min_dist == F32_INF) ? 1.0f : 1.0f + (0.0f + ren.glow_effect * powf(0.99f, min_min_dist));
        return Color(factor * ret_color.r, factor * ret_color.g, factor * ret_color.b);
    }
    return ret_color;
}

// primary ray of pixel (ix,iy) — render.rs:808-815
inline Vec3 primary_ray(const Env &ren, int32_t ix, int32_t iy) {
    Vec3 eye(1.0f, (float)(ix - ren.xres / 2) * 2.0f * ren.xfov / (float)ren.xres,
             (float)(-(iy - ren.yres / 2)) * 2.0f * ren.yfov / (float)ren.yres);
    return ren.cam_rot.transform(eye).normalized();
}

template <bool COUNT>
inline Color trace_pixel(const Env &ren, int32_t ix, int32_t iy, Instr *ins) {
    Vec3 vi = ren.cam_pos;
    Vec3 eye = primary_ray(ren, ix, iy);
    if (COUNT) ins->c.pixels++;
    return ren.use_raymarching ? raymarch<COUNT>(ren, vi, eye, 0, -1, 0, RC_PRIMARY, ins)
                               : raytrace<COUNT>(ren, vi, eye, 0, -1, 0, RC_PRIMARY, ins);
}

inline uint8_t quantize(float c) { return f32_as_u8(rs_min(c * 255.0f, 255.0f)); }  // main.rs:149

Env make_env(const rr_scene_desc *d, const rr_frame_params *p) {
    Env e;
    e.d = d;
    e.p = p;
    e.cam_pos = Vec3(p->cam_position);
    e.cam_rot = Quat(p->cam_rotation[0], p->cam_rotation[1], p->cam_rotation[2], p->cam_rotation[3]);
    e.light = Vec3(p->light);
    e.use_raymarching = p->use_raymarching != 0;
    e.glow_enabled = p->glow_enabled != 0;
    e.glow_effect = p->glow_effect;
    e.max_reflections = p->max_reflections;
    e.max_refractions = p->max_refractions;
    e.xres = p->xres;
    e.yres = p->yres;
    e.xfov = p->xfov;
    e.yfov = p->yfov;
    e.bg_kind = p->bg_kind;
    return e;
}

// rows of this shard, in packed order (see rr_frame_params.band_*)
std::vector<int32_t> shard_rows(const rr_frame_params *p) {
    std::vector<int32_t> rows;
    const int32_t cnt = p->band_count <= 1 ? 1 : p->band_count;
    const int32_t br = p->band_rows <= 0 ? 1 : p->band_rows;
    for (int32_t iy = 0; iy < p->yres; ++iy) {
        const int32_t slot = (iy / br) % cnt, span = p->band_span <= 1 ? 1 : p->band_span;
        if (cnt == 1 || (slot >= p->band_index && slot < p->band_index + span)) rows.push_back(iy);
    }
    return rows;
}

template <bool COUNT>
void render_rows(const Env &ren, const std::vector<int32_t> &rows, int threads, float *out_rgb, uint8_t *out_u8,
                 rr_ray_counts *counts, uint32_t *tags) {
    const int32_t W = ren.xres;
    // process_line, render.rs:806-827
    auto process_line = [&](size_t li, Instr *ins) {
        const int32_t iy = rows[li];
        for (int32_t ix = 0; ix < W; ++ix) {
            if (COUNT) ins->tag = 0;
            Color c = trace_pixel<COUNT>(ren, ix, iy, ins);
            const size_t o = (li * (size_t)W + (size_t)ix);
            if (out_rgb) { out_rgb[o * 3 + 0] = c.r; out_rgb[o * 3 + 1] = c.g; out_rgb[o * 3 + 2] = c.b; }
            if (out_u8) { out_u8[o * 3 + 0] = quantize(c.r); out_u8[o * 3 + 1] = quantize(c.g); out_u8[o * 3 + 2] = quantize(c.b); }
            if (COUNT && tags) tags[o] = ins->tag;
        }
    };
    auto merge = [&](const Instr &ins) {
        if (!counts) return;
        counts->pixels += ins.c.pixels; counts->primary += ins.c.primary; counts->reflect += ins.c.reflect;
        counts->refract += ins.c.refract; counts->shadow += ins.c.shadow; counts->object_tests += ins.c.object_tests;
        counts->march_steps += ins.c.march_steps; counts->bg_evals += ins.c.bg_evals;
        counts->sphere_tests += ins.c.sphere_tests; counts->sphere_hits += ins.c.sphere_hits;
    };
    if (threads <= 1) {  // render.rs:829-835
        Instr ins;
        for (size_t li = 0; li < rows.size(); ++li) process_line(li, &ins);
        merge(ins);
    } else {  // render.rs:836-898: N workers pulling rows from an atomic counter
        std::atomic<size_t> counter(0);
        std::vector<Instr> per(threads);
        std::vector<std::thread> th;
        for (int k = 0; k < threads; ++k) {
            th.emplace_back([&, k]() {
                for (;;) {
                    size_t li = counter.fetch_add(1);
                    if (rows.size() <= li) break;
                    process_line(li, &per[k]);
                }
            });
        }
        for (auto &t : th) t.join();
        for (auto &i : per) merge(i);
    }
}

}  // namespace

extern "C" {

int oracle_render(const rr_scene_desc *desc, const rr_frame_params *params, int threads, float *out_rgb,
                  uint8_t *out_u8, rr_ray_counts *counts, uint32_t *tags) {
    if (!desc || !params || params->xres < 0 || params->yres < 0) return RR_ERR_BAD_ARG;
    for (uint32_t i = 0; i < desc->n_objects; ++i)
        if (desc->objects[i].material < 0 || (uint32_t)desc->objects[i].material >= desc->n_materials) return RR_ERR_BAD_ARG;
    Env ren = make_env(desc, params);
    std::vector<int32_t> rows = shard_rows(params);
    if (counts) std::memset(counts, 0, sizeof(*counts));
    if (counts || tags) render_rows<true>(ren, rows, threads, out_rgb, out_u8, counts, tags);
    else render_rows<false>(ren, rows, threads, out_rgb, out_u8, nullptr, nullptr);
    return RR_OK;
}

float oracle_fmod(float f, float freq) { return m_fmod(f, freq); }
int32_t oracle_imod(int32_t f, int32_t freq) { return m_imod(f, freq); }
uint32_t oracle_umod(uint32_t f, uint32_t freq) { return m_umod(f, freq); }
void oracle_fimod(float f, float freq, float *frac, uint32_t *i) { m_fimod(f, freq, frac, i); }
void oracle_scale_pixel(float s, const uint8_t a[3], float out[3]) {
    for (int c = 0; c < 3; ++c) out[c] = s * (float)a[c];
}
void oracle_add_pixel(const float a[3], const float b[3], float out[3]) {
    for (int c = 0; c < 3; ++c) out[c] = a[c] + b[c];
}
float oracle_powi(float a, int32_t b) { return rs_powi(a, b); }
uint8_t oracle_quantize(float c) { return quantize(c); }
void oracle_quat_from_pyr(const float pyr[3], float o[4]) {
    Quat q = Quat::from_pyr(Vec3(pyr));
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}
void oracle_quat_mul(const float a[4], const float b[4], float o[4]) {
    Quat q = Quat(a[0], a[1], a[2], a[3]).mul(Quat(b[0], b[1], b[2], b[3]));
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}
void oracle_quat_transform(const float q[4], const float v[3], float o[3]) {
    Vec3 r = Quat(q[0], q[1], q[2], q[3]).transform(Vec3(v));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void oracle_quat_slerp(const float a[4], const float b[4], float t, float o[4]) {
    Quat q = Quat(a[0], a[1], a[2], a[3]).slerp(Quat(b[0], b[1], b[2], b[3]), t);
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}
void oracle_normalize(const float v[3], float o[3]) {
    Vec3 r = Vec3(v).normalized();
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void oracle_primary_ray(const rr_frame_params *p, int32_t ix, int32_t iy, float eye[3]) {
    rr_scene_desc d{};
    Env ren = make_env(&d, p);
    Vec3 e = primary_ray(ren, ix, iy);
    eye[0] = e.x; eye[1] = e.y; eye[2] = e.z;
}
static rr_object mk_obj(int kind, const float org[3], float r, const float n[3]) {
    rr_object o{};
    o.kind = kind; o.r = r;
    for (int c = 0; c < 3; ++c) { o.org[c] = org[c]; o.face_normal[c] = n ? n[c] : 0.0f; }
    return o;
}
float oracle_sphere_raycast(const float org[3], float r, const float vi[3], const float eye[3], float ray_length,
                            uint32_t flags) {
    rr_object o = mk_obj(RR_SPHERE, org, r, nullptr);
    return sphere_raycast<false>(o, Vec3(vi), Vec3(eye), ray_length, flags, nullptr);
}
float oracle_floor_raycast(const float org[3], const float n[3], const float vi[3], const float eye[3],
                           float ray_length) {
    rr_object o = mk_obj(RR_FLOOR, org, 0.0f, n);
    return floor_raycast(o, Vec3(vi), Vec3(eye), ray_length);
}
float oracle_sphere_distance(const float org[3], float r, const float vi[3]) {
    rr_object o = mk_obj(RR_SPHERE, org, r, nullptr);
    return sphere_distance(o, Vec3(vi));
}
float oracle_floor_distance(const float org[3], const float n[3], const float vi[3]) {
    rr_object o = mk_obj(RR_FLOOR, org, 0.0f, n);
    return floor_distance(o, Vec3(vi));
}
void oracle_bgcolor(const float light[3], const float dir[3], float out[3]) {
    rr_scene_desc d{};
    rr_frame_params p{};
    for (int c = 0; c < 3; ++c) p.light[c] = light[c];
    Env ren = make_env(&d, &p);
    Color c = bgcolor<false>(ren, Vec3(dir), nullptr);
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
void oracle_get_uv(const rr_material *m, const float pos[3], int32_t uvmap, float uv[2]) {
    get_uv(*m, Vec3(pos), uvmap, &uv[0], &uv[1]);
}
void oracle_lookup_texture(const rr_material *m, const rr_texture *tex, float u, float v, float out[3]) {
    Color c = lookup_texture<false>(*m, tex, u, v, nullptr);
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
void oracle_trace_pixel(const rr_scene_desc *desc, const rr_frame_params *params, int32_t ix, int32_t iy,
                        float out[3]) {
    Env ren = make_env(desc, params);
    Color c = trace_pixel<false>(ren, ix, iy, nullptr);
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}

}  // extern "C"
