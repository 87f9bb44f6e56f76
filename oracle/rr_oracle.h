/*
 * rr_oracle.h — C entry points of the CPU oracle (liboracle.so).
 *
 * TEST INFRASTRUCTURE ONLY. The oracle is a scalar-f32 CPU restatement of ray-rust's per-pixel
 * path, used as the checker in tests/, __graft_entry__.smoke() and as bench.py's cpu_baseline /
 * --impl reference arm. Nothing in the product path (ray-rust_b200/) links, imports or calls it.
 *
 * It consumes the same POD scene description as the device library (include/rr_ffi.h) so both
 * sides of a parity test see byte-identical inputs.
 */
#ifndef RR_ORACLE_H
#define RR_ORACLE_H

#include "../include/rr_ffi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* per-pixel classification tags (bit mask) written by the instrumented render, used to localise
 * device/oracle mismatches (SURVEY.md 8d "Parity report"). */
#define ORACLE_TAG_BG        (1u << 0)  /* bgproc contributed to the pixel                     */
#define ORACLE_TAG_REFLECT   (1u << 1)  /* a mirror bounce was followed                        */
#define ORACLE_TAG_REFRACT   (1u << 2)  /* a refraction child was traced                       */
#define ORACLE_TAG_SHADOWED  (1u << 3)  /* some shading() call found its point in shadow       */
#define ORACLE_TAG_LIT       (1u << 4)  /* some shading() call found its point lit             */
#define ORACLE_TAG_GRAZING   (1u << 5)  /* a sphere discriminant within a few ulp of EPSILON   */
#define ORACLE_TAG_WRAP      (1u << 6)  /* pattern coordinate within a few ulp of an integer   */
#define ORACLE_TAG_MAXITER   (1u << 7)  /* a march hit the 10 001-iteration cap                */
#define ORACLE_TAG_SUN       (1u << 8)  /* bgcolor sun glare branch (dot > 0.9)                */

/* Test-only uv mapping u=x, v=z: the floor mapping of the older revision that rendered
 * images/example.png. Not part of the reference's UVMap enum; the device library rejects it. */
#define ORACLE_UV_XZ_LEGACY 4

/* render() — src/render.rs:801-900. threads==1: serial row loop (:829-835); otherwise N threads
 * pulling rows from an atomic counter (:836-898). Honours params->band_* like the device path.
 * out_rgb: rows*xres*3 floats or NULL; out_u8: rows*xres*3 bytes or NULL; counts/tags may be NULL
 * (when both are NULL the un-instrumented build of the path runs: this is what gets timed). */
int oracle_render(const rr_scene_desc *desc, const rr_frame_params *params, int threads,
                  float *out_rgb, uint8_t *out_u8, rr_ray_counts *counts, uint32_t *tags);

/* per-function probes (each one calls the same code the render uses) */
float oracle_fmod(float f, float freq);                       /* modutil.rs:1-3   */
int32_t oracle_imod(int32_t f, int32_t freq);                 /* modutil.rs:4-6   */
uint32_t oracle_umod(uint32_t f, uint32_t freq);              /* modutil.rs:7-9   */
void oracle_fimod(float f, float freq, float *frac, uint32_t *i); /* modutil.rs:10-14 */
void oracle_scale_pixel(float s, const uint8_t a[3], float out[3]);     /* pixelutil.rs:11-13 */
void oracle_add_pixel(const float a[3], const float b[3], float out[3]); /* pixelutil.rs:4-10 */
float oracle_powi(float a, int32_t b);                        /* compiler-builtins __powisf2 */
uint8_t oracle_quantize(float c);                             /* main.rs:149 */
void oracle_quat_from_pyr(const float pyr[3], float out_xyzw[4]);       /* quat.rs:129-134 */
void oracle_quat_mul(const float a[4], const float b[4], float out[4]); /* quat.rs:63-72 */
void oracle_quat_transform(const float q[4], const float v[3], float out[3]); /* quat.rs:74-80 */
void oracle_quat_slerp(const float a[4], const float b[4], float t, float out[4]); /* quat.rs:97-127 */
void oracle_normalize(const float v[3], float out[3]);        /* vec3.rs:36-39 */
void oracle_primary_ray(const rr_frame_params *p, int32_t ix, int32_t iy, float eye[3]); /* render.rs:808-815 */
float oracle_sphere_raycast(const float org[3], float r, const float vi[3], const float eye[3],
                            float ray_length, uint32_t flags);          /* render.rs:447-471 */
float oracle_floor_raycast(const float org[3], const float n[3], const float vi[3],
                           const float eye[3], float ray_length);       /* render.rs:557-569 */
float oracle_sphere_distance(const float org[3], float r, const float vi[3]);  /* render.rs:473-475 */
float oracle_floor_distance(const float org[3], const float n[3], const float vi[3]); /* render.rs:571-573 */
void oracle_bgcolor(const float light[3], const float dir[3], float out[3]);   /* main.rs:231-260 */
void oracle_get_uv(const rr_material *m, const float pos[3], int32_t uvmap, float uv[2]); /* render.rs:220-233 */
void oracle_lookup_texture(const rr_material *m, const rr_texture *tex_or_null, float u, float v,
                           float out[3]);                               /* render.rs:249-317 */
/* one pixel, full path */
void oracle_trace_pixel(const rr_scene_desc *desc, const rr_frame_params *params, int32_t ix,
                        int32_t iy, float out[3]);

#ifdef __cplusplus
}
#endif
#endif
