/* examples/render_c.c — the drop-in boundary used from plain C: what a cgo / JNI / Rust `-sys` binding does, without any of
 * this repo's host layers. Builds a two-sphere scene over a floor, renders it on device 0 into a pinned host frame through
 * rr_render_rgb8 and writes a binary PPM.
 *
 *   gcc -O2 -I include examples/render_c.c -L ray-rust_b200 -lray_rust_b200 -Wl,-rpath,$PWD/ray-rust_b200 -lm -o /tmp/render_c
 *   /tmp/render_c 640 480 /tmp/out.ppm
 *
 * Exit status: 0 ok, 1 usage, 2 library error (e.g. no usable CUDA device: there is no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rr_ffi.h"

static int check(int rc, const char *what) {
    if (rc == RR_OK) return 0;
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, rr_last_error());
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s <width> <height> <out.ppm>\n", argv[0]); return 1; }
    const int w = atoi(argv[1]), h = atoi(argv[2]);
    if (w <= 0 || h <= 0) { fprintf(stderr, "bad size\n"); return 1; }

    rr_material mats[3];
    memset(mats, 0, sizeof mats);
    /* floor: yellow repeated gradation, like main.rs:156-169 */
    mats[0].diffuse[0] = 1.0f; mats[0].diffuse[1] = 1.0f; mats[0].pattern = RR_REPEATED_GRADATION; mats[0].pattern_scale = 300.0f;
    mats[0].pattern_angle_scale = 0.2f; mats[0].texture = -1;
    /* mirror */
    mats[1].specular[0] = mats[1].specular[1] = mats[1].specular[2] = 1.0f; mats[1].pn = 24; mats[1].pattern_scale = 1.0f; mats[1].texture = -1;
    /* glass */
    mats[2].specular[0] = mats[2].specular[1] = mats[2].specular[2] = 1.0f; mats[2].pn = 24; mats[2].t = 1.0f; mats[2].n = 1.5f;
    mats[2].pattern_scale = 1.0f; mats[2].texture = -1;

    rr_object objs[3];
    memset(objs, 0, sizeof objs);
    objs[0].kind = RR_FLOOR; objs[0].material = 0; objs[0].uvmap = RR_UV_ZX; objs[0].org[1] = -300.0f; objs[0].face_normal[1] = 1.0f;
    objs[1].kind = RR_SPHERE; objs[1].material = 1; objs[1].r = 80.0f; objs[1].org[0] = -120.0f; objs[1].org[1] = -220.0f; objs[1].org[2] = 172.0f;
    objs[2].kind = RR_SPHERE; objs[2].material = 2; objs[2].r = 100.0f; objs[2].org[0] = 90.0f; objs[2].org[1] = -200.0f; objs[2].org[2] = 150.0f;

    rr_scene_desc desc;
    memset(&desc, 0, sizeof desc);
    desc.n_objects = 3; desc.objects = objs; desc.n_materials = 3; desc.materials = mats;

    rr_frame_params p;
    memset(&p, 0, sizeof p);
    p.xres = w; p.yres = h; p.xfov = 1.0f; p.yfov = (float)h / (float)w;
    p.cam_position[1] = -150.0f; p.cam_position[2] = -300.0f;
    /* Quat::from_pyr((0, -pi/2, -pi/2)) of main.rs:262-266: looks along +z */
    p.cam_rotation[0] = -0.5f; p.cam_rotation[1] = -0.5f; p.cam_rotation[2] = -0.5f; p.cam_rotation[3] = 0.5f;
    const float l[3] = {50.0f, 60.0f, -50.0f}, ln = sqrtf(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    for (int k = 0; k < 3; ++k) p.light[k] = l[k] / ln;
    p.max_reflections = 3; p.max_refractions = 10; p.bg_kind = RR_BG_BGCOLOR; p.band_count = 1; p.band_rows = 1; p.band_span = 1;

    rr_scene *scene = NULL;
    void *frame = NULL;
    int rc = 2;
    if (check(rr_scene_create(&desc, 0, &scene), "rr_scene_create")) return 2;
    if (!check(rr_host_alloc((size_t)w * h * 3, &frame), "rr_host_alloc") &&
        !check(rr_render_rgb8(scene, &p, (uint8_t *)frame, 0), "rr_render_rgb8")) {
        float ms = 0.0f;
        rr_last_kernel_ms(scene, &ms);
        FILE *f = fopen(argv[3], "wb");
        if (f) {
            fprintf(f, "P6\n%d %d\n255\n", w, h);
            fwrite(frame, 1, (size_t)w * h * 3, f);
            fclose(f);
            printf("%dx%d rendered, kernel %.3f ms -> %s\n", w, h, ms, argv[3]);
            rc = 0;
        } else {
            perror(argv[3]);
        }
    }
    if (frame) rr_host_free(frame);
    rr_scene_destroy(scene);
    return rc;
}
