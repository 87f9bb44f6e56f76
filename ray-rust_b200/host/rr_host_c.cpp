// rr_host_c.cpp — small extern "C" facade over the C++ host layer so the Python tests can drive it
// (scene construction, YAML round trips, flattening) without a GPU.
#include <chrono>
#include <cstring>

#include <zlib.h>

#include "rr_host.hpp"

namespace {
thread_local std::string g_err;
char *dup(const std::string &s) {
    char *p = (char *)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
struct Handle {
    rr::RenderEnv ren;
    rr::FlatScene flat;
};
}  // namespace

extern "C" {

const char *rrh_last_error() { return g_err.c_str(); }
void rrh_free(void *p) { free(p); }

// kind: 0 = default scene (main.rs:154-276), 1 = synthetic scene (n_spheres, seed)
void *rrh_env_new(int kind, int width, int height, int raymarch, int glow_some, float glow, int n_spheres, uint64_t seed) {
    try {
        rr::RenderEnv ren = kind == 0 ? rr::default_scene(width, height, raymarch != 0, glow_some != 0, glow)
                                      : rr::synthetic_scene(width, height, n_spheres, seed);
        if (kind != 0) ren.use_raymarching(raymarch != 0).glow_effect(glow_some != 0, glow);
        return new Handle{std::move(ren), {}};
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void rrh_env_free(void *h) { delete (Handle *)h; }
char *rrh_env_serialize(void *h) { return dup(((Handle *)h)->ren.serialize()); }
int rrh_env_deserialize(void *h, const char *text) {
    try {
        ((Handle *)h)->ren.deserialize(text);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
// flatten into the handle; returns pointers valid until the next call / free
int rrh_env_flatten(void *h, rr_scene_desc *desc, rr_frame_params *params) {
    Handle *H = (Handle *)h;
    H->flat = rr::flatten(H->ren);
    *desc = H->flat.desc();
    *params = H->ren.frame_params();
    return 0;
}
int rrh_env_limits(void *h, int *max_reflections, int *max_refractions, int *n_keyframes) {
    Handle *H = (Handle *)h;
    *max_reflections = H->ren.max_reflections;
    *max_refractions = H->ren.max_refractions;
    *n_keyframes = (int)H->ren.camera_motion.size();
    return 0;
}
// render through the C++ render() signature: pointproc is invoked once per pixel
typedef void (*rrh_pointproc)(int x, int y, float r, float g, float b, void *user);
int rrh_render(void *h, rrh_pointproc cb, void *user, int thread_count) {
    try {
        rr::render(((Handle *)h)->ren, [&](int x, int y, const rr::RenderColor &c) { cb(x, y, c.r, c.g, c.b, user); }, thread_count);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
int rrh_render_rgb8(void *h, uint8_t *out) {
    try {
        rr::render_rgb8(((Handle *)h)->ren, out);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
int rrh_png_roundtrip(const uint8_t *rgb, uint32_t w, uint32_t h, const char *path, uint8_t *back) {
    try {
        rr::save_png_rgb8(path, rgb, w, h);
        auto t = rr::load_image_rgb8(path);
        if (!t || t->width != w || t->height != h) return -2;
        memcpy(back, t->rgb8.data(), (size_t)w * h * 3);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// image::open(path) as the texture loader sees it: 0 and the RGB8 pixels, or -2 when the file is not an RGB8 image
// (`out` may be null to query the size only).
int rrh_load_image(const char *path, uint32_t *w, uint32_t *h, uint8_t *out, uint64_t cap) {
    try {
        auto t = rr::load_image_rgb8(path);
        if (!t) return -2;
        *w = t->width; *h = t->height;
        if (out) {
            if (cap < t->rgb8.size()) return -3;
            memcpy(out, t->rgb8.data(), t->rgb8.size());
        }
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// PNG encode alone (image::save_buffer's share of main.rs:316-348): best-of-reps milliseconds and the encoded size.
int rrh_png_encode_ms(const uint8_t *rgb, uint32_t w, uint32_t h, int reps, double *ms_best, uint64_t *png_bytes) {
    try {
        double best = 1e300;
        size_t n = 0;
        for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<uint8_t> png = rr::encode_png_rgb8(rgb, w, h);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            best = ms < best ? ms : best;
            n = png.size();
        }
        *ms_best = best;
        *png_bytes = n;
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
// render_frames (render.rs:926-989) of the handle's camera_motion over `n_devices` GPUs (0 = all).
// mode 0: frames are only checksummed (CRC-32 of each frame into crcs[i], up to max_crcs), 1: each frame is PNG-encoded
// in memory like the CLI does before writing it. Returns the number of frames, wall seconds in *seconds.
int rrh_render_frames(void *h, int mode, int n_devices, double *seconds, uint32_t *crcs, int max_crcs) {
    try {
        Handle *H = (Handle *)h;
        std::vector<int> devices;
        for (int d = 0; d < n_devices; ++d) devices.push_back(d);
        int frames = 0;
        const uint32_t w = (uint32_t)H->ren.xres, hh = (uint32_t)H->ren.yres;
        const rr::Camera saved = H->ren.camera;
        auto t0 = std::chrono::steady_clock::now();
        rr::render_frames(H->ren, w, hh, [&](int i, const uint8_t *data, size_t len) {
            if (mode == 1) {
                std::vector<uint8_t> png = rr::encode_png_rgb8(data, w, hh);
                if (crcs && i < max_crcs) crcs[i] = (uint32_t)png.size();
            } else if (crcs && i < max_crcs) {
                crcs[i] = (uint32_t)crc32(0L, data, (uInt)len);
            }
            frames = i + 1;
        }, 0, devices);
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        H->ren.camera = saved;
        return frames;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
// camera poses of render_frames(): 7 floats per frame (position xyz, rotation xyzw); returns the frame count
int rrh_camera_motion(void *h, float *out, int max_frames) {
    Handle *H = (Handle *)h;
    const std::vector<rr::Camera> cams = rr::interpolate_camera_motion(H->ren, false);
    for (size_t i = 0; i < cams.size() && (int)i < max_frames; ++i) {
        float *o = out + 7 * i;
        o[0] = cams[i].position.x; o[1] = cams[i].position.y; o[2] = cams[i].position.z;
        o[3] = cams[i].rotation.x; o[4] = cams[i].rotation.y; o[5] = cams[i].rotation.z; o[6] = cams[i].rotation.w;
    }
    return (int)cams.size();
}
}
