// rr_host.hpp — C++ host layer above the C ABI: the mirror of ray-rust's public scene/render API.
//
// The reference is a Rust crate (src/lib.rs exports quat, render, vec3); no Rust toolchain exists in
// this image, so the host side that a Rust adapter would provide is written in C++ with the same
// names, argument meaning and error behaviour:
//   Vec3 (vec3.rs), Quat (quat.rs), RenderColor, RenderMaterial (+ builder methods), RenderSphere,
//   RenderFloor, RenderObject, Camera, CameraKeyframe, RenderEnv (+ builders, serialize/deserialize),
//   render(ren, pointproc, thread_count), render_frames(...) — render.rs:23-989.
// Everything that computes a pixel goes through include/rr_ffi.h to the CUDA kernels; there is no
// CPU rendering path in this layer.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rr_ffi.h"

namespace rr {

constexpr int MAX_REFLECTIONS = 3;   // render.rs:11
constexpr int MAX_REFRACTIONS = 10;  // render.rs:12
constexpr float PI = 3.14159265358979323846264338327950288f;

struct Vec3 {  // vec3.rs
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    static Vec3 zero() { return Vec3(); }
    float dot(const Vec3 &b) const { return x * b.x + y * b.y + z * b.z; }
    float squared_len() const { return x * x + y * y + z * z; }
    float len() const;
    Vec3 normalized() const;
    Vec3 operator+(const Vec3 &o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
    Vec3 operator-(const Vec3 &o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
    Vec3 operator*(float o) const { return Vec3(x * o, y * o, z * o); }
};

struct Quat {  // quat.rs
    float x = 0, y = 0, z = 0, w = 0;
    Quat() = default;
    Quat(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    Quat conjugated() const { return Quat(-x, -y, -z, w); }
    Quat mul(const Quat &o) const;
    Quat operator*(const Quat &o) const { return mul(o); }
    Vec3 transform(const Vec3 &v) const;
    float dot(const Quat &b) const { return x * b.x + y * b.y + z * b.z + w * b.w; }
    bool operator==(const Quat &o) const { return x == o.x && y == o.y && z == o.z && w == o.w; }
    Quat slerp(const Quat &o, float t) const;
    static Quat rotation(float p, float sx, float sy, float sz);
    static Quat from_pyr(const Vec3 &pyr);
};

struct RenderColor {  // render.rs:23-42
    float r = 0, g = 0, b = 0;
    RenderColor() = default;
    RenderColor(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {}
    static RenderColor zero() { return RenderColor(); }
};

enum class RenderPattern { Solid, Checkerboard, RepeatedGradation };  // render.rs:44-49
enum class UVMap { XY, YZ, ZX, LL };                                  // render.rs:51-57
enum class TextureFilter { Nearest, Bilinear };                       // render.rs:59-63

struct TextureRgb8 {  // what image::open() yields when it is DynamicImage::ImageRgb8 (render.rs:251)
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgb8;
};

class RenderMaterial {  // render.rs:82-181
public:
    RenderMaterial(std::string name, RenderColor diffuse, RenderColor specular, int pn, float t, float n);
    const std::string &get_name() const { return name_; }
    RenderMaterial &glow_dist(float v) { glow_dist_ = v; return *this; }
    RenderMaterial &frac(RenderColor f) { frac_ = f; return *this; }
    RenderMaterial &pattern(RenderPattern p) { pattern_ = p; return *this; }
    RenderMaterial &pattern_scale(float v) { pattern_scale_ = v; return *this; }
    RenderMaterial &pattern_angle_scale(float v) { pattern_angle_scale_ = v; return *this; }
    RenderMaterial &texture(const std::string &file);     // throws std::runtime_error("texture image file load failed")
    RenderMaterial &texture_ok(const std::string &file);  // ignores a failed load quietly
    RenderMaterial &texture_filter(TextureFilter f) { texture_filter_ = f; return *this; }

    std::string name_;
    RenderColor diffuse_, specular_;
    int pn_;
    float t_, n_;
    float glow_dist_ = 0.0f;
    RenderColor frac_{1.0f, 1.0f, 1.0f};
    RenderPattern pattern_ = RenderPattern::Solid;
    float pattern_scale_ = 1.0f, pattern_angle_scale_ = 1.0f;
    std::string texture_name_;
    std::shared_ptr<TextureRgb8> texture_;  // None when absent or not RGB8
    TextureFilter texture_filter_ = TextureFilter::Nearest;
};
using MaterialRef = std::shared_ptr<RenderMaterial>;  // Arc<RenderMaterial>

struct RenderObject {  // render.rs:585-589 (enum of RenderSphere | RenderFloor)
    enum Kind { Sphere, Floor } kind = Sphere;
    MaterialRef material;
    float r = 0.0f;   // sphere
    Vec3 org;         // centre
    Vec3 face_normal; // floor
    UVMap uvmap_ = UVMap::XY;
    RenderObject &uvmap(UVMap v) { uvmap_ = v; return *this; }
};
struct RenderSphere {  // render.rs:386-399
    static RenderObject make(MaterialRef m, float r, Vec3 org);
};
struct RenderFloor {  // render.rs:495-514
    static RenderObject make(MaterialRef m, Vec3 org, Vec3 face_normal);
    static RenderObject new_raw(MaterialRef m, Vec3 org, Vec3 face_normal) { return make(std::move(m), org, face_normal); }
};

struct Camera {  // render.rs:617-622
    Vec3 position, pyr;
    Quat rotation;
    Camera() = default;
    Camera(Vec3 pos, Vec3 pyr_) : position(pos), pyr(pyr_), rotation(Quat::from_pyr(pyr_)) {}
};
struct CameraKeyframe {  // render.rs:634-640
    Camera camera;
    Vec3 velocity;
    bool has_target = false;
    Vec3 camera_target;
    float duration = 0.0f;
};

struct DeserializeError : std::runtime_error {  // render.rs:341-366
    std::string s;
    explicit DeserializeError(const std::string &m) : std::runtime_error("Deserialize error: " + m), s(m) {}
};
struct RenderError : std::runtime_error {  // anyhow::Error of render()
    int code;
    RenderError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

enum class BgProc { BgColor, Black };  // RenderEnv.bgproc is a fn pointer in the reference (render.rs:661)

class DeviceScene;

class RenderEnv {  // render.rs:646-799
public:
    RenderEnv(Vec3 cam, Vec3 pyr, int xres, int yres, float xfov, float yfov, BgProc bgproc = BgProc::BgColor);
    RenderEnv &materials(std::map<std::string, MaterialRef> m) { materials_ = std::move(m); invalidate(); return *this; }
    RenderEnv &objects(std::vector<RenderObject> o) { objects_ = std::move(o); invalidate(); return *this; }
    RenderEnv &light(Vec3 l) { light_ = l.normalized(); return *this; }
    RenderEnv &use_raymarching(bool f) { use_raymarching_ = f; return *this; }
    RenderEnv &glow_effect(bool some, float v = 0.0f) { glow_some_ = some; glow_value_ = v; return *this; }
    std::string serialize() const;            // render.rs:735-760
    void deserialize(const std::string &s);   // render.rs:762-799; throws DeserializeError
    rr_frame_params frame_params() const;
    // Drops the device-resident copies of the scene. materials()/objects() call it; call it yourself after editing a
    // shared RenderMaterial in place through a MaterialRef (the device copy is a snapshot taken at the first render).
    void invalidate();

    Camera camera;
    std::vector<CameraKeyframe> camera_motion;
    int xres, yres;
    float xfov, yfov;
    std::map<std::string, MaterialRef> materials_;
    std::vector<RenderObject> objects_;
    Vec3 light_{0.0f, 0.0f, 1.0f};
    BgProc bgproc;
    bool use_raymarching_ = false;
    bool glow_some_ = false;
    float glow_value_ = 0.0f;
    int max_reflections = MAX_REFLECTIONS, max_refractions = MAX_REFRACTIONS;
    mutable std::map<int, std::shared_ptr<DeviceScene>> device_;  // flattened scene resident on GPU <key> (built lazily, locked)
};

// flattened POD view of a RenderEnv for rr_scene_create (pointers are into this object)
struct FlatScene {
    std::vector<rr_object> objects;
    std::vector<rr_material> materials;
    std::vector<rr_texture> textures;
    std::vector<std::shared_ptr<TextureRgb8>> keep;
    rr_scene_desc desc() const;
};
FlatScene flatten(const RenderEnv &ren);

class DeviceScene {  // owns an rr_scene handle
public:
    DeviceScene(const RenderEnv &ren, int device);
    ~DeviceScene();
    rr_scene *handle = nullptr;
};

// Page-locked host frame (rr_host_alloc): what every caller of this layer hands to the C ABI, so that device-to-host
// copies run asynchronously at full PCIe rate and long kernels can store straight into the frame.
class PinnedFrame {
public:
    PinnedFrame() = default;
    explicit PinnedFrame(size_t bytes) { resize(bytes); }
    PinnedFrame(const PinnedFrame &) = delete;
    PinnedFrame &operator=(const PinnedFrame &) = delete;
    PinnedFrame(PinnedFrame &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    ~PinnedFrame();
    void resize(size_t bytes);  // grows only; throws RenderError when no CUDA device is usable
    uint8_t *data() const { return p_; }
    size_t size() const { return n_; }
private:
    uint8_t *p_ = nullptr;
    size_t n_ = 0;
};

// Makes the scene resident on `device` now (CUDA context + flattened upload) instead of at the first render. The
// reference builds its RenderEnv before it starts the "Rendering time" clock (main.rs:154-316); the CLI does the same
// with the device copy of the scene.
void upload_scene(const RenderEnv &ren, int device = 0);

using PointProc = std::function<void(int, int, const RenderColor &)>;
// render(), render.rs:801-805. Calls pointproc(x, y, colour) once per pixel, row-major, on the caller
// thread. thread_count is accepted for signature compatibility and ignored (device grid).
void render(const RenderEnv &ren, const PointProc &pointproc, int thread_count, int device = 0);
// Fast path of the three stock callers: render() + the putpoint quantiser (main.rs:148-152).
void render_rgb8(const RenderEnv &ren, uint8_t *data, int device = 0);
// render_frames(), render.rs:926-989. frame_proc(i, data, len) receives frame i as `&[u8]` (RGB8, len = 3*width*height),
// in frame order, on the calling thread, exactly like the reference's FnMut(i32, &[u8]). Unlike the reference the frames
// are pipelined: all camera poses are interpolated first, then frames are dealt round-robin to `devices` (empty = every
// visible GPU) with two page-locked frames in flight per GPU (rr_render_rgb8_async), so while frame_proc encodes frame k
// the kernels and copies of the next frames are already running. width/height must equal ren.xres/yres (the reference's
// putpoint would panic on a mismatch); RenderError otherwise.
using FrameProc = std::function<void(int, const uint8_t *, size_t)>;
void render_frames(RenderEnv &ren, size_t width, size_t height, const FrameProc &frame_proc, int thread_count,
                   const std::vector<int> &devices = {});
// the camera poses render_frames() renders, in order (exposed for tests and tools)
std::vector<Camera> interpolate_camera_motion(const RenderEnv &ren, bool verbose = false);

// built-in scene of main.rs:154-276 and the synthetic scene of BASELINE configs[3]
RenderEnv default_scene(int width, int height, bool use_raymarching, bool glow_some, float glow_value);
RenderEnv synthetic_scene(int width, int height, int n_spheres = 1024, uint64_t seed = 20261018ull);

// run_webserver(), webserver.rs:324-333: serves /, /image, /render?x&y&z&yaw&pitch from the resident device scene
int run_webserver(const RenderEnv &ren, int width, int height, int port, int device = 0);

// PNG (image::save_buffer(.., ColorType::Rgb8) / image::open for textures)
void save_png_rgb8(const std::string &path, const uint8_t *rgb, uint32_t w, uint32_t h);
std::vector<uint8_t> encode_png_rgb8(const uint8_t *rgb, uint32_t w, uint32_t h);
std::shared_ptr<TextureRgb8> load_png_rgb8(const std::string &path);  // nullptr unless it decodes to RGB8
std::shared_ptr<TextureRgb8> load_jpeg_rgb8(const std::vector<uint8_t> &file_bytes);  // baseline 3-component JPEG (rr_jpeg.cpp), else nullptr
// image::open(path).ok() restricted to DynamicImage::ImageRgb8 (render.rs:165-181, :251): PNG or baseline JPEG by signature
std::shared_ptr<TextureRgb8> load_image_rgb8(const std::string &path);

}  // namespace rr
