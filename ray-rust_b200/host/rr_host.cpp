// rr_host.cpp — scene model, flattening and the render()/render_frames() entry points (see rr_host.hpp).
#include "rr_host.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>

namespace rr {

// ---- vec3.rs / quat.rs (host-side f32, same operation order as the reference) -----------------
float Vec3::len() const { return std::sqrt(squared_len()); }
Vec3 Vec3::normalized() const {
    float l = len();
    return Vec3(x / l, y / l, z / l);
}

Quat Quat::mul(const Quat &qb) const {  // quat.rs:63-72
    const Quat &qa = *this;
    return Quat(qa.y * qb.z - qa.z * qb.y + qa.x * qb.w + qa.w * qb.x, qa.z * qb.x - qa.x * qb.z + qa.y * qb.w + qa.w * qb.y,
                qa.x * qb.y - qa.y * qb.x + qa.z * qb.w + qa.w * qb.z, -qa.x * qb.x - qa.y * qb.y - qa.z * qb.z + qa.w * qb.w);
}
Vec3 Quat::transform(const Vec3 &v) const {  // quat.rs:74-80
    Quat qr = mul(Quat(v.x, v.y, v.z, 0.0f));
    Quat o = qr.mul(conjugated());
    return Vec3(o.x, o.y, o.z);
}
Quat Quat::rotation(float p, float sx, float sy, float sz) {  // quat.rs:92-95
    float len = sinf(p / 2.0f);
    return Quat(len * sx, len * sy, len * sz, cosf(p / 2.0f));
}
Quat Quat::from_pyr(const Vec3 &pyr) {  // quat.rs:129-134
    Quat mx = rotation(pyr.z, 1.0f, 0.0f, 0.0f);
    Quat my = rotation(pyr.y, 0.0f, 0.0f, 1.0f);
    Quat mp = rotation(pyr.x, 0.0f, 1.0f, 0.0f);
    return mx * my * mp;
}
Quat Quat::slerp(const Quat &o, float t) const {  // quat.rs:97-127
    float qr = dot(o);
    float ss = 1.0f - qr * qr;
    if (ss <= std::sqrt(1e-10f) || *this == o) return *this;
    float sp = std::sqrt(ss);
    float ph = acosf(qr);
    float pt = ph * t;
    float t1 = sinf(pt) / sp;
    float t0 = sinf(ph - pt) / sp;
    if (qr < 0.0f) t1 *= -1.0f;
    return Quat(x * t0 + o.x * t1, y * t0 + o.y * t1, z * t0 + o.z * t1, w * t0 + o.w * t1);
}

// ---- materials / objects ------------------------------------------------------------------------
RenderMaterial::RenderMaterial(std::string name, RenderColor diffuse, RenderColor specular, int pn, float t, float n)
    : name_(std::move(name)), diffuse_(diffuse), specular_(specular), pn_(pn), t_(t), n_(n) {}

RenderMaterial &RenderMaterial::texture(const std::string &file) {  // render.rs:165-174
    texture_name_ = file;
    texture_ = load_image_rgb8(file);
    if (!texture_) throw std::runtime_error("texture image file load failed");
    return *this;
}
RenderMaterial &RenderMaterial::texture_ok(const std::string &file) {  // render.rs:177-181
    texture_name_ = file;
    texture_ = load_image_rgb8(file);
    return *this;
}

RenderObject RenderSphere::make(MaterialRef m, float r, Vec3 org) {
    RenderObject o;
    o.kind = RenderObject::Sphere;
    o.material = std::move(m);
    o.r = r;
    o.org = org;
    return o;
}
RenderObject RenderFloor::make(MaterialRef m, Vec3 org, Vec3 face_normal) {
    RenderObject o;
    o.kind = RenderObject::Floor;
    o.material = std::move(m);
    o.org = org;
    o.face_normal = face_normal;
    return o;
}

RenderEnv::RenderEnv(Vec3 cam, Vec3 pyr, int xres_, int yres_, float xfov_, float yfov_, BgProc bg)
    : camera(cam, pyr), xres(xres_), yres(yres_), xfov(xfov_), yfov(yfov_), bgproc(bg) {}

rr_frame_params RenderEnv::frame_params() const {
    rr_frame_params p;
    std::memset(&p, 0, sizeof p);
    p.xres = xres; p.yres = yres; p.xfov = xfov; p.yfov = yfov;
    p.cam_position[0] = camera.position.x; p.cam_position[1] = camera.position.y; p.cam_position[2] = camera.position.z;
    p.cam_rotation[0] = camera.rotation.x; p.cam_rotation[1] = camera.rotation.y;
    p.cam_rotation[2] = camera.rotation.z; p.cam_rotation[3] = camera.rotation.w;
    p.light[0] = light_.x; p.light[1] = light_.y; p.light[2] = light_.z;
    p.use_raymarching = use_raymarching_ ? 1 : 0;
    p.glow_enabled = glow_some_ ? 1 : 0;
    p.glow_effect = glow_some_ ? glow_value_ : 0.0f;
    p.max_reflections = max_reflections; p.max_refractions = max_refractions;
    p.bg_kind = bgproc == BgProc::BgColor ? RR_BG_BGCOLOR : RR_BG_BLACK;
    p.band_rows = 0; p.band_index = 0; p.band_count = 1; p.band_span = 1;
    return p;
}

// ---- flatten to the C ABI -----------------------------------------------------------------------
rr_scene_desc FlatScene::desc() const {
    rr_scene_desc d;
    d.n_objects = (uint32_t)objects.size(); d.objects = objects.data();
    d.n_materials = (uint32_t)materials.size(); d.materials = materials.data();
    d.n_textures = (uint32_t)textures.size(); d.textures = textures.data();
    return d;
}

FlatScene flatten(const RenderEnv &ren) {
    FlatScene f;
    std::map<const RenderMaterial *, int> index;  // de-duplicate by Arc pointer
    for (const RenderObject &o : ren.objects_) {
        const RenderMaterial *m = o.material.get();
        auto it = index.find(m);
        int mi;
        if (it == index.end()) {
            mi = (int)f.materials.size();
            index[m] = mi;
            rr_material c;
            std::memset(&c, 0, sizeof c);
            c.diffuse[0] = m->diffuse_.r; c.diffuse[1] = m->diffuse_.g; c.diffuse[2] = m->diffuse_.b;
            c.specular[0] = m->specular_.r; c.specular[1] = m->specular_.g; c.specular[2] = m->specular_.b;
            c.pn = m->pn_; c.t = m->t_; c.n = m->n_; c.glow_dist = m->glow_dist_;
            c.frac[0] = m->frac_.r; c.frac[1] = m->frac_.g; c.frac[2] = m->frac_.b;
            c.pattern = (int)m->pattern_; c.pattern_scale = m->pattern_scale_; c.pattern_angle_scale = m->pattern_angle_scale_;
            c.texture_filter = (int)m->texture_filter_;
            c.texture = -1;
            if (m->texture_) {
                c.texture = (int)f.textures.size();
                rr_texture t;
                t.width = m->texture_->width; t.height = m->texture_->height; t.rgb8 = m->texture_->rgb8.data();
                f.textures.push_back(t);
                f.keep.push_back(m->texture_);
            }
            f.materials.push_back(c);
        } else {
            mi = it->second;
        }
        rr_object c;
        std::memset(&c, 0, sizeof c);
        c.kind = o.kind == RenderObject::Sphere ? RR_SPHERE : RR_FLOOR;
        c.material = mi;
        c.uvmap = (int)o.uvmap_;
        c.r = o.r;
        c.org[0] = o.org.x; c.org[1] = o.org.y; c.org[2] = o.org.z;
        if (o.kind == RenderObject::Floor) {
            c.face_normal[0] = o.face_normal.x; c.face_normal[1] = o.face_normal.y; c.face_normal[2] = o.face_normal.z;
        }
        f.objects.push_back(c);
    }
    return f;
}

static void check(int rc) {
    if (rc != RR_OK) throw RenderError(rc, rr_last_error());
}

DeviceScene::DeviceScene(const RenderEnv &ren, int device) {
    FlatScene f = flatten(ren);
    rr_scene_desc d = f.desc();
    check(rr_scene_create(&d, device, &handle));
}
DeviceScene::~DeviceScene() {
    if (handle) rr_scene_destroy(handle);
}

// One flattened copy of the scene per GPU, built at the first render on that GPU. The cache sits in a const RenderEnv
// (render() takes `&RenderEnv`, and several threads may render one environment, webserver.rs:268-280), so it is locked.
static std::mutex g_device_cache_mu;
static std::shared_ptr<DeviceScene> device_scene(const RenderEnv &ren, int device) {
    std::lock_guard<std::mutex> lk(g_device_cache_mu);
    auto it = ren.device_.find(device);
    if (it != ren.device_.end()) return it->second;
    auto d = std::make_shared<DeviceScene>(ren, device);
    ren.device_[device] = d;
    return d;
}
void RenderEnv::invalidate() {
    std::lock_guard<std::mutex> lk(g_device_cache_mu);
    device_.clear();
}

PinnedFrame::~PinnedFrame() {
    if (p_) rr_host_free(p_);
}
void PinnedFrame::resize(size_t bytes) {
    if (bytes <= n_) return;
    if (p_) rr_host_free(p_);
    p_ = nullptr;
    n_ = 0;
    void *q = nullptr;
    check(rr_host_alloc(bytes, &q));
    p_ = static_cast<uint8_t *>(q);
    n_ = bytes;
}

void upload_scene(const RenderEnv &ren, int device) { device_scene(ren, device); }

void render(const RenderEnv &ren, const PointProc &pointproc, int /*thread_count*/, int device) {
    auto scene = device_scene(ren, device);
    rr_frame_params p = ren.frame_params();
    const size_t n = (size_t)3 * ren.xres * ren.yres;
    if (n == 0) return;
    PinnedFrame buf(n * sizeof(float));
    const float *px = reinterpret_cast<const float *>(buf.data());
    check(rr_render_f32(scene->handle, &p, reinterpret_cast<float *>(buf.data())));
    for (int iy = 0; iy < ren.yres; ++iy)
        for (int ix = 0; ix < ren.xres; ++ix) {
            const float *c = &px[(size_t)3 * ((size_t)iy * ren.xres + ix)];
            pointproc(ix, iy, RenderColor(c[0], c[1], c[2]));
        }
}

void render_rgb8(const RenderEnv &ren, uint8_t *data, int device) {
    auto scene = device_scene(ren, device);
    rr_frame_params p = ren.frame_params();
    if (ren.xres == 0 || ren.yres == 0) return;
    check(rr_render_rgb8(scene->handle, &p, data, 0));
}

// hermite_interpolate, render.rs:907-924
static float hermite_f32(float t, float x0, float x1, float v0, float v1) {
    const float h = 1.0f;
    float d = x0, c = v0;
    float r = x1 - x0 - h * v0;
    float s = v1 - v0;
    float a = (h * s - 2.0f * r) / h / h / h;
    float b = (-h * s + 3.0f * r) / h / h;
    return a * t * t * t + b * t * t + c * t + d;
}
static Vec3 hermite(float t, const Vec3 &x0, const Vec3 &x1, const Vec3 &v0, const Vec3 &v1) {
    return Vec3(hermite_f32(t, x0.x, x1.x, v0.x, v1.x), hermite_f32(t, x0.y, x1.y, v0.y, v1.y), hermite_f32(t, x0.z, x1.z, v0.z, v1.z));
}

// The camera of every frame render_frames() produces, render.rs:935-971 (Hermite position, slerp or look-at rotation).
std::vector<Camera> interpolate_camera_motion(const RenderEnv &ren, bool verbose) {
    std::vector<Camera> out;
    Camera prev_camera = ren.camera;
    Vec3 prev_velocity = Vec3::zero();
    float total_frames = 0.0f;
    for (const auto &m : ren.camera_motion) total_frames += m.duration;
    int accum_frame = 0;
    const float frame_step = 0.5f;
    for (size_t n = 0; n < ren.camera_motion.size(); ++n) {
        const CameraKeyframe &frame = ren.camera_motion[n];
        const Vec3 v0 = prev_velocity, v1 = frame.velocity;
        if (verbose) printf("keyframe %zu / %zu, v0: %g,%g,%g\n", n, ren.camera_motion.size(), v0.x, v0.y, v0.z);
        const int count = (int)(frame.duration / frame_step);
        for (int i = 0; i < count; ++i) {
            const float f = (float)i / (frame.duration / frame_step);
            if (verbose) printf("Rendering frame %d / %g, v0: %g,%g\n", accum_frame, total_frames, v0.x, v0.y);
            Camera cam = ren.camera;
            cam.position = hermite(f, prev_camera.position, frame.camera.position, v0, v1);
            if (frame.has_target) {  // look-at, render.rs:961-967
                Vec3 delta = frame.camera_target - cam.position;
                float pitch = atan2f(delta.y, std::sqrt(delta.x * delta.x + delta.z * delta.z));
                float yaw = -atan2f(delta.z, delta.x);
                cam.rotation = Quat::rotation(yaw, 0.0f, 1.0f, 0.0f) * Quat::rotation(pitch, 0.0f, 0.0f, 1.0f) *
                               Quat::rotation(-PI / 2.0f, 1.0f, 0.0f, 0.0f);
            } else {
                cam.rotation = prev_camera.rotation.slerp(frame.camera.rotation, f);
            }
            out.push_back(cam);
            accum_frame += 1;
        }
        prev_camera = frame.camera;
        prev_velocity = frame.velocity;
    }
    return out;
}

void render_frames(RenderEnv &ren, size_t width, size_t height, const FrameProc &frame_proc, int thread_count,
                   const std::vector<int> &devices_in) {
    (void)thread_count;
    if ((long long)width != ren.xres || (long long)height != ren.yres)
        throw RenderError(RR_ERR_BAD_ARG, "render_frames: width/height differ from the environment's resolution");
    const std::vector<Camera> cams = interpolate_camera_motion(ren, true);
    if (cams.empty()) return;
    std::vector<int> devices = devices_in;
    if (devices.empty()) {
        int n = 0;
        check(rr_device_count(&n));
        for (int d = 0; d < n; ++d) devices.push_back(d);
        if (devices.empty()) throw RenderError(RR_ERR_CUDA, "no CUDA device");
    }
    if (devices.size() > cams.size()) devices.resize(cams.size());
    const size_t bytes = 3 * width * height;
    constexpr int DEPTH = 2;  // frames in flight per GPU (each has 4 lanes; two keep kernel, copy and frame_proc overlapped)
    struct Slot { std::shared_ptr<DeviceScene> scene; PinnedFrame buf; int32_t ticket = -1; };
    std::vector<Slot> slots(devices.size() * DEPTH);
    for (size_t k = 0; k < slots.size(); ++k) {
        slots[k].scene = device_scene(ren, devices[k % devices.size()]);
        slots[k].buf.resize(bytes ? bytes : 1);
    }
    rr_frame_params base = ren.frame_params();
    auto submit = [&](size_t i) {
        Slot &s = slots[i % slots.size()];
        rr_frame_params p = base;
        const Camera &c = cams[i];
        p.cam_position[0] = c.position.x; p.cam_position[1] = c.position.y; p.cam_position[2] = c.position.z;
        p.cam_rotation[0] = c.rotation.x; p.cam_rotation[1] = c.rotation.y; p.cam_rotation[2] = c.rotation.z; p.cam_rotation[3] = c.rotation.w;
        check(rr_render_rgb8_async(s.scene->handle, &p, s.buf.data(), 0, &s.ticket));
    };
    auto drain = [&]() {  // never leave tickets behind, whatever happens
        for (Slot &s : slots)
            if (s.ticket >= 0) { rr_render_wait(s.scene->handle, s.ticket, nullptr); s.ticket = -1; }
    };
    try {
        for (size_t i = 0; i < cams.size() && i < slots.size(); ++i) submit(i);
        for (size_t i = 0; i < cams.size(); ++i) {
            Slot &s = slots[i % slots.size()];
            const int32_t t = s.ticket;
            s.ticket = -1;
            check(rr_render_wait(s.scene->handle, t, nullptr));
            frame_proc((int)i, s.buf.data(), bytes);       // the next frames are rendering meanwhile
            if (i + slots.size() < cams.size()) submit(i + slots.size());
        }
    } catch (...) {
        drain();
        throw;
    }
    ren.camera = cams.back();  // the reference leaves the last interpolated pose in ren.camera
}

// ---- built-in scene, main.rs:154-276 --------------------------------------------------------------
RenderEnv default_scene(int width, int height, bool use_raymarching, bool glow_some, float glow_value) {
    const float xfov = 1.0f;
    const float yfov = (float)height / (float)width;  // main.rs:135-136
    std::map<std::string, MaterialRef> materials;
    auto floor_material = std::make_shared<RenderMaterial>("floor", RenderColor(1.0f, 1.0f, 0.0f), RenderColor(0.0f, 0.0f, 0.0f), 0, 0.0f, 0.0f);
    floor_material->pattern(RenderPattern::RepeatedGradation).pattern_scale(300.0f).pattern_angle_scale(0.2f).texture_ok("bar.png");
    materials["floor"] = floor_material;
    auto mirror = std::make_shared<RenderMaterial>("mirror", RenderColor(0.0f, 0.0f, 0.0f), RenderColor(1.0f, 1.0f, 1.0f), 24, 0.0f, 0.0f);
    mirror->frac(RenderColor(1.0f, 1.0f, 1.0f));
    auto red = std::make_shared<RenderMaterial>("red", RenderColor(0.8f, 0.0f, 0.0f), RenderColor(0.0f, 0.0f, 0.0f), 24, 0.0f, 0.0f);
    red->glow_dist(5.0f);
    auto transparent = std::make_shared<RenderMaterial>("transparent", RenderColor(0.0f, 0.0f, 0.0f), RenderColor(0.0f, 0.0f, 0.0f), 0, 1.0f, 1.5f);
    transparent->frac(RenderColor(1.49998f, 1.49999f, 1.5f));
    std::vector<RenderObject> objects;
    objects.push_back(RenderFloor::new_raw(floor_material, Vec3(0.0f, -300.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f)).uvmap(UVMap::ZX));
    objects.push_back(RenderSphere::make(mirror, 80.0f, Vec3(0.0f, -30.0f, 172.0f)));
    objects.push_back(RenderSphere::make(mirror, 80.0f, Vec3(-200.0f, -30.0f, 172.0f)));
    objects.push_back(RenderSphere::make(red, 80.0f, Vec3(-200.0f, -200.0f, 172.0f)));
    objects.push_back(RenderSphere::make(transparent, 100.0f, Vec3(70.0f, -200.0f, 150.0f)));
    RenderEnv ren(Vec3(0.0f, -150.0f, -300.0f), Vec3(0.0f, -PI / 2.0f, -PI / 2.0f), width, height, xfov, yfov);
    ren.materials(materials).objects(objects).light(Vec3(50.0f, 60.0f, -50.0f)).use_raymarching(use_raymarching).glow_effect(glow_some, glow_value);
    return ren;
}

// ---- synthetic scene; identical to ray_rust_b200.scene.synthetic_scene ----
namespace {
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    float unit() { return (float)(next() >> 40) * 5.9604644775390625e-8f; }  // 24-bit uniform, exact in f32
    float uniform(float a, float b) { return a + (b - a) * unit(); }
    int below(int n) { return (int)(next() % (uint64_t)n); }
};
}  // namespace

RenderEnv synthetic_scene(int width, int height, int n_spheres, uint64_t seed) {
    SplitMix64 rng(seed);
    auto floor_material = std::make_shared<RenderMaterial>("floor", RenderColor(1.0f, 1.0f, 0.0f), RenderColor(0.0f, 0.0f, 0.0f), 0, 0.0f, 0.0f);
    floor_material->pattern(RenderPattern::RepeatedGradation).pattern_scale(300.0f).pattern_angle_scale(0.2f);
    std::vector<MaterialRef> mats;
    for (int i = 0; i < 6; ++i) {
        float s = rng.uniform(0.5f, 1.0f);
        float d0 = rng.uniform(0.0f, 0.3f), d1 = rng.uniform(0.0f, 0.3f), d2 = rng.uniform(0.0f, 0.3f);
        mats.push_back(std::make_shared<RenderMaterial>("mirror" + std::to_string(i), RenderColor(d0, d1, d2), RenderColor(s, s, s), 24, 0.0f, 0.0f));
    }
    for (int i = 0; i < 5; ++i) {
        float d0 = rng.uniform(0.1f, 1.0f), d1 = rng.uniform(0.1f, 1.0f), d2 = rng.uniform(0.1f, 1.0f);
        auto m = std::make_shared<RenderMaterial>("diffuse" + std::to_string(i), RenderColor(d0, d1, d2), RenderColor(0.0f, 0.0f, 0.0f), 24, 0.0f, 0.0f);
        m->pattern(i % 2 == 0 ? RenderPattern::Solid : RenderPattern::Checkerboard).pattern_scale(10.0f);
        mats.push_back(m);
    }
    for (int i = 0; i < 5; ++i) {
        float t = rng.uniform(0.5f, 1.0f);
        float n = rng.uniform(1.2f, 1.8f);
        RenderColor spec(0.0f, 0.0f, 0.0f);
        if (i >= 3) {
            float s = rng.uniform(0.2f, 0.5f);
            spec = RenderColor(s, s, s);
        }
        mats.push_back(std::make_shared<RenderMaterial>("glass" + std::to_string(i), RenderColor(0.0f, 0.0f, 0.0f), spec, 0, t, n));
    }
    std::vector<RenderObject> objects;
    objects.push_back(RenderFloor::new_raw(floor_material, Vec3(0.0f, -300.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f)).uvmap(UVMap::ZX));
    for (int k = 0; k < n_spheres; ++k) {
        MaterialRef m = mats[rng.below((int)mats.size())];
        float r = rng.uniform(15.0f, 45.0f);
        float x = rng.uniform(-900.0f, 900.0f);
        float y = rng.uniform(-280.0f, 300.0f);
        float z = rng.uniform(-100.0f, 1900.0f);
        objects.push_back(RenderSphere::make(m, r, Vec3(x, y, z)));
    }
    std::map<std::string, MaterialRef> materials;
    materials[floor_material->name_] = floor_material;
    for (auto &m : mats) materials[m->name_] = m;
    RenderEnv ren(Vec3(0.0f, -150.0f, -300.0f), Vec3(0.0f, -PI / 2.0f, -PI / 2.0f), width, height, 1.0f, (float)height / (float)width);
    ren.materials(materials).objects(objects).light(Vec3(50.0f, 60.0f, -50.0f));
    return ren;
}

}  // namespace rr
