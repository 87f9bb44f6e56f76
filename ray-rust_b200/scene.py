"""Host-side mirror of ray-rust's scene model and render() entry point (Python binding).

Mirrors the public surface of the reference's `render.rs` / `main.rs` for the hot path:
RenderColor, RenderMaterial (+builder methods), RenderSphere, RenderFloor, RenderEnv (+builder
methods, serialize/deserialize), render(ren, pointproc, thread_count), and the built-in scene of
main.rs:154-276. All rendering goes through the C ABI (include/rr_ffi.h) to the CUDA kernels;
nothing here computes a pixel.

All host arithmetic that feeds the device (quaternion from pitch/yaw/roll, light normalisation,
yfov) is done in IEEE f32 in the reference's operation order, with sinf/cosf taken from the C
library exactly like Rust's f32::sin/cos on linux-gnu.
"""
import ctypes as C
import ctypes.util
import math

import numpy as np

from . import ffi

f32 = np.float32
_libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.sinf.restype = C.c_float
_libm.sinf.argtypes = [C.c_float]
_libm.cosf.restype = C.c_float
_libm.cosf.argtypes = [C.c_float]

PI = f32(math.pi)  # std::f32::consts::PI
MAX_REFLECTIONS = 3  # render.rs:11
MAX_REFRACTIONS = 10  # render.rs:12


def _v3(v):
    return (f32(v[0]), f32(v[1]), f32(v[2]))


class RenderColor:  # render.rs:23-42
    __slots__ = ("r", "g", "b")

    def __init__(self, r, g, b):
        self.r, self.g, self.b = f32(r), f32(g), f32(b)

    @staticmethod
    def new(r, g, b):
        return RenderColor(r, g, b)

    @staticmethod
    def zero():
        return RenderColor(0.0, 0.0, 0.0)

    def __iter__(self):
        return iter((self.r, self.g, self.b))

    def __repr__(self):
        return f"RenderColor({self.r}, {self.g}, {self.b})"


class Quat:  # quat.rs
    __slots__ = ("x", "y", "z", "w")

    def __init__(self, x, y, z, w):
        self.x, self.y, self.z, self.w = f32(x), f32(y), f32(z), f32(w)

    def mul(self, qb):  # quat.rs:63-72
        qa = self
        return Quat(
            qa.y * qb.z - qa.z * qb.y + qa.x * qb.w + qa.w * qb.x,
            qa.z * qb.x - qa.x * qb.z + qa.y * qb.w + qa.w * qb.y,
            qa.x * qb.y - qa.y * qb.x + qa.z * qb.w + qa.w * qb.z,
            -qa.x * qb.x - qa.y * qb.y - qa.z * qb.z + qa.w * qb.w,
        )

    @staticmethod
    def rotation(p, sx, sy, sz):  # quat.rs:92-95
        half = f32(p) / f32(2.0)
        ln = f32(_libm.sinf(float(half)))
        return Quat(ln * f32(sx), ln * f32(sy), ln * f32(sz), f32(_libm.cosf(float(half))))

    @staticmethod
    def from_pyr(pyr):  # quat.rs:129-134
        mx = Quat.rotation(pyr[2], 1.0, 0.0, 0.0)
        my = Quat.rotation(pyr[1], 0.0, 0.0, 1.0)
        mp = Quat.rotation(pyr[0], 0.0, 1.0, 0.0)
        return mx.mul(my).mul(mp)

    def as_tuple(self):
        return (self.x, self.y, self.z, self.w)


PATTERNS = {"Solid": ffi.RR_SOLID, "Checkerboard": ffi.RR_CHECKERBOARD, "RepeatedGradation": ffi.RR_REPEATED_GRADATION}
UVMAPS = {"XY": ffi.RR_UV_XY, "YZ": ffi.RR_UV_YZ, "ZX": ffi.RR_UV_ZX, "LL": ffi.RR_UV_LL}
FILTERS = {"Nearest": ffi.RR_NEAREST, "Bilinear": ffi.RR_BILINEAR}
_INV = lambda d: {v: k for k, v in d.items()}
PATTERN_NAMES, UVMAP_NAMES, FILTER_NAMES = _INV(PATTERNS), _INV(UVMAPS), _INV(FILTERS)


def _load_rgb8(path):
    """image::open(path) restricted to what the path honours: an RGB8 image (render.rs:251)."""
    try:
        from PIL import Image

        im = Image.open(path)
        if im.mode != "RGB":  # any other decoded format silently falls back to the pattern
            return None
        return np.ascontiguousarray(np.asarray(im, dtype=np.uint8))
    except Exception:
        return None


class RenderMaterial:  # render.rs:82-181
    def __init__(self, name, diffuse, specular, pn, t, n):
        self.name = str(name)
        self.diffuse = diffuse
        self.specular = specular
        self.pn = int(pn)
        self.t = f32(t)
        self.n = f32(n)
        self._glow_dist = f32(0.0)
        self._frac = RenderColor(1.0, 1.0, 1.0)
        self._pattern = "Solid"
        self._pattern_scale = f32(1.0)
        self._pattern_angle_scale = f32(1.0)
        self.texture_name = ""
        self._texture = None  # HxWx3 uint8 or None
        self.texture_filter = "Nearest"

    @staticmethod
    def new(name, diffuse, specular, pn, t, n):
        return RenderMaterial(name, diffuse, specular, pn, t, n)

    def get_name(self):
        return self.name

    def glow_dist(self, v):
        self._glow_dist = f32(v)
        return self

    def frac(self, frac):
        self._frac = frac
        return self

    def pattern(self, pattern):
        assert pattern in PATTERNS
        self._pattern = pattern
        return self

    def pattern_scale(self, v):
        self._pattern_scale = f32(v)
        return self

    def pattern_angle_scale(self, v):
        self._pattern_angle_scale = f32(v)
        return self

    def texture(self, texture_name):  # render.rs:165-174: error when the image cannot be opened
        self.texture_name = str(texture_name)
        self._texture = _load_rgb8(texture_name)
        if self._texture is None:
            raise IOError("texture image file load failed")
        return self

    def texture_ok(self, texture_name):  # render.rs:177-181: ignore quietly
        self.texture_name = str(texture_name)
        self._texture = _load_rgb8(texture_name)
        return self

    def texture_data(self, rgb8, filt="Nearest"):
        """Attach an in-memory RGB8 texture (tests)."""
        self._texture = np.ascontiguousarray(rgb8, dtype=np.uint8)
        self.texture_filter = filt
        return self

    def serialize(self):  # render.rs:183-199
        col = lambda c: {"r": float(c.r), "g": float(c.g), "b": float(c.b)}
        return {
            "name": self.name,
            "diffuse": col(self.diffuse),
            "specular": col(self.specular),
            "pn": self.pn,
            "t": float(self.t),
            "n": float(self.n),
            "glow_dist": float(self._glow_dist),
            "frac": col(self._frac),
            "pattern": self._pattern,
            "pattern_scale": float(self._pattern_scale),
            "pattern_angle_scale": float(self._pattern_angle_scale),
            "texture_name": self.texture_name,
            "texture_filter": self.texture_filter,
        }

    @staticmethod
    def deserialize(o):  # render.rs:201-218
        col = lambda c: RenderColor(c["r"], c["g"], c["b"])
        m = RenderMaterial(o["name"], col(o["diffuse"]), col(o["specular"]), o["pn"], o["t"], o["n"])
        m._glow_dist = f32(o["glow_dist"])
        m._frac = col(o["frac"])
        m._pattern = o["pattern"]
        m._pattern_scale = f32(o["pattern_scale"])
        m._pattern_angle_scale = f32(o["pattern_angle_scale"])
        m.texture_name = o["texture_name"]
        m._texture = _load_rgb8(o["texture_name"]) if o["texture_name"] else None
        m.texture_filter = o["texture_filter"]
        if m._pattern not in PATTERNS or m.texture_filter not in FILTERS:
            raise DeserializeError("serde_yaml::Error")
        return m


class DeserializeError(Exception):  # render.rs:341-366
    def __init__(self, s):
        super().__init__("Deserialize error: " + s)
        self.s = s


class RenderSphere:  # render.rs:378-427
    kind = ffi.RR_SPHERE

    def __init__(self, material, r, org):
        self.material = material
        self.r = f32(r)
        self.org = _v3(org)
        self._uvmap = "XY"

    @staticmethod
    def new(material, r, org):
        return RenderSphere(material, r, org)

    def uvmap(self, v):
        assert v in UVMAPS
        self._uvmap = v
        return self

    def serialize(self):
        x, y, z = self.org
        return {"Sphere": {"material": self.material.name, "r": float(self.r),
                           "org": {"x": float(x), "y": float(y), "z": float(z)}, "uvmap": self._uvmap}}


class RenderFloor:  # render.rs:487-537
    kind = ffi.RR_FLOOR

    def __init__(self, material, org, face_normal):
        self.material = material
        self.org = _v3(org)
        self.face_normal = _v3(face_normal)
        self.r = f32(0.0)
        self._uvmap = "XY"

    @staticmethod
    def new(material, org, face_normal):
        return RenderFloor(material, org, face_normal)

    new_raw = new

    def uvmap(self, v):
        assert v in UVMAPS
        self._uvmap = v
        return self

    def serialize(self):
        v = lambda p: {"x": float(p[0]), "y": float(p[1]), "z": float(p[2])}
        return {"Floor": {"material": self.material.name, "org": v(self.org),
                          "face_normal": v(self.face_normal), "uvmap": self._uvmap}}


class Camera:  # render.rs:617-622
    def __init__(self, position, pyr):
        self.position = _v3(position)
        self.pyr = _v3(pyr)
        self.rotation = Quat.from_pyr(self.pyr)


def _normalized(v):  # vec3.rs:36-39
    x, y, z = _v3(v)
    ln = np.sqrt(x * x + y * y + z * z)
    return (x / ln, y / ln, z / ln)


class RenderEnv:  # render.rs:646-799
    def __init__(self, cam, pyr, xres, yres, xfov, yfov, bgproc="bgcolor"):
        self.camera = Camera(cam, pyr)
        self.camera_motion = []
        self.xres, self.yres = int(xres), int(yres)
        self.xfov, self.yfov = f32(xfov), f32(yfov)
        self._materials = {}
        self._objects = []
        self._light = _v3((0.0, 0.0, 1.0))
        # bgproc is a host fn pointer in the reference (render.rs:661); only `bgcolor` exists.
        self.bgproc = bgproc
        self._use_raymarching = False
        self._glow_effect = None
        self.max_reflections = MAX_REFLECTIONS
        self.max_refractions = MAX_REFRACTIONS

    @staticmethod
    def new(cam, pyr, xres, yres, xfov, yfov, bgproc="bgcolor"):
        return RenderEnv(cam, pyr, xres, yres, xfov, yfov, bgproc)

    def materials(self, materials):
        self._materials = dict(materials)
        return self

    def objects(self, objects):
        self._objects = list(objects)
        return self

    def light(self, light):
        self._light = _normalized(light)
        return self

    def use_raymarching(self, f):
        self._use_raymarching = bool(f)
        return self

    def glow_effect(self, v):
        self._glow_effect = None if v is None else f32(v)
        return self

    # ---- YAML, render.rs:735-799 (schema: SURVEY.md appendix B) ----
    def serialize(self):
        import yaml

        v = lambda p: {"x": float(p[0]), "y": float(p[1]), "z": float(p[2])}
        mats = {}
        for o in self._objects:
            mats[o.material.name] = o.material.serialize()
        scene = {
            "camera": {"position": v(self.camera.position), "pyr": v(self.camera.pyr)},
            "camera_motion": [],
            "max_reflections": MAX_REFLECTIONS,  # the constants, not the env's values (render.rs:742-743)
            "max_refractions": MAX_REFRACTIONS,
            "materials": mats,
            "objects": [o.serialize() for o in self._objects],
        }
        return "---\n" + yaml.safe_dump(scene, sort_keys=False, default_flow_style=False)

    def deserialize(self, s):
        import yaml

        try:
            sc = yaml.safe_load(s)
            v = lambda p: (p["x"], p["y"], p["z"])
            mm = {k: RenderMaterial.deserialize(m) for k, m in sc["materials"].items()}
            cam = Camera(v(sc["camera"]["position"]), v(sc["camera"]["pyr"]))
            motion = list(sc["camera_motion"])
            max_refl, max_refr = int(sc["max_reflections"]), int(sc["max_refractions"])
            objs_serial = list(sc["objects"])
        except DeserializeError:
            raise
        except Exception:
            raise DeserializeError("serde_yaml::Error")
        self.camera = cam
        self.camera_motion = motion
        self.max_reflections, self.max_refractions = max_refl, max_refr
        self._materials = mm
        self._objects = []
        for o in objs_serial:
            (kind, body), = o.items()
            mat = self._materials.get(body["material"])
            if kind == "Sphere":
                if mat is None:
                    raise DeserializeError(f"RenderSphere couldn't find material {body['material']}")
                self._objects.append(RenderSphere(mat, body["r"], v(body["org"])).uvmap(body["uvmap"]))
            elif kind == "Floor":
                if mat is None:
                    raise DeserializeError(f"RenderFloor couldn't find material {body['material']}")
                self._objects.append(RenderFloor(mat, v(body["org"]), v(body["face_normal"])).uvmap(body["uvmap"]))
            else:
                raise DeserializeError("serde_yaml::Error")

    # ---- flatten to the C ABI ----
    def flatten(self):
        """-> FlatScene (rr_scene_desc + the arrays that back it)."""
        mats, mat_index, textures = [], {}, []
        for o in self._objects:
            m = o.material
            if id(m) not in mat_index:
                mat_index[id(m)] = len(mats)
                mats.append(m)
        cm = (ffi.rr_material * max(1, len(mats)))()
        keep = []
        for i, m in enumerate(mats):
            c = cm[i]
            c.diffuse[:] = [float(x) for x in m.diffuse]
            c.specular[:] = [float(x) for x in m.specular]
            c.pn, c.t, c.n, c.glow_dist = m.pn, float(m.t), float(m.n), float(m._glow_dist)
            c.frac[:] = [float(x) for x in m._frac]
            c.pattern = PATTERNS[m._pattern]
            c.pattern_scale, c.pattern_angle_scale = float(m._pattern_scale), float(m._pattern_angle_scale)
            c.texture_filter = FILTERS[m.texture_filter]
            if m._texture is not None:
                c.texture = len(textures)
                textures.append(m._texture)
            else:
                c.texture = -1
        co = (ffi.rr_object * max(1, len(self._objects)))()
        for i, o in enumerate(self._objects):
            c = co[i]
            c.kind, c.material, c.uvmap, c.r = o.kind, mat_index[id(o.material)], UVMAPS[o._uvmap], float(o.r)
            c.org[:] = [float(x) for x in o.org]
            if o.kind == ffi.RR_FLOOR:
                c.face_normal[:] = [float(x) for x in o.face_normal]
        ct = (ffi.rr_texture * max(1, len(textures)))()
        for i, t in enumerate(textures):
            ct[i].height, ct[i].width = t.shape[0], t.shape[1]
            ct[i].rgb8 = t.ctypes.data_as(C.POINTER(C.c_uint8))
            keep.append(t)
        desc = ffi.rr_scene_desc(len(self._objects), co, len(mats), cm, len(textures), ct)
        return FlatScene(desc, [cm, co, ct, keep])

    def frame_params(self, band_rows=0, band_index=0, band_count=1, band_span=1):
        p = ffi.rr_frame_params()
        p.xres, p.yres, p.xfov, p.yfov = self.xres, self.yres, float(self.xfov), float(self.yfov)
        p.cam_position[:] = [float(x) for x in self.camera.position]
        p.cam_rotation[:] = [float(x) for x in self.camera.rotation.as_tuple()]
        p.light[:] = [float(x) for x in self._light]
        p.use_raymarching = 1 if self._use_raymarching else 0
        p.glow_enabled = 0 if self._glow_effect is None else 1
        p.glow_effect = 0.0 if self._glow_effect is None else float(self._glow_effect)
        p.max_reflections, p.max_refractions = self.max_reflections, self.max_refractions
        p.bg_kind = ffi.RR_BG_BGCOLOR if self.bgproc == "bgcolor" else ffi.RR_BG_BLACK
        p.band_rows, p.band_index, p.band_count, p.band_span = band_rows, band_index, band_count, band_span
        return p


class FlatScene:
    def __init__(self, desc, keep):
        self.desc = desc
        self._keep = keep


def frame_rows(params):
    """Rows a call with `params` produces (whole frame, or this shard's bands)."""
    cnt = max(1, params.band_count)
    if cnt == 1:
        return params.yres
    br = max(1, params.band_rows)
    span = max(1, params.band_span)
    return sum(1 for iy in range(params.yres) if params.band_index <= (iy // br) % cnt < params.band_index + span)


class DeviceScene:
    """Owns an rr_scene handle (device-resident flattened scene)."""

    def __init__(self, ren, device=0):
        self.lib = ffi.load()
        self.flat = ren.flatten()
        h = C.c_void_p()
        ffi.check(self.lib.rr_scene_create(C.byref(self.flat.desc), device, C.byref(h)))
        self.handle = h
        self.device = device

    def close(self):
        if self.handle:
            self.lib.rr_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_culling(self, enabled):
        """enabled=False forces the brute-force object scan (tests compare it with the BVH path)."""
        ffi.check(self.lib.rr_scene_set_culling(self.handle, 1 if enabled else 0))

    def render_rgb8(self, params, out=None):
        rows = frame_rows(params)
        if out is None:
            out = np.empty((rows, params.xres, 3), dtype=np.uint8)
        ffi.check(self.lib.rr_render_rgb8(self.handle, C.byref(params), out.ctypes.data_as(C.c_void_p), 0))
        return out

    def render_f32(self, params):
        rows = frame_rows(params)
        out = np.empty((rows, params.xres, 3), dtype=np.float32)
        ffi.check(self.lib.rr_render_f32(self.handle, C.byref(params), out.ctypes.data_as(C.c_void_p)))
        return out

    def render_count(self, params, want_image=True):
        rows = frame_rows(params)
        out = np.empty((rows, params.xres, 3), dtype=np.uint8) if want_image else None
        counts = ffi.rr_ray_counts()
        ptr = out.ctypes.data_as(C.c_void_p) if out is not None else None
        ffi.check(self.lib.rr_render_count(self.handle, C.byref(params), ptr, 0, C.byref(counts)))
        return out, counts

    def render_rgb8_device(self, params, d_ptr, stream=None, row_stride=0):
        ffi.check(self.lib.rr_render_rgb8_device(self.handle, C.byref(params), C.c_void_p(d_ptr), row_stride,
                                                 C.c_void_p(stream) if stream else None))

    def render_f32_device(self, params, d_ptr, stream=None):
        ffi.check(self.lib.rr_render_f32_device(self.handle, C.byref(params), C.c_void_p(d_ptr),
                                                C.c_void_p(stream) if stream else None))

    def last_kernel_ms(self):
        ms = C.c_float()
        ffi.check(self.lib.rr_last_kernel_ms(self.handle, C.byref(ms)))
        return float(ms.value)


def render(ren, pointproc, thread_count=1, device=0):
    """render(), render.rs:801-805: calls pointproc(x, y, RenderColor) once per pixel, row-major.

    `thread_count` is kept for signature compatibility; the device grid replaces the reference's
    row-scheduler threads (render.rs:836-898), so it is ignored.
    """
    scene = DeviceScene(ren, device)
    try:
        img = scene.render_f32(ren.frame_params())
    finally:
        scene.close()
    for y in range(ren.yres):
        row = img[y]
        for x in range(ren.xres):
            pointproc(x, y, RenderColor(row[x, 0], row[x, 1], row[x, 2]))


# ---------------------------------------------------------------------------------------------
# built-in scene of main.rs:154-276
# ---------------------------------------------------------------------------------------------
def default_scene(width, height, use_raymarching=False, glow_effect=None):
    xfov = f32(1.0)
    yfov = f32(height) / f32(width)  # main.rs:135-136
    floor_material = (
        RenderMaterial.new("floor", RenderColor(1.0, 1.0, 0.0), RenderColor(0.0, 0.0, 0.0), 0, 0.0, 0.0)
        .pattern("RepeatedGradation").pattern_scale(300.0).pattern_angle_scale(0.2).texture_ok("bar.png")
    )
    mirror = RenderMaterial.new("mirror", RenderColor(0.0, 0.0, 0.0), RenderColor(1.0, 1.0, 1.0), 24, 0.0, 0.0).frac(
        RenderColor(1.0, 1.0, 1.0))
    red = RenderMaterial.new("red", RenderColor(0.8, 0.0, 0.0), RenderColor(0.0, 0.0, 0.0), 24, 0.0, 0.0).glow_dist(5.0)
    transparent = RenderMaterial.new("transparent", RenderColor(0.0, 0.0, 0.0), RenderColor(0.0, 0.0, 0.0), 0, 1.0,
                                     1.5).frac(RenderColor(1.49998, 1.49999, 1.5))
    materials = {"floor": floor_material}
    objects = [
        RenderFloor.new_raw(floor_material, (0.0, -300.0, 0.0), (0.0, 1.0, 0.0)).uvmap("ZX"),
        RenderSphere.new(mirror, 80.0, (0.0, -30.0, 172.0)),
        RenderSphere.new(mirror, 80.0, (-200.0, -30.0, 172.0)),
        RenderSphere.new(red, 80.0, (-200.0, -200.0, 172.0)),
        RenderSphere.new(transparent, 100.0, (70.0, -200.0, 150.0)),
    ]
    return (
        RenderEnv.new((0.0, -150.0, -300.0), (f32(0.0), -PI / f32(2.0), -PI / f32(2.0)), width, height, xfov, yfov)
        .materials(materials).objects(objects).light((50.0, 60.0, -50.0))
        .use_raymarching(use_raymarching).glow_effect(glow_effect)
    )


# ---------------------------------------------------------------------------------------------
# synthetic scene of BASELINE.json configs[3] (recipe: SURVEY.md 8d "Config 4")
# ---------------------------------------------------------------------------------------------
class SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def unit(self):
        """24-bit uniform in [0,1), exactly representable in f32."""
        return f32(self.next() >> 40) * f32(2.0 ** -24)

    def uniform(self, a, b):
        a, b = f32(a), f32(b)
        return a + (b - a) * self.unit()

    def below(self, n):
        return int(self.next() % n)


def synthetic_scene(width, height, n_spheres=1024, seed=20261018, use_raymarching=False, glow_effect=None):
    rng = SplitMix64(seed)
    floor_material = (
        RenderMaterial.new("floor", RenderColor(1.0, 1.0, 0.0), RenderColor(0.0, 0.0, 0.0), 0, 0.0, 0.0)
        .pattern("RepeatedGradation").pattern_scale(300.0).pattern_angle_scale(0.2)
    )
    mats = []
    for i in range(6):  # mirrors
        s = rng.uniform(0.5, 1.0)
        d = [rng.uniform(0.0, 0.3) for _ in range(3)]
        mats.append(RenderMaterial.new(f"mirror{i}", RenderColor(*d), RenderColor(s, s, s), 24, 0.0, 0.0))
    for i in range(5):  # diffuse
        d = [rng.uniform(0.1, 1.0) for _ in range(3)]
        m = RenderMaterial.new(f"diffuse{i}", RenderColor(*d), RenderColor(0.0, 0.0, 0.0), 24, 0.0, 0.0)
        mats.append(m.pattern("Solid" if i % 2 == 0 else "Checkerboard").pattern_scale(10.0))
    for i in range(5):  # glass
        t = rng.uniform(0.5, 1.0)
        n = rng.uniform(1.2, 1.8)
        if i >= 3:
            s = rng.uniform(0.2, 0.5)
            spec = RenderColor(s, s, s)
        else:
            spec = RenderColor(0.0, 0.0, 0.0)
        mats.append(RenderMaterial.new(f"glass{i}", RenderColor(0.0, 0.0, 0.0), spec, 0, t, n))
    objects = [RenderFloor.new_raw(floor_material, (0.0, -300.0, 0.0), (0.0, 1.0, 0.0)).uvmap("ZX")]
    for _ in range(n_spheres):
        m = mats[rng.below(len(mats))]
        r = rng.uniform(15.0, 45.0)
        x = rng.uniform(-900.0, 900.0)
        y = rng.uniform(-280.0, 300.0)
        z = rng.uniform(-100.0, 1900.0)
        objects.append(RenderSphere.new(m, r, (x, y, z)))
    materials = {m.name: m for m in [floor_material] + mats}
    xfov = f32(1.0)
    yfov = f32(height) / f32(width)
    return (
        RenderEnv.new((0.0, -150.0, -300.0), (f32(0.0), -PI / f32(2.0), -PI / f32(2.0)), width, height, xfov, yfov)
        .materials(materials).objects(objects).light((50.0, 60.0, -50.0))
        .use_raymarching(use_raymarching).glow_effect(glow_effect)
    )
