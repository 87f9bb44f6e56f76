"""CPU tests of the C++ host layer (ray-rust_b200/host): scene model, YAML, PNG, CLI error path.
It must agree with the Python binding bit for bit, because both feed the same C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ray-rust_b200", "host")


@pytest.fixture(scope="module")
def host(rr):
    subprocess.check_call(["make", "-C", HOST, "-s"])
    lib = C.CDLL(os.path.join(HOST, "libray_rust_host.so"))
    lib.rrh_env_new.restype = C.c_void_p
    lib.rrh_env_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64]
    lib.rrh_env_free.argtypes = [C.c_void_p]
    lib.rrh_env_serialize.restype = C.c_void_p
    lib.rrh_env_serialize.argtypes = [C.c_void_p]
    lib.rrh_free.argtypes = [C.c_void_p]
    lib.rrh_env_deserialize.argtypes = [C.c_void_p, C.c_char_p]
    lib.rrh_env_flatten.argtypes = [C.c_void_p, C.POINTER(rr.ffi.rr_scene_desc), C.POINTER(rr.ffi.rr_frame_params)]
    lib.rrh_env_limits.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.rrh_last_error.restype = C.c_char_p
    lib.rrh_png_roundtrip.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_void_p]
    lib.rrh_load_image.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p, C.c_uint64]
    return lib


def _serialize(host, h):
    p = host.rrh_env_serialize(h)
    s = C.string_at(p).decode()
    host.rrh_free(p)
    return s


def _flat_bytes(rr, desc, params):
    objs = C.string_at(desc.objects, C.sizeof(rr.ffi.rr_object) * desc.n_objects)
    mats = C.string_at(desc.materials, C.sizeof(rr.ffi.rr_material) * desc.n_materials)
    return objs, mats, bytes(params)


def _host_flat(rr, host, h):
    d, p = rr.ffi.rr_scene_desc(), rr.ffi.rr_frame_params()
    host.rrh_env_flatten(h, C.byref(d), C.byref(p))
    return _flat_bytes(rr, d, p)


def _py_flat(rr, ren):
    flat = ren.flatten()
    return _flat_bytes(rr, flat.desc, ren.frame_params()), flat


@pytest.mark.parametrize("march,glow", [(0, None), (1, 1.0)])
def test_default_scene_cpp_equals_python(rr, host, march, glow):
    h = host.rrh_env_new(0, 640, 480, march, 0 if glow is None else 1, glow or 0.0, 0, 0)
    (po, pm, pp), _keep = _py_flat(rr, rr.default_scene(640, 480, use_raymarching=bool(march), glow_effect=glow))
    co, cm, cp = _host_flat(rr, host, h)
    host.rrh_env_free(h)
    assert co == po and cm == pm and cp == pp


def test_synthetic_scene_cpp_equals_python(rr, host):
    h = host.rrh_env_new(1, 3840, 2160, 0, 0, 0.0, 1024, 20261018)
    (po, pm, pp), _keep = _py_flat(rr, rr.synthetic_scene(3840, 2160))
    co, cm, cp = _host_flat(rr, host, h)
    host.rrh_env_free(h)
    assert co == po and cm == pm and cp == pp


def test_yaml_cpp_writer_loads_in_pyyaml_and_python_binding(rr, host):
    h = host.rrh_env_new(1, 320, 200, 0, 0, 0.0, 40, 7)
    text = _serialize(host, h)
    assert text.startswith("---\n")
    doc = yaml.safe_load(text)
    assert doc["camera_motion"] == [] and doc["max_reflections"] == 3 and doc["max_refractions"] == 10
    assert len(doc["objects"]) == 41 and "Floor" in doc["objects"][0] and "Sphere" in doc["objects"][1]
    assert set(doc["materials"]["floor"]) == {"name", "diffuse", "specular", "pn", "t", "n", "glow_dist", "frac", "pattern",
                                              "pattern_scale", "pattern_angle_scale", "texture_name", "texture_filter"}
    ren = rr.default_scene(320, 200)
    ren.deserialize(text)                      # Python reader takes the C++ writer's output
    (po, pm, pp), _keep = _py_flat(rr, ren)
    co, cm, cp = _host_flat(rr, host, h)
    host.rrh_env_free(h)
    assert po == co and pp == cp
    assert sorted(pm[i:i + 72] for i in range(0, len(pm), 72)) == sorted(cm[i:i + 72] for i in range(0, len(cm), 72))


def test_yaml_cpp_reader_takes_python_writer_and_roundtrips(rr, host):
    src = rr.synthetic_scene(64, 48, n_spheres=30, seed=99)
    text = src.serialize()                     # PyYAML block style
    h = host.rrh_env_new(0, 64, 48, 0, 0, 0.0, 0, 0)
    assert host.rrh_env_deserialize(h, text.encode()) == 0, host.rrh_last_error()
    (po, _pm, pp), _keep = _py_flat(rr, src)
    co, _cm, cp = _host_flat(rr, host, h)
    assert po == co and pp == cp
    text2 = _serialize(host, h)
    h2 = host.rrh_env_new(0, 64, 48, 0, 0, 0.0, 0, 0)
    assert host.rrh_env_deserialize(h2, text2.encode()) == 0
    assert _host_flat(rr, host, h2) == _host_flat(rr, host, h)
    host.rrh_env_free(h)
    host.rrh_env_free(h2)


HAND_WRITTEN = """
# flow style, quoted strings, widened floats, .inf, ~ and a camera key frame
camera: {position: {x: 1, y: -2.5, z: 3e2}, pyr: {x: 0.0, y: -1.5707963705062866, z: 0}}
camera_motion:
  - camera:
      position: {x: 0, y: 0, z: 0}
      pyr: {x: 0, y: 0, z: 0}
    velocity: {x: 1, y: 0, z: 0}
    camera_target: ~
    duration: 2.0
  - camera: {position: {x: 5, y: 0, z: 0}, pyr: {x: 0, y: 1, z: 0}}
    velocity: {x: 0, y: 0, z: 0}
    camera_target: {x: 0, y: -30, z: 172}
    duration: 1.5
max_reflections: 5
max_refractions: 4
materials:
  "m 1":
    name: "m 1"
    diffuse: {r: 0.800000011920929, g: 0.0, b: 0.0}
    specular: {r: 0, g: 0, b: 0}
    pn: 24
    t: 0.0
    n: 0.0
    glow_dist: .inf
    frac: {r: 1.0, g: 1.0, b: 1.0}
    pattern: Checkerboard
    pattern_scale: 10
    pattern_angle_scale: 1
    texture_name: ''
    texture_filter: Bilinear
objects:
- Sphere:
    material: m 1
    r: 12.5
    org: {x: 1, y: 2, z: 3}
    uvmap: LL
- Floor: {material: "m 1", org: {x: 0, y: -300, z: 0}, face_normal: {x: 0, y: 1, z: 0}, uvmap: ZX}
"""


def test_yaml_hand_written_subset(rr, host):
    h = host.rrh_env_new(0, 32, 32, 0, 0, 0.0, 0, 0)
    assert host.rrh_env_deserialize(h, HAND_WRITTEN.encode()) == 0, host.rrh_last_error()
    d, p = rr.ffi.rr_scene_desc(), rr.ffi.rr_frame_params()
    host.rrh_env_flatten(h, C.byref(d), C.byref(p))
    assert d.n_objects == 2 and d.n_materials == 1
    assert d.objects[0].kind == rr.ffi.RR_SPHERE and d.objects[0].uvmap == rr.ffi.RR_UV_LL and d.objects[0].r == 12.5
    assert d.objects[1].kind == rr.ffi.RR_FLOOR and list(d.objects[1].face_normal) == [0.0, 1.0, 0.0]
    m = d.materials[0]
    assert m.diffuse[0] == np.float32(0.8) and m.pattern == rr.ffi.RR_CHECKERBOARD and m.texture_filter == rr.ffi.RR_BILINEAR
    assert m.glow_dist == float("inf") and m.texture == -1
    assert list(p.cam_position) == [1.0, -2.5, 300.0] and p.max_reflections == 5 and p.max_refractions == 4
    a, b, k = C.c_int(), C.c_int(), C.c_int()
    host.rrh_env_limits(h, C.byref(a), C.byref(b), C.byref(k))
    assert (a.value, b.value, k.value) == (5, 4, 2)
    # same text through the Python binding gives the same flattened scene
    ren = rr.default_scene(32, 32)
    ren.deserialize(HAND_WRITTEN)
    (po, pm, _pp), _keep = _py_flat(rr, ren)
    co, cm, _cp = _host_flat(rr, host, h)
    host.rrh_env_free(h)
    assert po == co and pm == cm


def test_yaml_errors(rr, host):
    h = host.rrh_env_new(0, 8, 8, 0, 0, 0.0, 0, 0)
    good = _serialize(host, h)
    assert host.rrh_env_deserialize(h, good.replace("material: red", "material: nosuch").encode()) == -1
    assert b"Deserialize error: RenderSphere couldn't find material nosuch" in host.rrh_last_error()
    assert host.rrh_env_deserialize(h, good.replace("material: floor", "material: gone", 1).encode()) == -1
    assert b"RenderFloor couldn't find material gone" in host.rrh_last_error()
    for bad in ("camera: 3", good.replace("max_reflections: 3\n", ""), good.replace("pattern: Solid", "pattern: Plaid"),
                good.replace("- Sphere:", "- Cube:", 1), "a: [1, 2"):
        assert host.rrh_env_deserialize(h, bad.encode()) == -1
        assert host.rrh_last_error() == b"Deserialize error: serde_yaml::Error"
    host.rrh_env_free(h)


def test_png_writer_and_reader(host, tmp_path):
    from PIL import Image

    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    back = np.zeros_like(img)
    path = str(tmp_path / "a.png")
    assert host.rrh_png_roundtrip(img.ctypes.data_as(C.c_void_p), 53, 37, path.encode(), back.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(back, img)
    assert np.array_equal(np.asarray(Image.open(path).convert("RGB")), img)   # a conforming decoder agrees
    # reader on files written by another encoder (filters 1-4 in use), and the not-RGB8 cases
    path2 = str(tmp_path / "b.png")
    grad = np.add.outer(np.arange(40), np.arange(60))[:, :, None].repeat(3, 2).astype(np.uint8)
    Image.fromarray(grad).save(path2, optimize=True)
    lib = host
    lib.rrh_env_new.restype = C.c_void_p
    # load through a material: use the CLI-independent helper by round-tripping the bytes
    import zlib  # noqa: F401
    # decode via the facade: save_png is not involved, only load_png
    buf = np.zeros_like(grad)
    # rrh_png_roundtrip rewrites the file, so call the reader through a tiny scene YAML instead
    h = host.rrh_env_new(0, 8, 8, 0, 0, 0.0, 0, 0)
    text = _serialize(host, h).replace('texture_name: bar.png', f'texture_name: {path2}')
    assert host.rrh_env_deserialize(h, text.encode()) == 0
    from ray_rust_b200 import ffi
    d, p = ffi.rr_scene_desc(), ffi.rr_frame_params()
    host.rrh_env_flatten(h, C.byref(d), C.byref(p))
    assert d.n_textures == 1 and (d.textures[0].width, d.textures[0].height) == (60, 40)
    got = np.ctypeslib.as_array(d.textures[0].rgb8, shape=(40, 60, 3))
    assert np.array_equal(got, grad)
    rgba = str(tmp_path / "c.png")
    Image.fromarray(np.dstack([grad, grad[:, :, :1]])).save(rgba)
    assert host.rrh_env_deserialize(h, text.replace(path2, rgba).encode()) == 0
    host.rrh_env_flatten(h, C.byref(d), C.byref(p))
    assert d.n_textures == 0   # RGBA is not ImageRgb8: silently falls back to the pattern (render.rs:251)
    host.rrh_env_free(h)


def _load_image(host, path):
    w, h = C.c_uint32(), C.c_uint32()
    rc = host.rrh_load_image(str(path).encode(), C.byref(w), C.byref(h), None, 0)
    if rc != 0:
        return rc
    out = np.zeros((h.value, w.value, 3), np.uint8)
    assert host.rrh_load_image(str(path).encode(), C.byref(w), C.byref(h), out.ctypes.data_as(C.c_void_p), out.size) == 0
    return out


def test_jpeg_textures_decode_like_a_conforming_decoder(host, tmp_path):
    """image::open() loads JPEG textures too (render.rs:165-181). The host decodes baseline and progressive JPEG itself (rr_jpeg.cpp); T.81
    leaves IDCT and chroma upsampling to the decoder within a level or so, hence the comparison with libjpeg-turbo (PIL) is
    a tolerance, not equality: every texel within 3 levels, 99 % within 1 (4:4:4, no upsampling: within 2)."""
    from PIL import Image

    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:77, 0:101]
    smooth = np.stack([128 + 100 * np.sin(xx / 9.0) * np.cos(yy / 13.0), 128 + 90 * np.cos(xx / 17.0 + yy / 5.0), (xx * 2 + yy) % 256], -1)
    img = np.clip(smooth + rng.normal(0, 6, smooth.shape), 0, 255).astype(np.uint8)
    for name, kw, tol in (("444", dict(subsampling=0), 2), ("422", dict(subsampling=1), 3), ("420", dict(subsampling=2), 3),
                          ("420q50", dict(subsampling=2, quality=50), 3), ("opt", dict(subsampling=0, optimize=True, quality=95), 2)):
        path = tmp_path / f"t_{name}.jpg"
        Image.fromarray(img).save(path, "JPEG", quality=kw.pop("quality", 90), **kw)
        ref = np.asarray(Image.open(path).convert("RGB")).astype(int)
        got = _load_image(host, path)
        assert isinstance(got, np.ndarray) and got.shape == ref.shape, (name, got)
        d = np.abs(got.astype(int) - ref)
        assert d.max() <= tol, (name, int(d.max()))
        assert (d <= 1).mean() >= 0.99, (name, float((d <= 1).mean()))
    # restart intervals, and sizes that are not a multiple of the MCU
    small = img[:19, :23]
    path = tmp_path / "rst.jpg"
    Image.fromarray(small).save(path, "JPEG", quality=92, subsampling=2, restart_marker_blocks=1)
    assert b"\xff\xdd" in path.read_bytes()  # DRI present
    got, ref = _load_image(host, path), np.asarray(Image.open(path).convert("RGB")).astype(int)
    assert np.abs(got.astype(int) - ref).max() <= 3
    # what image::open() does not turn into ImageRgb8, or this decoder does not cover: quietly "no texture"
    grey = tmp_path / "grey.jpg"
    Image.fromarray(img[:, :, 0]).save(grey, "JPEG")
    assert _load_image(host, grey) == -2           # ImageLuma8: ignored by render.rs:251
    # progressive frames (spectral selection + successive approximation; libjpeg's default scan script), also with restart
    # intervals and subsampled chroma
    for name, kw in (("p444", dict(subsampling=0)), ("p420", dict(subsampling=2, quality=80)),
                     ("p422rst", dict(subsampling=1, restart_marker_blocks=3)), ("p420small", dict(subsampling=2))):
        path = tmp_path / f"{name}.jpg"
        src = small if name == "p420small" else img
        Image.fromarray(src).save(path, "JPEG", progressive=True, quality=kw.pop("quality", 90), **kw)
        assert b"\xff\xc2" in path.read_bytes()    # SOF2
        ref = np.asarray(Image.open(path).convert("RGB")).astype(int)
        got = _load_image(host, path)
        assert isinstance(got, np.ndarray) and got.shape == ref.shape, (name, got)
        d = np.abs(got.astype(int) - ref)
        assert d.max() <= 3 and (d <= 1).mean() >= 0.99, (name, int(d.max()), float((d <= 1).mean()))
    trunc = tmp_path / "trunc.jpg"
    trunc.write_bytes((tmp_path / "t_420.jpg").read_bytes()[:300])
    assert _load_image(host, trunc) == -2
    assert _load_image(host, tmp_path / "missing.jpg") == -2
    # through the scene YAML, like a reference scene that names a .jpg texture
    h = host.rrh_env_new(0, 8, 8, 0, 0, 0.0, 0, 0)
    text = _serialize(host, h).replace('texture_name: bar.png', f'texture_name: {tmp_path / "t_444.jpg"}')
    assert host.rrh_env_deserialize(h, text.encode()) == 0
    from ray_rust_b200 import ffi
    d, p = ffi.rr_scene_desc(), ffi.rr_frame_params()
    host.rrh_env_flatten(h, C.byref(d), C.byref(p))
    assert d.n_textures == 1 and (d.textures[0].width, d.textures[0].height) == (101, 77)
    host.rrh_env_free(h)


def test_cli_has_no_cpu_fallback(host, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([os.path.join(HOST, "ray-rust"), "16", "12", "-o", str(tmp_path / "x.png")], capture_output=True, text=True)
    assert r.returncode == 1
    assert "Value for width: 16" in r.stdout and "Value for threads: 8" in r.stdout and "Value for output:" in r.stdout
    assert "Error:" in r.stderr and not (tmp_path / "x.png").exists()
    r = subprocess.run([os.path.join(HOST, "ray-rust"), "16"], capture_output=True, text=True)
    assert r.returncode == 2 and "USAGE" in r.stderr
