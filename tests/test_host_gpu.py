"""GPU tests of the C++ host layer and the `ray-rust` CLI: the reference-shaped entry points
(render(ren, pointproc, threads), the CLI flags, -s/-d) drive the CUDA path and agree with the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ray-rust_b200", "host")
CLI = os.path.join(HOST, "ray-rust")
HOSTLIB = os.path.join(HOST, "libray_rust_host.so")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-C", HOST, "-s"])


def _png(path):
    from PIL import Image

    return np.asarray(Image.open(path).convert("RGB"))


def _close(dev, ref, min_exact=0.999):
    d = np.abs(dev.astype(int) - ref.astype(int)).max(axis=2)
    assert (d <= 1).mean() >= 0.9995 and (d == 0).mean() >= min_exact, ((d == 0).mean(), (d <= 1).mean(), d.max())


def test_cli_default_scene(rr, oracle, tmp_path):
    out = tmp_path / "foo.png"
    r = subprocess.run([CLI, "320", "240", "-t", "4", "-o", str(out)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    for line in ("Value for width: 320", "Value for height: 240", "Value for threads: 4", "Rendering time: "):
        assert line in r.stdout
    _close(_png(out), oracle.render(rr.default_scene(320, 240))["u8"])


def test_cli_raymarch_glow(rr, oracle, tmp_path):
    out = tmp_path / "m.png"
    r = subprocess.run([CLI, "160", "120", "-m", "-g", "1.0", "-o", str(out)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    assert "Value for gloweffect: 1" in r.stdout
    _close(_png(out), oracle.render(rr.default_scene(160, 120, use_raymarching=True, glow_effect=1.0))["u8"], 0.998)


def test_cli_serialize_then_deserialize(rr, oracle, tmp_path):
    """BASELINE config 4 flow: a scene written with -s and reloaded with -d renders the same frame."""
    ren = rr.synthetic_scene(200, 112, n_spheres=200)
    scene = tmp_path / "scene.yaml"
    scene.write_text(ren.serialize())
    a, b, again = tmp_path / "a.png", tmp_path / "b.png", tmp_path / "again.yaml"
    r = subprocess.run([CLI, "200", "112", "-d", str(scene), "-s", str(again), "-o", str(a)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([CLI, "200", "112", "-d", str(again), "-o", str(b)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(_png(a), _png(b))
    _close(_png(a), oracle.render(ren, threads=os.cpu_count())["u8"], 0.998)
    bad = tmp_path / "bad.yaml"
    bad.write_text(ren.serialize().replace("material: glass0", "material: nosuch"))
    r = subprocess.run([CLI, "8", "8", "-d", str(bad)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "Deserialize error: RenderSphere couldn't find material nosuch" in r.stderr


def test_cpp_render_pointproc(rr, oracle):
    lib = C.CDLL(os.path.join(HOST, "libray_rust_host.so"))
    lib.rrh_env_new.restype = C.c_void_p
    lib.rrh_env_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64]
    CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p)
    lib.rrh_render.argtypes = [C.c_void_p, CB, C.c_void_p, C.c_int]
    lib.rrh_render_rgb8.argtypes = [C.c_void_p, C.c_void_p]
    lib.rrh_env_free.argtypes = [C.c_void_p]
    w, h = 48, 32
    env = lib.rrh_env_new(0, w, h, 0, 0, 0.0, 0, 0)
    got = np.zeros((h, w, 3), dtype=np.float32)
    order = []

    def cb(x, y, r, g, b, _):
        got[y, x] = (r, g, b)
        order.append((y, x))

    assert lib.rrh_render(env, CB(cb), None, 8) == 0
    assert order == [(y, x) for y in range(h) for x in range(w)]  # once per pixel, row-major
    ref = oracle.render(rr.default_scene(w, h), want_f32=True, want_tags=True)
    clean = (ref["tags"] & 1) == 0
    assert np.array_equal(got.view(np.uint32)[clean], ref["f32"].view(np.uint32)[clean])
    img = np.zeros((h, w, 3), dtype=np.uint8)
    assert lib.rrh_render_rgb8(env, img.ctypes.data_as(C.c_void_p)) == 0
    d = np.abs(img.astype(int) - ref["u8"].astype(int)).max()
    assert d <= 1
    lib.rrh_env_free(env)


def test_cli_camera_motion_render_frames(rr, oracle, tmp_path):
    """render_frames (render.rs:926-989): a scene file with camera key frames renders one PNG per frame
    (`<output><i>.png`), camera interpolated with the reference's Hermite / slerp / look-at rules."""
    import ctypes as C
    import yaml

    ren = rr.default_scene(160, 90)
    doc = yaml.safe_load(ren.serialize())
    cam0 = doc["camera"]
    doc["camera_motion"] = [
        {"camera": {"position": {"x": 40.0, "y": -120.0, "z": -280.0}, "pyr": {"x": 0.1, "y": -1.4, "z": -1.5707964}},
         "velocity": {"x": 10.0, "y": 0.0, "z": 5.0}, "camera_target": None, "duration": 1.0},
        {"camera": {"position": {"x": 80.0, "y": -100.0, "z": -260.0}, "pyr": {"x": 0.0, "y": -1.2, "z": -1.5707964}},
         "velocity": {"x": 0.0, "y": 0.0, "z": 0.0}, "camera_target": {"x": 0.0, "y": -30.0, "z": 172.0}, "duration": 1.0},
    ]
    scene = tmp_path / "motion.yaml"
    scene.write_text(yaml.safe_dump(doc))
    r = subprocess.run([CLI, "160", "90", "-d", str(scene), "-o", str(tmp_path / "f")], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    assert "keyframe 0 / 2" in r.stdout and "Rendering frame 3" in r.stdout
    frames = [_png(tmp_path / f"f{i}.png") for i in range(4)]  # 2 key frames x (duration 1.0 / frame_step 0.5)
    assert not (tmp_path / "f4.png").exists()

    # oracle side: interpolate the camera exactly like render.rs:907-970 (f32) and render each frame
    f32 = np.float32
    lib = oracle.load()
    import ctypes.util

    libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.atan2f.restype = C.c_float
    libm.atan2f.argtypes = [C.c_float, C.c_float]

    def hermite(t, x0, x1, v0, v1):
        h = f32(1.0)
        d, c = x0, v0
        r_ = x1 - x0 - h * v0
        s = v1 - v0
        a = (h * s - f32(2.0) * r_) / h / h / h
        b = (-h * s + f32(3.0) * r_) / h / h
        return a * t * t * t + b * t * t + c * t + d

    def v3(d):
        return [f32(d["x"]), f32(d["y"]), f32(d["z"])]

    def from_pyr(p):
        out = (C.c_float * 4)()
        lib.oracle_quat_from_pyr(oracle.fa(*p), out)
        return [f32(x) for x in out]

    prev_pos, prev_rot, prev_vel = v3(cam0["position"]), from_pyr(v3(cam0["pyr"])), [f32(0)] * 3
    k = 0
    for key in doc["camera_motion"]:
        kpos, krot, v1 = v3(key["camera"]["position"]), from_pyr(v3(key["camera"]["pyr"])), v3(key["velocity"])
        for i in range(2):
            f = f32(i) / (f32(key["duration"]) / f32(0.5))
            pos = [hermite(f, prev_pos[c], kpos[c], prev_vel[c], v1[c]) for c in range(3)]
            if key["camera_target"] is None:
                out = (C.c_float * 4)()
                lib.oracle_quat_slerp(oracle.fa(*prev_rot), oracle.fa(*krot), float(f), out)
                rot = [f32(x) for x in out]
            else:
                # look-at, render.rs:961-967 restated in f32 with libm's atan2f/sinf/cosf (what Rust's f32 methods call)
                tgt = v3(key["camera_target"])
                dx, dy, dz = tgt[0] - pos[0], tgt[1] - pos[1], tgt[2] - pos[2]
                pitch = f32(libm.atan2f(float(dy), float(np.sqrt(dx * dx + dz * dz))))
                yaw = -f32(libm.atan2f(float(dz), float(dx)))
                q = rr.Quat.rotation(yaw, 0.0, 1.0, 0.0).mul(rr.Quat.rotation(pitch, 0.0, 0.0, 1.0)).mul(
                    rr.Quat.rotation(-f32(np.pi) / f32(2.0), 1.0, 0.0, 0.0))
                rot = list(q.as_tuple())
            e = rr.default_scene(160, 90)
            e.camera.position = tuple(pos)
            e.camera.rotation = rr.Quat(*rot)
            _close(frames[k], oracle.render(e)["u8"], 0.999)
            k += 1
        prev_pos, prev_rot, prev_vel = kpos, krot, v1
    assert not np.array_equal(frames[0], frames[1])


def test_render_frames_pipeline_equals_single_frames(rr, tmp_path):
    """render_frames (render.rs:926-989) deals frames to lanes / GPUs and keeps several in flight; every frame must be
    byte-identical to rendering that camera pose alone, in order, whatever the number of devices."""
    import ctypes as C
    import zlib

    import torch
    import yaml

    lib = C.CDLL(HOSTLIB)
    lib.rrh_env_new.restype = C.c_void_p
    lib.rrh_env_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64]
    lib.rrh_env_free.argtypes = [C.c_void_p]
    lib.rrh_env_deserialize.argtypes = [C.c_void_p, C.c_char_p]
    lib.rrh_render_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_int]
    lib.rrh_camera_motion.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
    w, h = 320, 180
    ren = rr.default_scene(w, h)
    doc = yaml.safe_load(ren.serialize())
    doc["camera_motion"] = [
        {"camera": {"position": {"x": 40.0, "y": -120.0, "z": -280.0}, "pyr": {"x": 0.1, "y": -1.4, "z": -1.5707964}},
         "velocity": {"x": 10.0, "y": 0.0, "z": 5.0}, "camera_target": None, "duration": 3.5},
        {"camera": {"position": {"x": 80.0, "y": -100.0, "z": -260.0}, "pyr": {"x": 0.0, "y": -1.2, "z": -1.5707964}},
         "velocity": {"x": 0.0, "y": 0.0, "z": 0.0}, "camera_target": {"x": 0.0, "y": -30.0, "z": 172.0}, "duration": 2.0},
    ]
    env = lib.rrh_env_new(0, w, h, 0, 0, 0.0, 0, 0)
    assert env and lib.rrh_env_deserialize(env, yaml.safe_dump(doc).encode()) == 0
    poses = (C.c_float * (7 * 32))()
    n = lib.rrh_camera_motion(env, poses, 32)
    assert n == 11  # int(3.5 / 0.5) + int(2.0 / 0.5)
    scene = rr.DeviceScene(ren, 0)
    want = []
    for i in range(n):
        p = ren.frame_params()
        p.cam_position[:] = poses[7 * i:7 * i + 3]
        p.cam_rotation[:] = poses[7 * i + 3:7 * i + 7]
        want.append(zlib.crc32(scene.render_rgb8(p).tobytes()))
    scene.close()
    assert len(set(want)) == n
    for ndev in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        crcs = (C.c_uint32 * 32)()
        sec = C.c_double()
        assert lib.rrh_render_frames(env, 0, ndev, C.byref(sec), crcs, 32) == n
        assert list(crcs[:n]) == want, f"{ndev} device(s)"
    lib.rrh_env_free(env)
