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
