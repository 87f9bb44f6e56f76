"""examples/render_c.c: the C ABI consumed from plain C (no Python, no C++ host layer), as a cgo / JNI / Rust -sys binding
would. CPU: it compiles against include/rr_ffi.h with gcc, links the shared library and fails loudly without a device.
GPU: its frame equals the oracle's render of the same scene."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ray-rust_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory, rr):
    rr.ffi.load()  # the library must exist (built by __graft_entry__.build())
    out = str(tmp_path_factory.mktemp("c_example") / "render_c")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "render_c.c"), "-L", LIBDIR, "-lray_rust_b200",
                           f"-Wl,-rpath,{LIBDIR}", "-lm", "-o", out])
    return out


def test_c_example_builds_and_has_no_cpu_fallback(exe, tmp_path):
    import torch

    assert subprocess.run([exe], capture_output=True).returncode == 1          # usage
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "64", "48", str(tmp_path / "o.ppm")], capture_output=True, text=True)
    assert r.returncode == 2 and "rr_scene_create failed" in r.stderr and not (tmp_path / "o.ppm").exists()


@pytest.mark.gpu
def test_c_example_frame_matches_oracle(exe, tmp_path, rr, oracle):
    from ray_rust_b200.scene import Quat, RenderColor, RenderFloor, RenderMaterial, RenderSphere

    w, h = 320, 200
    out = tmp_path / "o.ppm"
    r = subprocess.run([exe, str(w), str(h), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = out.read_bytes()
    head = f"P6\n{w} {h}\n255\n".encode()
    assert raw.startswith(head) and len(raw) == len(head) + w * h * 3
    img = np.frombuffer(raw[len(head):], dtype=np.uint8).reshape(h, w, 3)
    floor = (RenderMaterial.new("floor", RenderColor(1.0, 1.0, 0.0), RenderColor(0.0, 0.0, 0.0), 0, 0.0, 0.0)
             .pattern("RepeatedGradation").pattern_scale(300.0).pattern_angle_scale(0.2))
    mirror = RenderMaterial.new("mirror", RenderColor(0.0, 0.0, 0.0), RenderColor(1.0, 1.0, 1.0), 24, 0.0, 0.0)
    glass = RenderMaterial.new("glass", RenderColor(0.0, 0.0, 0.0), RenderColor(1.0, 1.0, 1.0), 24, 1.0, 1.5)
    base = rr.default_scene(w, h)
    ren = rr.RenderEnv.new((0.0, -150.0, -300.0), tuple(base.camera.pyr), w, h, 1.0, np.float32(h) / np.float32(w))
    ren.objects([RenderFloor.new_raw(floor, (0.0, -300.0, 0.0), (0.0, 1.0, 0.0)).uvmap("ZX"),
                 RenderSphere.new(mirror, 80.0, (-120.0, -220.0, 172.0)), RenderSphere.new(glass, 100.0, (90.0, -200.0, 150.0))])
    ren.light((50.0, 60.0, -50.0))
    ren.camera.rotation = Quat(-0.5, -0.5, -0.5, 0.5)   # the example hard-codes the exact quaternion
    assert bytes(ren.frame_params())[:64] != b"" and tuple(ren.frame_params().cam_rotation) == (-0.5, -0.5, -0.5, 0.5)
    ref = oracle.render(ren, threads=os.cpu_count() or 1)["u8"]
    d = np.abs(img.astype(int) - ref.astype(int)).max(axis=2)
    assert (d <= 1).mean() >= 0.999 and (d == 0).mean() >= 0.99, ((d == 0).mean(), (d <= 1).mean(), d.max())
