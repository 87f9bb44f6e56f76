"""CPU check of the KERNEL SOURCE LOGIC: ray-rust_b200/csrc/rr_trace.cuh and rr_march.cuh are compiled for the
host (tests/hostsim, CUDA built-ins replaced by stand-ins, -ffp-contract=off) and every pixel is compared with the
oracle. This runs in the no-GPU container, so a logic regression in trace_pixel/march_pixel is caught before a
GPU is involved, including the exact BVH cull (host builder csrc/rr_bvh.h + bvh_scan_ordered). (GPU code generation and
the store path are covered by the -m gpu tests.)"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS = os.path.join(ROOT, "tests", "hostsim")
sys.path.insert(0, os.path.join(ROOT, "tests"))


@pytest.fixture(scope="module")
def hostsim(rr):
    so = os.path.join(HS, "libhostsim.so")
    deps = [os.path.join(HS, f) for f in ("hostsim.cpp", "cuda_stub.h")] + [
        os.path.join(ROOT, "ray-rust_b200", "csrc", f) for f in ("rr_device.cuh", "rr_trace.cuh", "rr_march.cuh", "rr_bvh.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-o", so, os.path.join(HS, "hostsim.cpp")])
    lib = C.CDLL(so)
    lib.hostsim_render_f32.argtypes = [C.POINTER(rr.ffi.rr_scene_desc), C.POINTER(rr.ffi.rr_frame_params), C.c_void_p,
                                       C.POINTER(rr.ffi.rr_ray_counts)]

    lib.hostsim_render_f32_ex.argtypes = lib.hostsim_render_f32.argtypes + [C.c_int, C.POINTER(C.c_int)]

    def render(ren, culling=False, general=False):
        flat = ren.flatten()
        p = ren.frame_params()
        out = np.empty((p.yres, p.xres, 3), dtype=np.float32)
        cnt = rr.ffi.rr_ray_counts()
        used = C.c_int(0)
        assert lib.hostsim_render_f32_ex(C.byref(flat.desc), C.byref(p), out.ctypes.data_as(C.c_void_p), C.byref(cnt),
                                         (1 if culling else 0) | (2 if general else 0), C.byref(used)) == 0
        if culling:
            assert used.value == 1, "the builder produced no tree for this scene"
        return out, cnt

    return render


def _check(ren, oracle, hostsim, culling=False, general=False):
    ref = oracle.render(ren, threads=os.cpu_count() or 1, want_f32=True, want_tags=True, want_counts=True)
    out, cnt = hostsim(ren, culling, general)
    # same libm on both sides here, so even bgcolor / glow pixels must agree bit for bit
    a, b = out.view(np.uint32), ref["f32"].view(np.uint32)
    nan = np.isnan(out)
    assert np.array_equal(nan, np.isnan(ref["f32"]))
    assert np.array_equal(a[~nan], b[~nan])
    assert cnt.as_dict() == ref["counts"].as_dict()


@pytest.mark.parametrize("march", [False, True])
def test_default_scene_logic(rr, oracle, hostsim, march):
    _check(rr.default_scene(96, 72, use_raymarching=march, glow_effect=1.0 if march else None), oracle, hostsim)


def test_default_scene_general_instance(rr, oracle, hostsim):
    """The built-in scene fits the SceneHead, so the device renders it with the head-only instance (no count checks, no
    tail loops, padded never-hit slots); the general instance must give the same bits."""
    _check(rr.default_scene(96, 72), oracle, hostsim, general=True)


@pytest.mark.parametrize("n_spheres", [0, 1, 2, 3, 4, 5, 7])
def test_head_sizes_logic(rr, oracle, hostsim, n_spheres):
    """Sphere counts around the head size: odd pairs, padded slots, a tail of one; head-only and general instance."""
    ren = rr.synthetic_scene(48, 32, n_spheres=n_spheres, seed=11 + n_spheres)
    _check(ren, oracle, hostsim)
    _check(ren, oracle, hostsim, general=True)


def test_synthetic_scene_logic(rr, oracle, hostsim):
    _check(rr.synthetic_scene(64, 36, n_spheres=60), oracle, hostsim)


@pytest.mark.parametrize("march", [False, True])
def test_random_scenes_logic(rr, oracle, hostsim, march):
    from test_random_scenes_gpu import _random_env

    for seed in range(12 if not march else 5):
        _check(_random_env(rr, 1000 + seed + (500 if march else 0), march), oracle, hostsim)


def test_bvh_logic_synthetic(rr, oracle, hostsim):
    """The BVH instance (SAH build, ordered stack traversal, FMA slabs) against the brute-force oracle, bit for bit."""
    _check(rr.synthetic_scene(96, 54, n_spheres=300), oracle, hostsim, culling=True)
    _check(rr.synthetic_scene(64, 36, n_spheres=1024), oracle, hostsim, culling=True)
    _check(rr.synthetic_scene(48, 27, n_spheres=24, seed=7), oracle, hostsim, culling=True)   # smallest tree


def test_bvh_logic_hard_cases(rr, oracle, hostsim):
    """Duplicates (exact ties: lowest index must win whatever the visiting order), tiny and huge spheres, a camera
    inside a sphere, and a scene far from the coordinate origin (the FMA slab form rounds (o +- e)/d on its own)."""
    from ray_rust_b200.scene import RenderColor, RenderMaterial, RenderSphere, RenderFloor

    def env(spheres, cam=(0.0, -150.0, -300.0), floor_y=300.0):
        mats = {
            "m": RenderMaterial.new("m", RenderColor(0.8, 0.2, 0.2), RenderColor(0.3, 0.3, 0.3), 24, 0.0, 0.0),
            "g": RenderMaterial.new("g", RenderColor(0.0, 0.0, 0.0), RenderColor(0.2, 0.2, 0.2), 0, 0.8, 1.5),
            "f": RenderMaterial.new("f", RenderColor(1.0, 1.0, 0.0), RenderColor(0.0, 0.0, 0.0), 0, 0.0, 0.0).pattern("Checkerboard").pattern_scale(100.0),
        }
        objs = [RenderFloor.new(mats["f"], (cam[0], floor_y, cam[2]), (0.0, -1.0, 0.0))]
        for i, (c, r) in enumerate(spheres):
            objs.append(RenderSphere.new(mats["g" if i % 3 == 0 else "m"], r, c))
        base = rr.default_scene(72, 40)
        ren = rr.RenderEnv.new(cam, tuple(base.camera.pyr), 72, 40, 1.0, 40.0 / 72.0)
        return ren.materials(mats).objects(objs).light((50.0, 60.0, -50.0))

    rng = np.random.default_rng(5)
    pts = [((float(x), float(y), float(z)), float(r)) for x, y, z, r in
           zip(rng.uniform(-400, 400, 40), rng.uniform(-200, 200, 40), rng.uniform(0, 900, 40), rng.uniform(5, 60, 40))]
    _check(env(pts + pts[:10]), oracle, hostsim, culling=True)                                       # exact duplicates
    _check(env(pts + [((0.0, 0.0, 400.0), 1e-3), ((50.0, 0.0, 500.0), 700.0)]), oracle, hostsim, culling=True)
    _check(env(pts + [((0.0, -150.0, -300.0), 90.0)]), oracle, hostsim, culling=True)                # camera inside a glass sphere
    off = (1.0e6, -2.0e6, 3.0e6)
    far = [((c[0] + off[0], c[1] + off[1], c[2] + off[2]), r) for c, r in pts]
    _check(env(far, cam=(off[0], off[1] - 150.0, off[2] - 300.0), floor_y=off[1] + 300.0), oracle, hostsim, culling=True)


def test_march_bvh_logic(rr, oracle, hostsim):
    """Ray-march mode of large scenes: the distance scan pruned through the BVH (point-to-box distance, rr_march.cuh MBVH)
    against the brute-force oracle, bit for bit, with and without the glow pass, incl. a scene far from the origin."""
    _check(rr.synthetic_scene(64, 36, n_spheres=300, use_raymarching=True), oracle, hostsim, culling=True)
    _check(rr.synthetic_scene(48, 27, n_spheres=1024, use_raymarching=True, glow_effect=0.7), oracle, hostsim, culling=True)
    _check(rr.synthetic_scene(40, 24, n_spheres=24, seed=11, use_raymarching=True, glow_effect=1.0), oracle, hostsim, culling=True)


def test_fmod_2pi_is_fmodf(hostsim):
    """bgcolor's fmodf(x, 2*pi) is replaced on the device by a three-instruction exact remainder (rr_device.cuh fmod_2pi):
    identical bits for every float in [1, 1024) (bgcolor's arguments lie in [58, 256]) and the fmodf fallback outside."""
    so = os.path.join(HS, "libhostsim.so")
    lib = C.CDLL(so)
    lib.hostsim_fmod_2pi_mismatches.restype = C.c_longlong
    lib.hostsim_fmod_2pi_mismatches.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float)]
    bad = C.c_float(0)
    for lo, hi in ((1.0, 1024.0), (1e-30, 1e-29), (4000.0, 4200.0), (0.001, 0.002)):
        assert lib.hostsim_fmod_2pi_mismatches(lo, hi, C.byref(bad)) == 0, bad.value


def test_march_tile_rotation_hint(rr, hostsim):
    """The ray-march kernel serves its tile rows from the floor's horizon on (FrameParams::march_tile_rot, csrc/rr_device.cuh):
    a host-side scheduling hint. The built-in camera looks along the floor, so the horizon is the middle row; trace-mode
    frames keep image order (band shards too: the C ABI's to_dev, exercised on the GPU box)."""
    lib = C.CDLL(os.path.join(HS, "libhostsim.so"))  # (built by the hostsim fixture)

    def rot(ren, p=None):
        flat = ren.flatten()
        p = p or ren.frame_params()
        return lib.hostsim_march_tile_rot(C.byref(flat.desc), C.byref(p))

    lib.hostsim_march_tile_rot.argtypes = [C.POINTER(rr.ffi.rr_scene_desc), C.POINTER(rr.ffi.rr_frame_params)]
    march = rr.default_scene(3840, 2160, use_raymarching=True, glow_effect=1.0)
    assert rot(march) == 2160 // 2 // 4 - 2
    assert rot(rr.default_scene(640, 480, use_raymarching=True)) == 480 // 2 // 4 - 2
    assert rot(rr.default_scene(3840, 2160)) == 0                       # ray-trace mode
    up = rr.default_scene(640, 480, use_raymarching=True)
    up.camera.rotation = rr.scene.Quat.from_pyr((np.float32(0.0), np.float32(0.0), np.float32(0.0)))  # looking straight up/down the axis: no horizon row
    r = rot(up)
    assert 0 <= r < 480 // 4
