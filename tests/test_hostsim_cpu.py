"""CPU check of the KERNEL SOURCE LOGIC: ray-rust_b200/csrc/rr_trace.cuh and rr_march.cuh are compiled for the
host (tests/hostsim, CUDA built-ins replaced by stand-ins, -ffp-contract=off) and every pixel is compared with the
oracle. This runs in the no-GPU container, so a logic regression in trace_pixel/march_pixel is caught before a
GPU is involved. (GPU code generation, the BVH instance and the store path are covered by the -m gpu tests.)"""
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
        os.path.join(ROOT, "ray-rust_b200", "csrc", f) for f in ("rr_device.cuh", "rr_trace.cuh", "rr_march.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-o", so, os.path.join(HS, "hostsim.cpp")])
    lib = C.CDLL(so)
    lib.hostsim_render_f32.argtypes = [C.POINTER(rr.ffi.rr_scene_desc), C.POINTER(rr.ffi.rr_frame_params), C.c_void_p,
                                       C.POINTER(rr.ffi.rr_ray_counts)]

    def render(ren):
        flat = ren.flatten()
        p = ren.frame_params()
        out = np.empty((p.yres, p.xres, 3), dtype=np.float32)
        cnt = rr.ffi.rr_ray_counts()
        assert lib.hostsim_render_f32(C.byref(flat.desc), C.byref(p), out.ctypes.data_as(C.c_void_p), C.byref(cnt)) == 0
        return out, cnt

    return render


def _check(ren, oracle, hostsim):
    ref = oracle.render(ren, want_f32=True, want_tags=True, want_counts=True)
    out, cnt = hostsim(ren)
    # same libm on both sides here, so even bgcolor / glow pixels must agree bit for bit
    a, b = out.view(np.uint32), ref["f32"].view(np.uint32)
    nan = np.isnan(out)
    assert np.array_equal(nan, np.isnan(ref["f32"]))
    assert np.array_equal(a[~nan], b[~nan])
    assert cnt.as_dict() == ref["counts"].as_dict()


@pytest.mark.parametrize("march", [False, True])
def test_default_scene_logic(rr, oracle, hostsim, march):
    _check(rr.default_scene(96, 72, use_raymarching=march, glow_effect=1.0 if march else None), oracle, hostsim)


def test_synthetic_scene_logic(rr, oracle, hostsim):
    _check(rr.synthetic_scene(64, 36, n_spheres=60), oracle, hostsim)


@pytest.mark.parametrize("march", [False, True])
def test_random_scenes_logic(rr, oracle, hostsim, march):
    from test_random_scenes_gpu import _random_env

    for seed in range(12 if not march else 5):
        _check(_random_env(rr, 1000 + seed + (500 if march else 0), march), oracle, hostsim)
