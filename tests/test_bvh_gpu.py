"""The exact-culling (BVH) instance of the trace kernel must be bit-identical to the brute-force scan
of render.rs:993-1018, on the device (f32 frames compared as raw bits) and against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1


def _both(rr, ren):
    scene = rr.DeviceScene(ren, 0)
    p = ren.frame_params()
    scene.set_culling(True)
    a = scene.render_f32(p)
    a8 = scene.render_rgb8(p)
    scene.set_culling(False)
    b = scene.render_f32(p)
    b8 = scene.render_rgb8(p)
    scene.close()
    return a, b, a8, b8


def _random_scene(rr, seed, n, w=192, h=108, spread=1.0, tiny=False, dup=False, inside=False, badfloor=False):
    rng = np.random.default_rng(seed)
    RC = rr.RenderColor
    floor = rr.RenderMaterial.new("floor", RC(1, 1, 0), RC(0, 0, 0), 0, 0.0, 0.0).pattern("RepeatedGradation").pattern_scale(300.0)
    mats = []
    for i in range(12):
        kind = i % 3
        if kind == 0:
            mats.append(rr.RenderMaterial.new(f"mir{i}", RC(*rng.uniform(0, 0.3, 3)), RC(*([rng.uniform(0.5, 1)] * 3)), 24, 0.0, 0.0))
        elif kind == 1:
            mats.append(rr.RenderMaterial.new(f"dif{i}", RC(*rng.uniform(0.1, 1, 3)), RC(0, 0, 0), 16, 0.0, 0.0).pattern(
                "Checkerboard").pattern_scale(7.0))
        else:
            mats.append(rr.RenderMaterial.new(f"gla{i}", RC(0, 0, 0), RC(*([rng.uniform(0, 0.4)] * 3)), 0, rng.uniform(0.4, 1),
                                              rng.uniform(1.1, 1.9)))
    objs = [rr.RenderFloor.new_raw(floor, (0, -300, 0), (0, 1, 0)).uvmap("ZX")]
    for k in range(n):
        r = rng.uniform(0.01, 0.5) if (tiny and k % 3 == 0) else rng.uniform(10, 60)
        c = (rng.uniform(-900, 900) * spread, rng.uniform(-280, 300) * spread, rng.uniform(-100, 1900) * spread)
        objs.append(rr.RenderSphere.new(mats[rng.integers(len(mats))], r, c))
        if dup and k % 5 == 0:   # an exact duplicate with another material: the lower index must win every tie
            objs.append(rr.RenderSphere.new(mats[rng.integers(len(mats))], r, c))
    if badfloor:
        # an un-normalised floor normal that is not object 0: bounces off it change |eye| (appendix A Q22), the
        # reference's sphere test then is no longer the geometric one and such rays must bypass the BVH
        objs.insert(len(objs) // 2, rr.RenderFloor.new_raw(mats[0], (20, -310, -48), (-0.47, 0.82, 1.31)).uvmap("XY"))
    if inside:
        objs.append(rr.RenderSphere.new(mats[2], 500.0, (0, -150, -300)))  # the camera sits inside a glass sphere
    f32 = np.float32
    ren = (rr.RenderEnv.new((0, -150, -300), (f32(0), -rr.scene.PI / f32(2), -rr.scene.PI / f32(2)), w, h, 1.0, f32(h) / f32(w))
           .objects(objs).light((50, 60, -50)))
    if badfloor:
        ren.max_reflections = 5
    return ren


@pytest.mark.parametrize("kw", [
    dict(seed=1, n=30), dict(seed=2, n=200), dict(seed=3, n=500, tiny=True), dict(seed=4, n=300, dup=True),
    dict(seed=5, n=100, spread=40.0), dict(seed=6, n=150, inside=True), dict(seed=7, n=1500, w=128, h=72),
    dict(seed=8, n=64, spread=0.05), dict(seed=9, n=120, badfloor=True), dict(seed=10, n=400, badfloor=True, dup=True),
])
def test_bvh_equals_bruteforce_bits(rr, kw):
    a, b, a8, b8 = _both(rr, _random_scene(rr, **kw))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(a8, b8)


def test_bvh_vs_oracle_synthetic(rr, oracle):
    ren = rr.synthetic_scene(320, 180)
    ref = oracle.render(ren, threads=NCPU, want_f32=True, want_tags=True, want_counts=True)
    scene = rr.DeviceScene(ren, 0)
    f = scene.render_f32(ren.frame_params())
    _, cnt = scene.render_count(ren.frame_params(), want_image=False)
    scene.close()
    clean = (ref["tags"] & 1) == 0
    assert np.array_equal(f.view(np.uint32)[clean], ref["f32"].view(np.uint32)[clean])
    assert cnt.as_dict() == ref["counts"].as_dict()   # reference-equivalent counts, whatever was culled


def test_bvh_far_origins(rr):
    """Rays that start very far away (horizon floor points, a camera at 1e6) get boxes inflated to match
    the reference test's round-off at that distance; the result must still be the brute-force one."""
    ren = _random_scene(rr, seed=11, n=120)
    ren.camera.position = (np.float32(0), np.float32(2.0e5), np.float32(-1.0e6))
    a, b, _, _ = _both(rr, ren)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("kw", [dict(seed=21, n=30), dict(seed=22, n=300, dup=True), dict(seed=23, n=150, inside=True),
                                dict(seed=24, n=700, tiny=True), dict(seed=25, n=100, spread=40.0)])
@pytest.mark.parametrize("glow", [None, 0.8])
def test_march_bvh_equals_linear_scan_bits(rr, kw, glow):
    """Ray-march mode of large scenes: the distance scan pruned through the BVH (MBVH instance) against the linear scan
    (culling off), f32 bit for bit and byte for byte."""
    ren = _random_scene(rr, w=96, h=54, **kw)
    ren.use_raymarching(True).glow_effect(glow)
    a, b, a8, b8 = _both(rr, ren)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(a8, b8)


def test_march_bvh_vs_oracle(rr, oracle):
    ren = rr.synthetic_scene(160, 90, n_spheres=400, use_raymarching=True, glow_effect=1.0)
    ref = oracle.render(ren, threads=NCPU, want_counts=True)
    scene = rr.DeviceScene(ren, 0)
    img, cnt = scene.render_count(ren.frame_params())
    scene.close()
    d = np.abs(img.astype(int) - ref["u8"].astype(int)).max(axis=2)
    assert (d <= 1).mean() >= 0.9995 and (d == 0).mean() >= 0.995, ((d == 0).mean(), d.max())
    assert cnt.as_dict() == ref["counts"].as_dict()
