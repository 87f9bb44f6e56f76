"""Randomised parity: many small random scenes (object counts 0..40 incl. several floors, every material
kind, random cameras, depth limits, both modes) rendered by the device and by the oracle. f32 bit-exact on
every pixel that involved no libm transcendental, <= 1 LSB elsewhere."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1


def _random_env(rr, seed, march):
    rng = np.random.default_rng(seed)
    RC = rr.RenderColor
    f32 = np.float32
    mats = []
    for i in range(int(rng.integers(1, 7))):
        kind = rng.integers(0, 4)
        spec = float(rng.uniform(0, 1)) if kind in (0, 3) else 0.0
        m = rr.RenderMaterial.new(f"m{i}", RC(*rng.uniform(0, 1, 3)), RC(spec, spec, spec), int(rng.choice([0, 1, 5, 24])),
                                  float(rng.uniform(0.3, 1)) if kind == 2 else 0.0, float(rng.uniform(1.05, 2.0)))
        m.pattern(["Solid", "Checkerboard", "RepeatedGradation"][int(rng.integers(0, 3))]).pattern_scale(float(rng.uniform(5, 300)))
        if march and rng.random() < 0.3:
            m.glow_dist(float(rng.uniform(0.5, 6)))
        if rng.random() < 0.15:
            m.texture_data(rng.integers(0, 256, size=(int(rng.integers(2, 9)), int(rng.integers(2, 9)), 3), dtype=np.uint8),
                           "Bilinear" if rng.random() < 0.5 else "Nearest")
        mats.append(m)
    objs = []
    for _ in range(int(rng.integers(0, 4))):
        nrm = rng.normal(size=3)
        nrm[1] = abs(nrm[1]) + 0.5
        objs.append(rr.RenderFloor.new_raw(mats[int(rng.integers(len(mats)))], (float(rng.uniform(-50, 50)), float(rng.uniform(-400, -200)),
                                           float(rng.uniform(-50, 50))), tuple(float(x) for x in nrm)).uvmap(["XY", "YZ", "ZX"][int(rng.integers(0, 3))]))
    for _ in range(int(rng.integers(0, 38 if not march else 10))):
        objs.append(rr.RenderSphere.new(mats[int(rng.integers(len(mats)))], float(rng.uniform(5, 120)),
                                        (float(rng.uniform(-500, 500)), float(rng.uniform(-280, 200)), float(rng.uniform(-100, 900)))).uvmap(
            ["XY", "YZ", "ZX"][int(rng.integers(0, 3))]))
    rng.shuffle(objs)
    w, h = (int(rng.integers(3, 12)) * 8, int(rng.integers(3, 10)) * 4) if not march else (48, 32)
    pyr = (f32(rng.uniform(-0.3, 0.3)), f32(-np.pi / 2 + rng.uniform(-0.4, 0.4)), f32(-np.pi / 2 + rng.uniform(-0.2, 0.2)))
    ren = (rr.RenderEnv.new((float(rng.uniform(-100, 100)), float(rng.uniform(-200, 0)), float(rng.uniform(-400, -200))), pyr, w, h,
                            1.0, f32(h) / f32(w))
           .objects(objs).light(tuple(float(x) for x in rng.normal(size=3))).use_raymarching(march)
           .glow_effect(float(rng.uniform(0.2, 2)) if (march and rng.random() < 0.7) else None))
    ren.max_reflections = int(rng.integers(0, 6))
    ren.max_refractions = int(rng.integers(0, 12))
    return ren


@pytest.mark.parametrize("march", [False, True])
def test_random_scenes(rr, oracle, march):
    n = 40 if not march else 16
    for seed in range(n):
        ren = _random_env(rr, 1000 + seed + (500 if march else 0), march)
        ref = oracle.render(ren, threads=NCPU, want_f32=True, want_tags=True, want_counts=True)
        scene = rr.DeviceScene(ren, 0)
        p = ren.frame_params()
        f = scene.render_f32(p)
        u8, cnt = scene.render_count(p)
        scene.close()
        d = np.abs(u8.astype(int) - ref["u8"].astype(int)).max(initial=0)
        assert d <= 1, (seed, d)
        glow = ren._glow_effect is not None
        clean = (ref["tags"] & 1) == 0 if not glow else np.zeros_like(ref["tags"], dtype=bool)  # powf in the glow factor
        a, b = f.view(np.uint32)[clean], ref["f32"].view(np.uint32)[clean]
        na, nb = np.isnan(f[clean]), np.isnan(ref["f32"][clean])
        assert np.array_equal(na, nb), seed
        assert np.array_equal(a[~na], b[~nb]), seed
        assert cnt.as_dict() == ref["counts"].as_dict(), seed
