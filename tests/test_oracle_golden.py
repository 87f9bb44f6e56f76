"""CPU tests that pin the oracle against every reference-authored vector for the hot path.

  * modutil.rs:16-55    fmod/imod/fimod/umod known-answer tests (exact asserts, fimod 1e-6)
  * pixelutil.rs:15-46  add_pixel/scale_pixel known-answer tests
  * images/example.png  the only rendered artefact in the reference (fixture: golden/example_png.npz)
"""
import ctypes as C
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fmod_kat(oracle):  # modutil.rs:16-23
    lib = oracle.load()
    assert lib.oracle_fmod(2.5, 2.5) == 0.0
    assert lib.oracle_fmod(2.5, 5.0) == 2.5
    assert lib.oracle_fmod(1.25, 2.0) == 1.25
    assert lib.oracle_fmod(5.0, 2.5) == 0.0
    assert lib.oracle_fmod(-2.75, 5.5) == 2.75


def test_imod_kat(oracle):  # modutil.rs:25-31
    lib = oracle.load()
    assert lib.oracle_imod(3, 5) == 3
    assert lib.oracle_imod(5, 3) == 2
    assert lib.oracle_imod(-2, 3) == 1
    assert lib.oracle_imod(-5, 7) == 2


def test_fimod_kat(oracle):  # modutil.rs:33-48
    lib = oracle.load()
    for (f, q), (ef, ei) in [((3.2, 5.0), (0.2, 3)), ((5.7, 3.0), (0.7, 2)), ((-2.5, 3.0), (0.5, 0)),
                             ((-5.9, 7.0), (0.1, 1))]:
        fr, i = C.c_float(), C.c_uint32()
        lib.oracle_fimod(f, q, C.byref(fr), C.byref(i))
        assert abs(fr.value - ef) < 1e-6
        assert i.value == ei


def test_umod_kat(oracle):  # modutil.rs:49-55
    lib = oracle.load()
    assert lib.oracle_umod(3, 5) == 3
    assert lib.oracle_umod(5, 3) == 2
    assert lib.oracle_umod(4, 3) == 1
    assert lib.oracle_umod(9, 7) == 2


def test_pixelutil_kat(oracle):  # pixelutil.rs:15-46
    lib = oracle.load()
    out = (C.c_float * 3)()
    lib.oracle_add_pixel(oracle.fa(1, 2, 3), oracle.fa(10, 20, 30), out)
    assert list(out) == [11.0, 22.0, 33.0]
    lib.oracle_add_pixel(oracle.fa(10, 20, 30), oracle.fa(1, 2, 3), out)
    assert list(out) == [11.0, 22.0, 33.0]
    lib.oracle_scale_pixel(3.5, (C.c_uint8 * 3)(1, 2, 3), out)
    assert list(out) == [3.5, 7.0, 10.5]
    sc = (C.c_float * 3)()
    lib.oracle_scale_pixel(2.0, (C.c_uint8 * 3)(1, 2, 3), sc)
    lib.oracle_add_pixel(oracle.fa(10, 20, 30), sc, out)
    assert list(out) == [12.0, 24.0, 36.0]


def test_powi_and_quantize(oracle):
    lib = oracle.load()
    a = np.float32(0.9371)
    # __powisf2 order for 24 = 0b11000: r = a^8 * a^16 built by squaring
    a2 = a * a; a4 = a2 * a2; a8 = a4 * a4; a16 = a8 * a8
    assert np.float32(lib.oracle_powi(float(a), 24)) == np.float32(a8 * a16)
    assert lib.oracle_powi(2.0, 0) == 1.0
    assert lib.oracle_powi(2.0, -2) == 0.25
    assert lib.oracle_quantize(2.0) == 255           # bgcolor's sun returns 2.0 (main.rs:245)
    assert lib.oracle_quantize(-0.5) == 0
    assert lib.oracle_quantize(float("nan")) == 255  # NaN.min(255.) == 255.
    assert lib.oracle_quantize(0.5) == 127           # truncation toward zero
    assert lib.oracle_quantize(1.0) == 255


def _example():
    return np.load(os.path.join(GOLD, "example_png.npz"))["rgb"]


def test_example_png_sky_current_scene(oracle, rr):
    """The sky region of example.png is reproduced by the CURRENT default scene (main.rs:154-276):
    pins primary-ray generation, Quat::transform, bgcolor and the u8 quantiser."""
    ex = _example()
    out = oracle.render(rr.default_scene(640, 480), want_tags=True)
    d = np.abs(ex.astype(int) - out["u8"].astype(int)).max(axis=2)
    top = d[:100]
    assert (top == 0).mean() >= 0.9995
    assert (top <= 1).all()
    # render.rs:840-843 counts for this config: 548 885 rays, 2.50 M object tests
    c = out["counts"]
    assert c.pixels == 640 * 480 and c.rays() == 548885 and c.object_tests == 2502740


def _legacy_scene(rr):
    """The scene example.png was rendered from: the current default scene without the second mirror
    sphere and with the floor mapped u=x, v=z (an older revision of main.rs)."""
    from ray_rust_b200 import scene as S

    S.UVMAPS.setdefault("XZ_LEGACY", 4)  # oracle-only mapping, see oracle/rr_oracle.h
    ren = rr.default_scene(640, 480)
    del ren._objects[2]
    ren._objects[0]._uvmap = "XZ_LEGACY"
    return ren


def test_example_png_pins_intersection_shading_reflection(oracle, rr):
    """Outside the glass sphere (whose refraction code changed after the image was made), the old
    scene rendered by this oracle reproduces example.png: sphere and floor intersection, Phong
    shading, hard shadows and mirror reflection are pinned by a reference-authored image."""
    from oracle.binding import load  # noqa: F401

    ex = _example()
    out = oracle.render(_legacy_scene(rr), want_tags=True)
    tags = out["tags"]
    d = np.abs(ex.astype(int) - out["u8"].astype(int)).max(axis=2)
    REFRACT, WRAP, SUN = 1 << 2, 1 << 6, 1 << 8
    m = (tags & (REFRACT | WRAP | SUN)) == 0
    assert m.sum() > 270000                       # > 88 % of the frame is covered by the check
    assert (d[m] == 0).mean() >= 0.996
    assert (d[m] <= 1).mean() >= 0.9995
    # every class of pixel the path produces is represented in the pinned set
    for bit in (1 << 0, 1 << 1, 1 << 3, 1 << 4):   # bg, reflect, shadowed, lit
        sel = m & ((tags & bit) != 0)
        assert sel.sum() > 1000
        assert (d[sel] <= 1).mean() >= 0.99
    # the mirror sphere interior (reflections of floor, sky and the red sphere)
    refl = m & ((tags & (1 << 1)) != 0)
    assert (d[refl] == 0).mean() >= 0.975


def test_threaded_equals_serial(oracle, rr):
    """render.rs:829-898: the N-thread row scheduler produces the same frame as the serial loop."""
    ren = rr.default_scene(160, 120)
    a = oracle.render(ren, threads=1, want_f32=True, want_counts=True)
    b = oracle.render(ren, threads=4, want_f32=True, want_counts=True)
    assert np.array_equal(a["f32"].view(np.uint32), b["f32"].view(np.uint32))
    assert np.array_equal(a["u8"], b["u8"])
    assert a["counts"].as_dict() == b["counts"].as_dict()


def test_march_miss_repeats_background(oracle, rr):
    """Appendix A Q15: in march mode a primary miss adds the sky three times."""
    t = oracle.render(rr.default_scene(64, 48), want_f32=True)["f32"]
    m = oracle.render(rr.default_scene(64, 48, use_raymarching=True), want_f32=True)["f32"]
    sky_t, sky_m = t[2, 5], m[2, 5]
    acc = np.float32(0) + sky_t
    acc = acc + sky_t
    acc = acc + sky_t
    assert np.array_equal(acc.view(np.uint32), sky_m.view(np.uint32))


def test_band_sharding_oracle(oracle, rr):
    ren = rr.default_scene(96, 70)
    full = oracle.render(ren)["u8"]
    rows = []
    for k in range(3):
        p = ren.frame_params(band_rows=8, band_index=k, band_count=3)
        rows.append(oracle.render(ren, params=p)["u8"])
    assert sum(r.shape[0] for r in rows) == 70
    out = np.zeros_like(full)
    for k in range(3):
        ys = [y for y in range(70) if (y // 8) % 3 == k]
        out[ys] = rows[k]
    assert np.array_equal(out, full)
    # unequal shares (band_span): rank 0 owns 2 of every 4 band slots, ranks 1 and 2 one each
    from ray_rust_b200 import bands
    spans, period = bands.weighted_spans(3, 2, 1)
    assert spans == [(0, 2), (2, 1), (3, 1)] and period == 4
    out2, tot = np.zeros_like(full), 0
    for idx, span in spans:
        p = ren.frame_params(band_rows=8, band_index=idx, band_count=period, band_span=span)
        got = oracle.render(ren, params=p)["u8"]
        ys = bands.span_rows(70, 8, idx, span, period)
        assert got.shape[0] == len(ys) == rr.frame_rows(p)
        out2[ys] = got
        tot += len(ys)
    assert tot == 70 and np.array_equal(out2, full)
