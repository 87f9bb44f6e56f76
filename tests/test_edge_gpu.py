"""Degenerate inputs: the device must follow the reference's IEEE behaviour (NaN/inf propagation, saturating
casts, f32::min/max NaN rules) exactly as the oracle does, not merely "not crash"."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1


def _scene(rr, objs, w=96, h=64, light=(50, 60, -50), march=False, glow=None, cam=(0, -150, -300)):
    f32 = np.float32
    return (rr.RenderEnv.new(cam, (f32(0), -rr.scene.PI / f32(2), -rr.scene.PI / f32(2)), w, h, 1.0, f32(h) / f32(w))
            .objects(objs).light(light).use_raymarching(march).glow_effect(glow))


def _cmp(rr, oracle, ren, exact_f32=True):
    ref = oracle.render(ren, threads=NCPU, want_f32=True, want_tags=True)
    scene = rr.DeviceScene(ren, 0)
    f = scene.render_f32(ren.frame_params())
    u8 = scene.render_rgb8(ren.frame_params())
    scene.close()
    d = np.abs(u8.astype(int) - ref["u8"].astype(int)).max()
    assert d <= 1, d
    if exact_f32:
        clean = (ref["tags"] & 1) == 0
        a, b = f.view(np.uint32)[clean], ref["f32"].view(np.uint32)[clean]
        nan_a, nan_b = np.isnan(f[clean]), np.isnan(ref["f32"][clean])
        assert np.array_equal(nan_a, nan_b)
        assert np.array_equal(a[~nan_a], b[~nan_b])


def _mats(rr):
    RC = rr.RenderColor
    return dict(
        floor=rr.RenderMaterial.new("floor", RC(1, 1, 0), RC(0, 0, 0), 0, 0.0, 0.0).pattern("RepeatedGradation").pattern_scale(300.0),
        red=rr.RenderMaterial.new("red", RC(0.8, 0.1, 0.1), RC(0.2, 0.2, 0.2), 24, 0.0, 0.0),
        glass=rr.RenderMaterial.new("glass", RC(0.1, 0.1, 0.1), RC(0.3, 0.3, 0.3), 8, 0.8, 1.5),
        mirror=rr.RenderMaterial.new("mirror", RC(0, 0, 0), RC(1, 1, 1), 24, 0.0, 0.0),
        negpn=rr.RenderMaterial.new("negpn", RC(0.5, 0.5, 0.9), RC(0.1, 0.1, 0.1), -3, 0.0, 0.0),      # powi with a negative exponent
        zeroscale=rr.RenderMaterial.new("zs", RC(0.5, 0.9, 0.5), RC(0, 0, 0), 4, 0.0, 0.0).pattern("Checkerboard").pattern_scale(0.0),
        nzero=rr.RenderMaterial.new("nzero", RC(0.1, 0.1, 0.1), RC(0, 0, 0), 0, 0.9, 0.0),              # 1/frac with frac = 0
    )


@pytest.mark.parametrize("march", [False, True])
def test_degenerate_objects(rr, oracle, march):
    m = _mats(rr)
    objs = [
        rr.RenderFloor.new_raw(m["floor"], (0, -300, 0), (0, 1, 0)).uvmap("ZX"),
        rr.RenderSphere.new(m["red"], -60.0, (-150, -100, 150)),        # negative radius
        rr.RenderSphere.new(m["mirror"], 0.0, (0, -100, 100)),          # zero radius
        rr.RenderSphere.new(m["glass"], 2000.0, (0, -150, -300)),       # camera inside a huge glass sphere
        rr.RenderSphere.new(m["negpn"], 70.0, (150, -120, 200)),
        rr.RenderSphere.new(m["zeroscale"], 50.0, (-40, -220, 60)),     # uv = x/0 -> inf/NaN -> floor() as i32 saturates
        rr.RenderSphere.new(m["nzero"], 45.0, (60, -240, 40)),
        rr.RenderFloor.new_raw(m["red"], (0, 0, 1500), (0.0, 0.0, 0.0)),  # zero normal: w == 0 -> 0/0
    ]
    _cmp(rr, oracle, _scene(rr, objs, march=march, glow=0.5 if march else None))


def test_zero_light_and_huge_coordinates(rr, oracle):
    m = _mats(rr)
    objs = [rr.RenderFloor.new_raw(m["floor"], (0, -300, 0), (0, 1, 0)).uvmap("ZX"),
            rr.RenderSphere.new(m["mirror"], 80.0, (0, -30, 172)), rr.RenderSphere.new(m["glass"], 100.0, (70, -200, 150))]
    _cmp(rr, oracle, _scene(rr, objs, light=(0, 0, 0)), exact_f32=True)          # normalized(0) = NaN light
    far = [rr.RenderFloor.new_raw(m["floor"], (0, -3e30, 0), (0, 1, 0)).uvmap("ZX"),
           rr.RenderSphere.new(m["red"], 1e19, (0, 0, 3e19)), rr.RenderSphere.new(m["glass"], 1e-30, (0, -150, -299))]
    _cmp(rr, oracle, _scene(rr, far))                                            # overflow to inf inside the tests


def test_nan_camera(rr, oracle):
    m = _mats(rr)
    objs = [rr.RenderFloor.new_raw(m["floor"], (0, -300, 0), (0, 1, 0)).uvmap("ZX"), rr.RenderSphere.new(m["red"], 80.0, (0, -30, 172))]
    ren = _scene(rr, objs, w=16, h=8)
    ren.camera.position = (np.float32("nan"), np.float32(0), np.float32(0))
    _cmp(rr, oracle, ren)


@pytest.mark.parametrize("w,h", [(1536, 1003), (2048, 1537), (1283, 997)])
def test_chunked_host_pipeline_odd_rows(rr, w, h):
    """rr_render_rgb8's kernel/copy pipeline (geometric chunk plan, chunks of whole 4-row tiles, a ragged last chunk)
    against one device-resident launch of the same frame, for frames big enough to be chunked, packed and padded."""
    import ctypes as C

    import torch

    ren = rr.default_scene(w, h)
    scene = rr.DeviceScene(ren, 0)
    p = ren.frame_params()
    dev = torch.empty(h * w * 3, dtype=torch.uint8, device="cuda:0")
    scene.render_rgb8_device(p, dev.data_ptr())
    torch.cuda.synchronize()
    whole = dev.cpu().numpy().reshape(h, w, 3)
    assert np.array_equal(scene.render_rgb8(p), whole)            # pageable destination
    stride = w * 3 + 8
    host = C.c_void_p()
    rr.ffi.check(scene.lib.rr_host_alloc(stride * h, C.byref(host)))
    pinned = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_uint8)), shape=(h, stride))
    pinned[:] = 0xEE
    rr.ffi.check(scene.lib.rr_render_rgb8(scene.handle, C.byref(p), host, stride))
    assert np.array_equal(pinned[:, : w * 3].reshape(h, w, 3), whole) and (pinned[:, w * 3:] == 0xEE).all()
    del pinned
    scene.lib.rr_host_free(host)
    scene.close()


def test_many_cameras_one_handle(rr, oracle):
    """One resident scene rendered from 40 different cameras and then from the first ones again: more cameras than the
    handle keeps primary-ray tables for (rr_ffi.cu, RR_PTABS = 32), so table slots are evicted and refilled, and a camera
    whose table is still cached must bind it without a refill. Every frame must be the oracle's."""
    ren = rr.default_scene(96, 64)
    scene = rr.DeviceScene(ren, 0)
    f32 = np.float32
    order = list(range(40)) + [0, 1, 39, 20, 0]
    for k in order:
        pyr = (f32(0.01 * k), -rr.scene.PI / f32(2) + f32(0.003 * k), -rr.scene.PI / f32(2) - f32(0.002 * k))
        ren.camera = rr.scene.Camera((f32(2.0 * k), f32(-150.0), f32(-300.0 + k)), pyr)
        got = scene.render_f32(ren.frame_params())
        ref = oracle.render(ren, threads=NCPU, want_f32=True, want_tags=True)
        clean = (ref["tags"] & 1) == 0
        assert np.array_equal(got.view(np.uint32)[clean], ref["f32"].view(np.uint32)[clean]), k
        if k % 7 == 0:  # resolution / fov changes re-key the tables too
            small = rr.default_scene(48, 32)
            small.camera = ren.camera
            g2 = scene.render_rgb8(small.frame_params())
            assert np.abs(g2.astype(int) - oracle.render(small, threads=NCPU)["u8"].astype(int)).max() <= 1
    scene.close()


def test_shared_reciprocal_divisions_equal_ieee_divisions(rr):
    """normalized() and get_uv() share one refined reciprocal per divisor (csrc/rr_device.cuh SharedRcp); the device
    self-test runs 2^26 hashed operand sets (zeros, subnormals, huge values, infinities and NaN included) through that
    path and through plain `/` and counts the sets that differ in any bit."""
    import ctypes as C

    lib = rr.ffi.load()
    for seed in (1, 20261018):
        bad = C.c_uint64(1)
        rr.ffi.check(lib.rr_selftest_normalize(0, 1 << 26, seed, C.byref(bad)))
        assert bad.value == 0


def test_march_row_profile_changes_the_order_not_the_frame(rr, oracle):
    """Ray-march launches record the longest tile of every tile row and later launches of the same geometry serve the rows
    longest first (MarchProfile, csrc/rr_ffi.cu). Scheduling only: repeated frames of one handle — first without a profile,
    then with the first frame's, then with an updated one — are byte-identical, also after a camera move and back, for a
    band shard, and equal to a fresh handle's frame."""
    import numpy as np

    ren = rr.default_scene(640, 360, use_raymarching=True, glow_effect=1.0)
    scene = rr.DeviceScene(ren, 0)
    p = ren.frame_params()
    frames = [scene.render_rgb8(p) for _ in range(4)]
    for f in frames[1:]:
        assert np.array_equal(f, frames[0])
    moved = rr.default_scene(640, 360, use_raymarching=True, glow_effect=1.0)
    moved.camera.position = (moved.camera.position[0] + 5.0, moved.camera.position[1] + 20.0, moved.camera.position[2])
    other = scene.render_rgb8(moved.frame_params())          # same geometry key, another camera: the old profile is only a hint
    assert not np.array_equal(other, frames[0])
    assert np.array_equal(scene.render_rgb8(p), frames[0])
    shard = ren.frame_params(16, 1, 3)
    a, b = scene.render_rgb8(shard), scene.render_rgb8(shard)
    assert np.array_equal(a, b)
    scene.close()
    fresh = rr.DeviceScene(ren, 0)
    assert np.array_equal(fresh.render_rgb8(p), frames[0])
    fresh2 = rr.DeviceScene(moved, 0)
    assert np.array_equal(fresh2.render_rgb8(moved.frame_params()), other)
    assert np.array_equal(fresh.render_rgb8(shard), a)
    fresh.close(); fresh2.close()
    ref = oracle.render(ren, threads=NCPU)
    assert np.abs(frames[0].astype(int) - ref["u8"].astype(int)).max() <= 1
