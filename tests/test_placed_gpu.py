"""Placed output (multi-GPU path): shards render their row bands straight to the image rows of one frame,
in device memory (the pointer may be NVLink peer memory in the real multi-process run) or in host memory.
Exercised here on one GPU by letting a single process play every rank in turn."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,band,w,h,march", [(2, 16, 256, 150, False), (8, 16, 640, 360, False), (3, 4, 96, 70, True), (4, 5, 64, 37, False)])
def test_placed_device_and_host(rr, n, band, w, h, march):
    import torch

    ren = rr.default_scene(w, h, use_raymarching=march, glow_effect=1.0 if march else None)
    scene = rr.DeviceScene(ren, 0)
    lib = scene.lib
    full = scene.render_rgb8(ren.frame_params())
    frame = torch.zeros(h * w * 3, dtype=torch.uint8, device="cuda:0")
    host = np.zeros((h, w, 3), dtype=np.uint8)
    for k in range(n):
        p = ren.frame_params(band, k, n)
        rr.ffi.check(lib.rr_render_rgb8_placed_device(scene.handle, C.byref(p), C.c_void_p(frame.data_ptr()), 0, None))
        rr.ffi.check(lib.rr_render_rgb8_placed(scene.handle, C.byref(p), host.ctypes.data_as(C.c_void_p), 0))
    torch.cuda.synchronize()
    assert np.array_equal(frame.cpu().numpy().reshape(h, w, 3), full)
    assert np.array_equal(host, full)
    # padded host rows
    stride = w * 3 + 12
    padded = np.full((h, stride), 0xCD, dtype=np.uint8)
    for k in range(n):
        p = ren.frame_params(band, k, n)
        rr.ffi.check(lib.rr_render_rgb8_placed(scene.handle, C.byref(p), padded.ctypes.data_as(C.c_void_p), stride))
    assert np.array_equal(padded[:, : w * 3].reshape(h, w, 3), full) and (padded[:, w * 3:] == 0xCD).all()
    scene.close()


def test_placed_big_frame_chunked(rr):
    """A frame large enough for the chunked copy pipeline (several chunks of whole bands)."""
    ren = rr.default_scene(3840, 2160)
    scene = rr.DeviceScene(ren, 0)
    full = scene.render_rgb8(ren.frame_params())
    host = np.zeros_like(full)
    for k in range(2):
        p = ren.frame_params(16, k, 2)
        rr.ffi.check(scene.lib.rr_render_rgb8_placed(scene.handle, C.byref(p), host.ctypes.data_as(C.c_void_p), 0))
    whole = np.zeros_like(full)
    rr.ffi.check(scene.lib.rr_render_rgb8_placed(scene.handle, C.byref(ren.frame_params()), whole.ctypes.data_as(C.c_void_p), 0))
    scene.close()
    assert np.array_equal(host, full) and np.array_equal(whole, full)


def test_device_alloc_and_ipc_export(rr):
    lib = rr.ffi.load()
    p = C.c_void_p()
    rr.ffi.check(lib.rr_device_alloc(0, 1 << 20, C.byref(p)))
    handle = (C.c_uint8 * 64)()
    rr.ffi.check(lib.rr_ipc_export(p, handle))
    assert any(handle)
    rr.ffi.check(lib.rr_device_free(0, p))
    assert lib.rr_ipc_export(None, handle) == rr.ffi.RR_ERR_BAD_ARG
    buf = np.zeros(1 << 16, dtype=np.uint8)
    rr.ffi.check(lib.rr_host_register(buf.ctypes.data_as(C.c_void_p), buf.nbytes))
    rr.ffi.check(lib.rr_host_unregister(buf.ctypes.data_as(C.c_void_p)))


@pytest.mark.parametrize("n,w,h,march", [(4, 640, 360, False), (3, 96, 70, True), (8, 64, 24, False)])
def test_placed_signal_and_fence_wait(rr, n, w, h, march):
    """Fused completion (rr_render_rgb8_placed_signal_device + rr_fence_wait_device): every shard's kernel publishes its
    epoch word; the owner's stream waits for all of them. One process plays every rank, each on its own stream, with the
    wait queued FIRST so that it really has to wait for the flags."""
    import torch

    ren = rr.default_scene(w, h, use_raymarching=march, glow_effect=1.0 if march else None)
    scene = rr.DeviceScene(ren, 0)
    lib = scene.lib
    full = scene.render_rgb8(ren.frame_params())
    frame = torch.zeros(h * w * 3, dtype=torch.uint8, device="cuda:0")
    flags = torch.zeros(n + 1, dtype=torch.int32, device="cuda:0")   # n completion words + 1 status word
    fptr, sptr = C.c_void_p(flags.data_ptr()), C.c_void_p(flags.data_ptr() + 4 * n)
    owner, worker = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for epoch in (1, 2):
        frame.zero_()
        torch.cuda.synchronize()
        rr.ffi.check(lib.rr_fence_wait_device(0, fptr, n, epoch, 20000, sptr, C.c_void_p(owner.cuda_stream)))
        with torch.cuda.stream(owner):
            snapshot = frame.clone()          # ordered behind the wait on the owner's stream only
        for k in range(n):                    # 8 shards of a 24-row frame: the last two own no rows at all
            p = ren.frame_params(4, k, n)
            rr.ffi.check(lib.rr_render_rgb8_placed_signal_device(scene.handle, C.byref(p), C.c_void_p(frame.data_ptr()), 0, fptr, epoch,
                                                                 C.c_void_p(worker.cuda_stream)))
        torch.cuda.synchronize()
        assert flags.cpu().tolist() == [epoch] * n + [0]
        assert np.array_equal(snapshot.cpu().numpy().reshape(h, w, 3), full)
    # a missing shard: the wait gives up after its timeout and reports it in the status word
    rr.ffi.check(lib.rr_fence_wait_device(0, fptr, n, 3, 50, sptr, C.c_void_p(owner.cuda_stream)))
    torch.cuda.synchronize()
    assert flags.cpu().tolist()[n] == 1
    # argument checks
    p = ren.frame_params(4, 0, n)
    assert lib.rr_render_rgb8_placed_signal_device(scene.handle, C.byref(p), C.c_void_p(frame.data_ptr()), 0, None, 1, C.c_void_p(worker.cuda_stream)) == rr.ffi.RR_ERR_BAD_ARG
    scene.close()


@pytest.mark.parametrize("w,h,march,synth", [(640, 360, False, False), (256, 150, True, False), (256, 144, False, True)])
def test_placed_unequal_band_spans(rr, w, h, march, synth):
    """rr_frame_params.band_span: the frame's owner renders a larger share (3 of every 8 band slots, five other shards one
    each). Placed, signalled launches of all shards must assemble the 1-GPU frame; packed launches must hold exactly the
    shard's rows; the host-frame and un-interleave paths refuse unequal spans."""
    import torch
    from ray_rust_b200 import bands

    ren = rr.synthetic_scene(w, h, n_spheres=64) if synth else rr.default_scene(w, h, use_raymarching=march, glow_effect=1.0 if march else None)
    scene = rr.DeviceScene(ren, 0)
    lib = scene.lib
    full = scene.render_rgb8(ren.frame_params())
    spans, period = bands.weighted_spans(6, 3, 1)
    frame = torch.zeros(h * w * 3, dtype=torch.uint8, device="cuda:0")
    flags = torch.zeros(8, dtype=torch.int32, device="cuda:0")
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    for r, (idx, span) in enumerate(spans):
        p = ren.frame_params(4, idx, period, span)
        word = C.c_void_p(flags.data_ptr() + 4 * (r - idx))   # rr_ffi.h: d_flags[band_index] is written
        rr.ffi.check(lib.rr_render_rgb8_placed_signal_device(scene.handle, C.byref(p), C.c_void_p(frame.data_ptr()), 0, word, 7, C.c_void_p(st.cuda_stream)))
        packed = scene.render_rgb8(p)
        assert np.array_equal(packed, full[bands.span_rows(h, 4, idx, span, period)])
    torch.cuda.synchronize()
    assert flags.cpu().tolist() == [7] * 6 + [0, 0]
    assert np.array_equal(frame.cpu().numpy().reshape(h, w, 3), full)
    p = ren.frame_params(4, 0, period, 3)
    host = np.zeros((h, w, 3), dtype=np.uint8)
    assert lib.rr_render_rgb8_placed(scene.handle, C.byref(p), host.ctypes.data_as(C.c_void_p), 0) == rr.ffi.RR_ERR_UNSUPPORTED
    assert lib.rr_bands_unpack_device(C.byref(p), C.c_void_p(frame.data_ptr()), 0, C.c_void_p(frame.data_ptr()), None) == rr.ffi.RR_ERR_UNSUPPORTED
    scene.close()


@pytest.mark.parametrize("kind", ["march", "synthetic"])
def test_zero_copy_pinned_host_frame(rr, kind):
    """Long kernels store straight into a page-locked host frame (no device buffer, no DMA copy); the bytes must equal
    the pageable-buffer path, for packed and padded rows, whole frames and row-band shards placed into one frame."""
    w, h = 256, 144
    ren = (rr.default_scene(w, h, use_raymarching=True, glow_effect=1.0) if kind == "march"
           else rr.synthetic_scene(w, h, n_spheres=100))
    scene = rr.DeviceScene(ren, 0)
    lib = scene.lib
    p = ren.frame_params()
    pageable = scene.render_rgb8(p)                       # numpy buffer: device buffer + cudaMemcpy
    stride = w * 3 + 16
    host = C.c_void_p()
    rr.ffi.check(lib.rr_host_alloc(stride * h, C.byref(host)))
    pinned = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_uint8)), shape=(h, stride))
    pinned[:] = 0xCD
    rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), host, 0))
    assert np.array_equal(pinned.reshape(-1)[: h * w * 3].reshape(h, w, 3), pageable)
    pinned[:] = 0xCD
    rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), host, stride))
    assert np.array_equal(pinned[:, : w * 3].reshape(h, w, 3), pageable) and (pinned[:, w * 3:] == 0xCD).all()
    pinned[:] = 0xCD
    for k in range(3):
        pk = ren.frame_params(16, k, 3)
        rr.ffi.check(lib.rr_render_rgb8_placed(scene.handle, C.byref(pk), host, stride))
    assert np.array_equal(pinned[:, : w * 3].reshape(h, w, 3), pageable) and (pinned[:, w * 3:] == 0xCD).all()
    del pinned
    lib.rr_host_free(host)
    scene.close()


def test_fence_signal_memset_read(rr):
    """The stand-alone publisher + wait pair (device-side barrier / doorbell primitive) and the raw device helpers."""
    import torch

    lib = rr.ffi.load()
    p = C.c_void_p()
    rr.ffi.check(lib.rr_device_alloc(0, 256, C.byref(p)))
    rr.ffi.check(lib.rr_device_memset(0, p, 0, 256))
    a, b = torch.cuda.Stream(), torch.cuda.Stream()
    # b waits for word 3 to reach 7; a publishes 5 (not enough), then 9
    rr.ffi.check(lib.rr_fence_wait_device(0, C.c_void_p(p.value + 12), 1, 7, 20000, C.c_void_p(p.value + 64), C.c_void_p(b.cuda_stream)))
    rr.ffi.check(lib.rr_fence_signal_device(0, C.c_void_p(p.value + 16), 1, C.c_void_p(b.cuda_stream)))   # ordered behind the wait
    rr.ffi.check(lib.rr_fence_signal_device(0, C.c_void_p(p.value + 12), 5, C.c_void_p(a.cuda_stream)))
    a.synchronize()
    words = (C.c_uint32 * 64)()
    rr.ffi.check(lib.rr_device_read(0, p, words, 256))
    assert words[3] == 5 and words[4] == 0          # the waiter is still spinning: its follow-up store has not happened
    rr.ffi.check(lib.rr_fence_signal_device(0, C.c_void_p(p.value + 12), 9, C.c_void_p(a.cuda_stream)))
    torch.cuda.synchronize()
    rr.ffi.check(lib.rr_device_read(0, p, words, 256))
    assert words[3] == 9 and words[4] == 1 and words[16] == 0   # released, no timeout recorded
    assert lib.rr_fence_signal_device(0, None, 1, None) == rr.ffi.RR_ERR_BAD_ARG
    rr.ffi.check(lib.rr_device_free(0, p))


@pytest.mark.parametrize("march", [False, True])
def test_overlapping_launches_of_one_handle(rr, march):
    """Kernels of ONE scene handle queued on several streams run concurrently; each launch (ray-trace AND ray-march mode)
    owns one of 64 (tile queue, block counter) slots that reset themselves, so 12 overlapping launches on 6 streams must
    all produce the frame, and the slot ring must survive wrapping."""
    import torch

    w, h = (1024, 576) if not march else (512, 288)
    ren = rr.default_scene(w, h, use_raymarching=march, glow_effect=1.0 if march else None)
    scene = rr.DeviceScene(ren, 0)
    p = ren.frame_params()
    ref = torch.empty(h * w * 3, dtype=torch.uint8, device="cuda:0")
    scene.render_rgb8_device(p, ref.data_ptr())
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(6)]
    bufs = [torch.zeros_like(ref) for _ in range(12)]
    for rep in range(7 if not march else 3):               # 84 launches: the 64-slot ring wraps
        for b in bufs:
            b.zero_()
        torch.cuda.synchronize()
        for i, b in enumerate(bufs):
            scene.render_rgb8_device(p, b.data_ptr(), stream=streams[i % 6].cuda_stream)
        torch.cuda.synchronize()
        for b in bufs:
            assert torch.equal(b, ref)
    scene.close()


@pytest.mark.parametrize("march", [False, True])
def test_concurrent_host_renders_of_one_handle(rr, march):
    """Eight host threads call the blocking rr_render_rgb8 on ONE handle (the web server's situation,
    webserver.rs:268-280): the handle runs up to four of them at a time on separate lanes; all frames must be right."""
    import threading

    w, h = 640, 360
    ren = rr.default_scene(w, h, use_raymarching=march, glow_effect=1.0 if march else None)
    scene = rr.DeviceScene(ren, 0)
    views = []
    for k in range(8):
        r = rr.default_scene(w, h, use_raymarching=march, glow_effect=1.0 if march else None)
        r.camera.position = np.asarray((0.0, -150.0 + 10.0 * k, -300.0 + 7.0 * k), dtype=np.float32)
        views.append(r.frame_params())
    ref = [scene.render_rgb8(v).copy() for v in views]
    assert not np.array_equal(ref[0], ref[7])
    out = [None] * 8
    err = []

    def work(k):
        try:
            for _ in range(3):
                out[k] = scene.render_rgb8(views[k])
        except Exception as e:  # noqa: BLE001
            err.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not err
    for k in range(8):
        assert np.array_equal(out[k], ref[k])
    scene.close()


def test_async_submit_wait(rr):
    """rr_render_rgb8_async / rr_render_wait (render_frames support): several frames of one handle in flight, waited for
    out of order; a fifth submission waits for a lane; stale tickets are rejected."""
    lib = rr.ffi.load()
    w, h = 800, 448
    ren = rr.default_scene(w, h)
    scene = rr.DeviceScene(ren, 0)
    n = 6
    views, hosts, tickets = [], [], []
    for k in range(n):
        r = rr.default_scene(w, h)
        r.camera.position = np.asarray((0.0, -150.0 + 12.0 * k, -300.0), dtype=np.float32)
        views.append(r.frame_params())
        ptr = C.c_void_p()
        rr.ffi.check(lib.rr_host_alloc(w * h * 3, C.byref(ptr)))
        hosts.append(ptr)
    ref = [scene.render_rgb8(v).copy() for v in views]
    for k in range(4):
        t = C.c_int32(-1)
        rr.ffi.check(lib.rr_render_rgb8_async(scene.handle, C.byref(views[k]), hosts[k], 0, C.byref(t)))
        assert t.value >= 0
        tickets.append(t.value)
    ms = C.c_float()
    for k in (2, 0):                                        # out of order
        rr.ffi.check(lib.rr_render_wait(scene.handle, tickets[k], C.byref(ms)))
        assert ms.value > 0.0
        assert lib.rr_render_wait(scene.handle, tickets[k], None) == rr.ffi.RR_ERR_BAD_ARG   # waited for twice
    for k in (4, 5):                                        # lanes freed above are reused
        t = C.c_int32(-1)
        rr.ffi.check(lib.rr_render_rgb8_async(scene.handle, C.byref(views[k]), hosts[k], 0, C.byref(t)))
        tickets.append(t.value)
    for k in (1, 3, 4, 5):
        rr.ffi.check(lib.rr_render_wait(scene.handle, tickets[k], None))
    for k in range(n):
        got = np.frombuffer(C.string_at(hosts[k], w * h * 3), dtype=np.uint8).reshape(h, w, 3)
        assert np.array_equal(got, ref[k])
        lib.rr_host_free(hosts[k])
    assert lib.rr_render_wait(scene.handle, 12345, None) == rr.ffi.RR_ERR_BAD_ARG
    scene.close()


def test_shared_device_frame_single_rank(rr):
    """ray_rust_b200.multi.SharedDeviceFrame with one rank (no process group): render + completion word + wait + download."""
    from ray_rust_b200.multi import SharedDeviceFrame

    ren = rr.default_scene(640, 360)
    scene = rr.DeviceScene(ren, 0)
    frame = SharedDeviceFrame(None, 0, 1, 0, 640, 360)
    for _ in range(3):
        frame.render(scene, ren)
    import torch

    torch.cuda.synchronize()
    assert frame.epoch == 3 and not frame.timed_out()
    assert np.array_equal(frame.download(), scene.render_rgb8(ren.frame_params()))
    frame.close()
    scene.close()
