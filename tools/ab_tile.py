"""A/B: 8x4 warp tiles (packed output) vs 32x1 row tiles (placed output) on one GPU, kernel-only ms."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ray_rust_b200 as rr
lib = rr.ffi.load()
for (w, h) in [(3840, 2160), (7680, 4320)]:
    ren = rr.default_scene(w, h)
    scene = rr.DeviceScene(ren, 0)
    p = ren.frame_params()
    buf = torch.empty(w * h * 3, dtype=torch.uint8, device="cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    res = {}
    for name in ("8x4", "32x1", "8x4", "32x1"):
        ms = []
        for i in range(12):
            flush.fill_(i); torch.cuda.synchronize()
            if name == "8x4":
                scene.render_rgb8_device(p, buf.data_ptr())
            else:
                rr.ffi.check(lib.rr_render_rgb8_placed_device(scene.handle, C.byref(p), C.c_void_p(buf.data_ptr()), 0, None))
            ms.append(scene.last_kernel_ms())
        res.setdefault(name, []).append(sorted(ms[2:])[len(ms[2:]) // 2])
    print(w, h, res)
    scene.close()
