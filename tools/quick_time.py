"""Quick device timings of the BASELINE configs (kernel-only via rr_last_kernel_ms, e2e via wall clock)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ray_rust_b200 as rr

lib = rr.ffi.load()
a, b = C.c_float(), C.c_float()
rr.ffi.check(lib.rr_fp32_peak_tflops(0, C.byref(a), C.byref(b)))
print(f"fp32 peak: unfused {a.value:.2f} TFLOP/s, ffma {b.value:.2f} TFLOP/s", flush=True)

which = sys.argv[1:] or ["trace4k", "trace8k", "march4k", "synth4k"]
cfgs = {
    "trace4k": lambda: rr.default_scene(3840, 2160),
    "trace8k": lambda: rr.default_scene(7680, 4320),
    "march4k": lambda: rr.default_scene(3840, 2160, use_raymarching=True, glow_effect=1.0),
    "synth4k": lambda: rr.synthetic_scene(3840, 2160),
    "synthmarch1080": lambda: rr.synthetic_scene(1920, 1080, use_raymarching=True, glow_effect=1.0),
}
for name in which:
    ren = cfgs[name]()
    scene = rr.DeviceScene(ren, 0)
    p = ren.frame_params()
    buf = torch.empty(p.yres * p.xres * 3, dtype=torch.uint8, device="cuda:0")
    ms = []
    for i in range(6):
        scene.render_rgb8_device(p, buf.data_ptr())
        ms.append(scene.last_kernel_ms())
    host = C.c_void_p()
    rr.ffi.check(lib.rr_host_alloc(p.yres * p.xres * 3, C.byref(host)))
    e2e = []
    for i in range(5):
        t = time.perf_counter()
        rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), host, 0))
        e2e.append((time.perf_counter() - t) * 1e3)
    _, cnt = scene.render_count(p, want_image=False)
    rays = cnt.rays()
    print(f"{name}: kernel ms {['%.3f' % m for m in ms]} e2e ms {['%.3f' % m for m in e2e]} rays {rays} "
          f"-> {rays / min(ms[1:]) / 1e3:.1f} Mrays/s kernel, {rays / min(e2e[1:]) / 1e3:.1f} Mrays/s e2e; counts {cnt.as_dict()}",
          flush=True)
    lib.rr_host_free(host)
    scene.close()
