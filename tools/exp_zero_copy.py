"""Experiment: let the render kernel store the RGB8 frame straight into pinned (mapped) host memory instead of
device memory + cudaMemcpy. Compares wall time of: A rr_render_rgb8 (kernel + D2H copy), B zero-copy 8x4 tiles,
C zero-copy 32x1 row tiles (placed instance). Checks the bytes."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ray_rust_b200 as rr
lib = rr.ffi.load()
cfgs = {
    "trace4k": lambda: rr.default_scene(3840, 2160),
    "march4k": lambda: rr.default_scene(3840, 2160, use_raymarching=True, glow_effect=1.0),
    "synth4k": lambda: rr.synthetic_scene(3840, 2160),
}
for name in sys.argv[1:] or list(cfgs):
    ren = cfgs[name](); scene = rr.DeviceScene(ren, 0); p = ren.frame_params()
    n = p.xres * p.yres * 3
    hosts = []
    for _ in range(3):
        h = C.c_void_p(); rr.ffi.check(lib.rr_host_alloc(n, C.byref(h))); hosts.append(h)
    st = torch.cuda.Stream(); sp = C.c_void_p(st.cuda_stream)
    def a(): rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), hosts[0], 0))
    def b():
        rr.ffi.check(lib.rr_render_rgb8_device(scene.handle, C.byref(p), hosts[1], 0, sp)); st.synchronize()
    def c():
        rr.ffi.check(lib.rr_render_rgb8_placed_device(scene.handle, C.byref(p), hosts[2], 0, sp)); st.synchronize()
    out = {}
    for rep in range(2):
        for k, fn in (("A copy", a), ("B zero-copy 8x4", b), ("C zero-copy 32x1", c)):
            for _ in range(2): fn()
            t0 = time.perf_counter()
            for _ in range(8): fn()
            out.setdefault(k, []).append((time.perf_counter() - t0) / 8 * 1e3)
    arrs = [np.ctypeslib.as_array(C.cast(h, C.POINTER(C.c_uint8)), shape=(n,)) for h in hosts]
    same = bool(np.array_equal(arrs[0], arrs[1]) and np.array_equal(arrs[0], arrs[2]))
    print(name, {k: ["%.3f" % v for v in vs] for k, vs in out.items()}, "identical" if same else "MISMATCH", flush=True)
    for h in hosts: lib.rr_host_free(h)
    scene.close()
