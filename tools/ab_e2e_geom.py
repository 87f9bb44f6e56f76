"""A/B e2e chunk plans (RR_E2E_GEOM="d,g" geometric vs the uniform default) in subprocesses on the same box."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import ctypes as C, time, sys
sys.path.insert(0, %r)
import ray_rust_b200 as rr
lib = rr.ffi.load()
for (w,h) in ((3840,2160),(7680,4320)):
    ren = rr.default_scene(w,h); scene = rr.DeviceScene(ren,0); p = ren.frame_params()
    host = C.c_void_p(); rr.ffi.check(lib.rr_host_alloc(w*h*3, C.byref(host)))
    for _ in range(5): rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), host, 0))
    ts=[]
    for _ in range(40):
        t=time.perf_counter(); rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), host, 0)); ts.append(time.perf_counter()-t)
    ts.sort(); print("%%dx%%d median %%.3f ms min %%.3f ms" %% (w,h,ts[20]*1e3, ts[0]*1e3), end="; ")
print()
''' % root
plans = [None, "48,1.5", "64,1.7", "96,1.7", "96,2.0", "128,1.8", "64,2.0", "192,2.0", "32,1.5"]
for rep in range(2):
    for g in plans:
        env = dict(os.environ)
        if g: env["RR_E2E_GEOM"] = g
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(f"geom={g}:", out.stdout.strip() or out.stderr[-300:], flush=True)
