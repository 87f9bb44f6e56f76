"""A/B the e2e chunking parameters (env RR_E2E_MAX_CHUNKS / RR_E2E_CHUNK_KB) in subprocesses on the same box."""
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
    for _ in range(30):
        t=time.perf_counter(); rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), host, 0)); ts.append(time.perf_counter()-t)
    ts.sort(); print("%%dx%%d median %%.3f ms min %%.3f ms" %% (w,h,ts[15]*1e3, ts[0]*1e3), end="; ")
print()
''' % root
for rep in range(2):
    for mc, kb in ((8, 4096), (16, 1536), (16, 1024), (24, 1024), (32, 768), (32, 512)):
        env = dict(os.environ, RR_E2E_MAX_CHUNKS=str(mc), RR_E2E_CHUNK_KB=str(kb))
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(f"chunks<={mc} chunk>={kb}KB:", out.stdout.strip() or out.stderr[-300:], flush=True)
