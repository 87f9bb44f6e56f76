"""A/B kernel timings of several builds of libray_rust_b200.so on one box (works across ABI revisions: binds only the
entry points it needs, so a round-1 build can be timed beside the current one).

  python tools/ab_kernel.py [--cfg=trace4k,trace8k,synth4k,march4k] [--reps=N] lib1.so lib2.so ...   ("default" = in-tree build)

Each (lib, cfg) runs in its own subprocess, interleaved over two passes; prints median kernel ms, e2e ms into a pinned
frame, and the CRC32 of the frame (builds that claim bit-identical output must print the same CRC)."""
import ctypes as C
import os
import subprocess
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFGS = {
    "trace4k": lambda rr: rr.default_scene(3840, 2160),
    "trace8k": lambda rr: rr.default_scene(7680, 4320),
    "march4k": lambda rr: rr.default_scene(3840, 2160, use_raymarching=True, glow_effect=1.0),
    "synth4k": lambda rr: rr.synthetic_scene(3840, 2160),
    "trace1080": lambda rr: rr.default_scene(1920, 1080),
}


def worker(libpath, cfg, reps):
    import numpy as np
    from importlib import import_module

    scene_mod = import_module("ray_rust_b200").scene
    ffi = import_module("ray_rust_b200").ffi
    import ray_rust_b200 as rr

    lib = C.CDLL(libpath)
    P = C.c_void_p
    lib.rr_last_error.restype = C.c_char_p
    lib.rr_scene_create.argtypes = [C.POINTER(ffi.rr_scene_desc), C.c_int, C.POINTER(P)]
    lib.rr_scene_destroy.argtypes = [P]
    lib.rr_render_rgb8_device.argtypes = [P, C.POINTER(ffi.rr_frame_params), P, C.c_size_t, P]
    lib.rr_render_rgb8.argtypes = [P, C.POINTER(ffi.rr_frame_params), P, C.c_size_t]
    lib.rr_last_kernel_ms.argtypes = [P, C.POINTER(C.c_float)]
    lib.rr_host_alloc.argtypes = [C.c_size_t, C.POINTER(P)]
    lib.rr_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(P)]
    lib.rr_device_memset.argtypes = [C.c_int, P, C.c_int, C.c_size_t]

    def ck(rc):
        if rc != 0:
            raise RuntimeError(f"rc={rc}: {lib.rr_last_error().decode()}")

    ren = CFGS[cfg](rr)
    flat = ren.flatten()
    p = ren.frame_params()
    h = P()
    ck(lib.rr_scene_create(C.byref(flat.desc), 0, C.byref(h)))
    nbytes = p.xres * p.yres * 3
    dbuf, flush, host = P(), P(), P()
    ck(lib.rr_device_alloc(0, nbytes, C.byref(dbuf)))
    ck(lib.rr_device_alloc(0, 256 << 20, C.byref(flush)))
    ck(lib.rr_host_alloc(nbytes, C.byref(host)))
    ms = []
    for i in range(reps + 3):
        ck(lib.rr_device_memset(0, flush, i & 255, 256 << 20))  # evict the previous frame from L2
        ck(lib.rr_render_rgb8_device(h, C.byref(p), dbuf, 0, None))
        v = C.c_float()
        ck(lib.rr_last_kernel_ms(h, C.byref(v)))
        if i >= 3:
            ms.append(v.value)
    e2e = []
    for i in range(max(4, reps // 2)):
        t = time.perf_counter()
        ck(lib.rr_render_rgb8(h, C.byref(p), host, 0))
        e2e.append((time.perf_counter() - t) * 1e3)
    crc = zlib.crc32(C.string_at(host, nbytes))
    ms.sort(); e2e.sort()
    print(f"{os.path.basename(libpath):28s} {cfg:10s} kernel ms med {ms[len(ms) // 2]:.4f} min {ms[0]:.4f}  e2e ms med {e2e[len(e2e) // 2]:.3f} min {e2e[0]:.3f}  crc {crc:08x}",
          flush=True)
    lib.rr_scene_destroy(h)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--worker":
        return worker(args[1], args[2], int(args[3]))
    cfgs = ["trace4k", "synth4k"]
    reps = 15
    libs = []
    for a in args:
        if a.startswith("--cfg="):
            cfgs = a[6:].split(",")
        elif a.startswith("--reps="):
            reps = int(a[7:])
        else:
            libs.append(os.path.join(ROOT, "ray-rust_b200", "libray_rust_b200.so") if a == "default" else os.path.join(ROOT, a))
    for _pass in range(2):
        for cfg in cfgs:
            for lib in libs:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", lib, cfg, str(reps)], capture_output=True, text=True)
                sys.stdout.write(r.stdout)
                if r.returncode != 0:
                    print(f"{lib} {cfg}: FAILED\n{r.stderr[-2000:]}", flush=True)


if __name__ == "__main__":
    main()
