"""Where the work of the ray-march frame (BASELINE config 3) is: march steps per pixel and per 8x4 warp tile, from the kernel
source compiled for the CPU.   python tools/march_cost_map.py [W H]   (11 s for 3840x2160 on 8 cores)"""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_rust_b200 as rr
so = "/tmp/libmarch_cost_map.so"
subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-o", so, os.path.join(ROOT, "tools", "march_cost_map.cpp")])
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ren = rr.default_scene(W, H, use_raymarching=True, glow_effect=1.0)
flat, p = ren.flatten(), ren.frame_params()
steps = np.zeros((H, W), np.uint32)
assert C.CDLL(so).march_cost_map(C.byref(flat.desc), C.byref(p), steps.ctypes.data_as(C.c_void_p), os.cpu_count() or 1) == 0
t = steps[:H // 4 * 4, :W // 8 * 8].reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3)).astype(np.int64)  # a warp tile costs its longest lane
print(f"{W}x{H}: {int(steps.sum())} march steps, longest pixel {int(steps.max())}; {t.size} warp tiles, sum of tile maxima {int(t.sum())}, "
      f"lane utilisation {steps.sum() / (32 * t.sum()):.3f}")
long_ = t >= 20000
print(f"tiles with >= 20 000 steps: {int(long_.sum())} = {100 * t[long_].sum() / t.sum():.1f} % of the frame's tile-steps, in tile rows "
      f"{int(np.where(long_.any(axis=1))[0].min())}..{int(np.where(long_.any(axis=1))[0].max())} of {t.shape[0]}")
q = np.sort(t.ravel())[::-1]
print("tile cost quantiles (steps):", {f"{100 * f:g}%": int(q[min(int(f * q.size), q.size - 1)]) for f in (0, 0.01, 0.017, 0.02, 0.05, 0.5)})
