"""Lane-utilisation model of the BVH trace instance (tools/simt_model.cpp): python tools/simt_model.py [W H]"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_rust_b200 as rr
so = "/tmp/libsimt_model.so"
subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tools", "simt_model.cpp")])
lib = C.CDLL(so)
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ren = rr.synthetic_scene(W, H)
flat, p = ren.flatten(), ren.frame_params()
shapes = [(8, 4, 0), (32, 1, 0), (8, 8, 1), (16, 8, 1), (16, 16, 1), (32, 16, 1), (32, 32, 1), (64, 32, 1)]
arr = (C.c_int * (3 * len(shapes)))(*[v for s in shapes for v in s])
out = (C.c_double * (4 * len(shapes)))()
assert lib.model_run(C.byref(flat.desc), C.byref(p), arr, len(shapes), out) == 0
base = out[2]
for i, s in enumerate(shapes):
    print(f"tile {s[0]:3d}x{s[1]:<3d} {'refill' if s[2] else 'static'}: lane work utilisation {out[4*i]:.3f}  ray-slot occupancy {out[4*i+1]:.3f}  "
          f"modelled cost {out[4*i+2]/base:.3f} of 8x4 static  rays {int(out[4*i+3])}")
