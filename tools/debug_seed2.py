import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ray_rust_b200 as rr
from oracle import binding as ob
from test_random_scenes_gpu import _random_env

def bad(ren, f32=True):
    ref = ob.render(ren, threads=8, want_f32=True, want_tags=True)
    sc = rr.DeviceScene(ren, 0)
    f = sc.render_f32(ren.frame_params()); u8 = sc.render_rgb8(ren.frame_params()); sc.close()
    clean = (ref["tags"] & 1) == 0
    b = (f.view(np.uint32) != ref["f32"].view(np.uint32)).any(axis=2) & clean
    d8 = np.abs(u8.astype(int) - ref["u8"].astype(int)).max(axis=2)
    return int(b.sum()), int((d8 > 1).sum()), b, ref, f

ren = _random_env(rr, 1000, False)
print("base", bad(ren)[:2], "refl/refr", ren.max_reflections, ren.max_refractions)
for refl, refr in [(0, 0), (1, 0), (3, 0), (0, 5), (3, 10), (5, 11)]:
    ren.max_reflections, ren.max_refractions = refl, refr
    print("refl", refl, "refr", refr, bad(ren)[:2])
ren = _random_env(rr, 1000, False)
objs = list(ren._objects)
floor_i = [i for i, o in enumerate(objs) if o.kind == 1]
print("floor index", floor_i, "n objects", len(objs))
# move the floor to index 0
o2 = [objs[floor_i[0]]] + [o for i, o in enumerate(objs) if i != floor_i[0]]
ren.objects(o2); print("floor first:", bad(ren)[:2])
# normalise the floor normal
ren = _random_env(rr, 1000, False)
for o in ren._objects:
    if o.kind == 1:
        n = np.array(o.face_normal, dtype=np.float64); n /= np.linalg.norm(n); o.face_normal = tuple(np.float32(x) for x in n)
print("unit normal:", bad(ren)[:2])
ren = _random_env(rr, 1000, False)
for o in ren._objects: o.material.pn = 0
print("pn=0:", bad(ren)[:2])
ren = _random_env(rr, 1000, False)
for o in ren._objects: o.material.pn = 2
print("pn=2:", bad(ren)[:2])
ren = _random_env(rr, 1000, False)
nb, n8, b, ref, f = bad(ren)
ys, xs = np.nonzero(b)
for y, x in list(zip(ys, xs))[:8]:
    print("px", x, y, "dev", f[y, x], "ref", ref["f32"][y, x], "tags", bin(ref["tags"][y, x]))
