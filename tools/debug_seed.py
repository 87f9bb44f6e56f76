import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ray_rust_b200 as rr
from oracle import binding as ob
from test_random_scenes_gpu import _random_env

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
march = len(sys.argv) > 2 and sys.argv[2] == "march"
ren = _random_env(rr, seed, march)

def diff(ren, culling=True):
    ref = ob.render(ren, threads=8, want_f32=True, want_tags=True)
    sc = rr.DeviceScene(ren, 0); sc.set_culling(culling)
    u8 = sc.render_rgb8(ren.frame_params()); f = sc.render_f32(ren.frame_params()); sc.close()
    d = np.abs(u8.astype(int) - ref["u8"].astype(int)).max(axis=2)
    return d, ref, f

d, ref, f = diff(ren)
print("objects", len(ren._objects), "kinds", [o.kind for o in ren._objects], "refl/refr", ren.max_reflections, ren.max_refractions, "res", ren.xres, ren.yres)
print("bad pixels", int((d > 1).sum()), "max", d.max())
d2, _, _ = diff(ren, culling=False)
print("culling off: bad", int((d2 > 1).sum()))
# greedy removal of objects while the mismatch persists
objs = list(ren._objects)
i = 0
while i < len(objs):
    trial = objs[:i] + objs[i + 1:]
    ren.objects(trial)
    dd, _, _ = diff(ren)
    if (dd > 1).sum() > 0:
        objs = trial
    else:
        i += 1
ren.objects(objs)
d, ref, f = diff(ren)
print("minimal scene:", len(objs), "objects; bad", int((d > 1).sum()))
for o in objs:
    m = o.material
    print(" ", "sphere" if o.kind == 0 else "floor", "r", o.r, "org", o.org, "n", getattr(o, "face_normal", None), "uv", o._uvmap,
          "| mat", m.name, "dif", tuple(m.diffuse), "spec", tuple(m.specular), "pn", m.pn, "t", m.t, "n", m.n, m._pattern, m._pattern_scale,
          "tex", None if m._texture is None else m._texture.shape, m.texture_filter)
ys, xs = np.nonzero(d > 1)
for y, x in list(zip(ys, xs))[:6]:
    print("px", x, y, "dev", f[y, x], "ref", ref["f32"][y, x], "tags", bin(ref["tags"][y, x]))
print("cam", ren.camera.position, ren.camera.rotation.as_tuple(), "light", ren._light)
