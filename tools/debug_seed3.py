import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ray_rust_b200 as rr
from oracle import binding as ob
from test_random_scenes_gpu import _random_env
for seed in (1000, 1001, 1002, 1003, 1004, 1005, 1006, 1007):
    ren = _random_env(rr, seed, False)
    ref = ob.render(ren, threads=8, want_f32=True, want_tags=True)
    clean = (ref["tags"] & 1) == 0
    res = []
    for cull in (True, False):
        sc = rr.DeviceScene(ren, 0); sc.set_culling(cull)
        f = sc.render_f32(ren.frame_params()); sc.close()
        res.append(int(((f.view(np.uint32) != ref["f32"].view(np.uint32)).any(axis=2) & clean).sum()))
    ns = sum(1 for o in ren._objects if o.kind == 0)
    print("seed", seed, "spheres", ns, "floors", len(ren._objects) - ns, "bad with BVH", res[0], "brute force", res[1])
