"""Large randomised device-vs-oracle sweep (same generator as tests/test_random_scenes_gpu.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ray_rust_b200 as rr
from oracle import binding as ob
from test_random_scenes_gpu import _random_env
n_trace, n_march = int(sys.argv[1]), int(sys.argv[2])
bad = []
for march, n in ((False, n_trace), (True, n_march)):
    for seed in range(n):
        ren = _random_env(rr, 5000 + seed + (100000 if march else 0), march)
        ref = ob.render(ren, threads=os.cpu_count(), want_f32=True, want_tags=True, want_counts=True)
        sc = rr.DeviceScene(ren, 0); p = ren.frame_params()
        f = sc.render_f32(p); u8, cnt = sc.render_count(p); sc.close()
        d = np.abs(u8.astype(int) - ref["u8"].astype(int)).max(initial=0)
        glow = ren._glow_effect is not None
        clean = (ref["tags"] & 1) == 0 if not glow else np.zeros_like(ref["tags"], dtype=bool)
        a, b = f.view(np.uint32)[clean], ref["f32"].view(np.uint32)[clean]
        na, nb = np.isnan(f[clean]), np.isnan(ref["f32"][clean])
        ok = d <= 1 and np.array_equal(na, nb) and np.array_equal(a[~na], b[~nb]) and cnt.as_dict() == ref["counts"].as_dict()
        if not ok:
            bad.append((march, seed, int(d)))
            print("MISMATCH", march, seed, d, flush=True)
print("done: trace", n_trace, "march", n_march, "mismatching scenes:", bad)
