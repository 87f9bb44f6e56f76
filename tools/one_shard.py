"""Render a few device-resident frames of ONE shard (for ncu captures of small launches): python tools/one_shard.py trace8k 16 [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ray_rust_b200 as rr
cfgs = {"trace4k": lambda: rr.default_scene(3840, 2160), "trace8k": lambda: rr.default_scene(7680, 4320), "synth4k": lambda: rr.synthetic_scene(3840, 2160)}
ren = cfgs[sys.argv[1]]()
nb = int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scene = rr.DeviceScene(ren, 0)
p = ren.frame_params(16, 0, nb)
buf = torch.empty(p.yres * p.xres * 3, dtype=torch.uint8, device="cuda:0")
for i in range(n):
    scene.render_rgb8_device(p, buf.data_ptr())
    print(sys.argv[1], nb, "kernel ms", scene.last_kernel_ms(), flush=True)
scene.close()
