"""Render a few device-resident frames of one config (for ncu captures): python tools/one_frame.py synth4k [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ray_rust_b200 as rr
cfgs = {
    "trace4k": lambda: rr.default_scene(3840, 2160),
    "trace8k": lambda: rr.default_scene(7680, 4320),
    "march4k": lambda: rr.default_scene(3840, 2160, use_raymarching=True, glow_effect=1.0),
    "synth4k": lambda: rr.synthetic_scene(3840, 2160),
}
ren = cfgs[sys.argv[1]]()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene = rr.DeviceScene(ren, 0)
p = ren.frame_params()
buf = torch.empty(p.yres * p.xres * 3, dtype=torch.uint8, device="cuda:0")
for i in range(n):
    scene.render_rgb8_device(p, buf.data_ptr())
    print(sys.argv[1], "kernel ms", scene.last_kernel_ms(), flush=True)
scene.close()
