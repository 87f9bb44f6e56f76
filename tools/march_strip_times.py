"""How long is the longest warp tile of every 16-row strip of the ray-march frame? Each strip is rendered alone (band
parameters: band k of yres/16), so it has fewer tiles than the GPU has resident warps and its kernel time is the time of its
slowest tile running (nearly) by itself: the dependent-chain profile of the image, row by row.
    python tools/march_strip_times.py [W H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ray_rust_b200 as rr
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ren = rr.default_scene(W, H, use_raymarching=True, glow_effect=1.0)
scene = rr.DeviceScene(ren, 0)
n = H // 16
buf = torch.empty(16 * W * 3, dtype=torch.uint8, device="cuda:0")
whole = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda:0")
scene.render_rgb8_device(ren.frame_params(), whole.data_ptr())
scene.render_rgb8_device(ren.frame_params(), whole.data_ptr())
print(f"whole frame {scene.last_kernel_ms():.3f} ms")
tot = 0.0
out = []
for k in range(n):
    p = ren.frame_params(16, k, n)
    scene.render_rgb8_device(p, buf.data_ptr())
    scene.render_rgb8_device(p, buf.data_ptr())
    ms = scene.last_kernel_ms()
    tot += ms
    out.append(ms)
print(f"sum of strips {tot:.2f} ms; longest strip {max(out):.3f} ms")
for k in range(n):
    print(f"rows {16*k:4d}-{16*k+15:4d}: {out[k]:.3f} ms")
scene.close()
