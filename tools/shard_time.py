"""Kernel time of ONE rank's shard of a band-sharded frame, on one GPU (what bench.py reports as per_rank_kernel_ms):
    python tools/shard_time.py [trace8k|synth4k|trace4k] [band_count] [reps]
Environment knobs of the tile schedule (RR_STATIC_16THS, RR_SUB_TAIL_16THS) are read by the library at first launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ray_rust_b200 as rr
cfg = sys.argv[1] if len(sys.argv) > 1 else "trace8k"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 15
ren = {"trace8k": lambda: rr.default_scene(7680, 4320), "trace4k": lambda: rr.default_scene(3840, 2160),
       "synth4k": lambda: rr.synthetic_scene(3840, 2160)}[cfg]()
scene = rr.DeviceScene(ren, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
whole = ren.frame_params()
buf = torch.empty(whole.yres * whole.xres * 3, dtype=torch.uint8, device="cuda:0")
st = torch.cuda.Stream()  # not the default stream (handle 0 = NULL = blocking launch on the handle's own stream)
torch.cuda.set_stream(st)
prewarm = ren.frame_params(16, 0, 13) if os.environ.get("RR_PREWARM") else None  # a 1/13 shard: same kernel instance, other rows
def t(p):
    ms = []
    for i in range(reps + 3):
        if not os.environ.get('RR_NOFLUSH'): flush.fill_(i & 255)
        if prewarm is not None: scene.render_rgb8_device(prewarm, buf.data_ptr(), stream=st.cuda_stream)  # untimed: warms the SMs' instruction/constant caches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); scene.render_rgb8_device(p, buf.data_ptr(), stream=st.cuda_stream); b.record(st)
        torch.cuda.synchronize()
        if i >= 3: ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]
full = t(whole)
shards = [t(ren.frame_params(16, r, nb)) for r in (0, nb // 2, nb - 1)]
print(f"{cfg} prewarm={os.environ.get('RR_PREWARM','0')} noflush={os.environ.get('RR_NOFLUSH','0')} static16={os.environ.get('RR_STATIC_16THS','-')} subtail16={os.environ.get('RR_SUB_TAIL_16THS','-')}: full {full:.4f} ms, ideal shard {full/nb:.4f}, "
      f"shards(1/{nb}) {' '.join(f'{x:.4f}' for x in shards)}", flush=True)
scene.close()
