"""Render one frame of a scene over N GPUs of one box and save it (rank 0), with the 1-GPU frame as the check.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/render_multi_gpu.py 7680 4320 [-m] [-g G] [-d scene.yaml] [-o out.png]

Uses ray_rust_b200.multi.SharedDeviceFrame: rows and completion words go into rank 0's memory over NVLink, no collective.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import ray_rust_b200 as rr
from ray_rust_b200.multi import SharedDeviceFrame


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("width", type=int)
    ap.add_argument("height", type=int)
    ap.add_argument("-m", "--raymarch", action="store_true")
    ap.add_argument("-g", "--gloweffect", type=float, default=None)
    ap.add_argument("-d", "--deserialize_file", default=None)
    ap.add_argument("-o", "--output", default="foo.png")
    ap.add_argument("--frames", type=int, default=5)
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ren = rr.default_scene(a.width, a.height, use_raymarching=a.raymarch, glow_effect=a.gloweffect)
    if a.deserialize_file:
        ren.deserialize(open(a.deserialize_file).read())
    scene = rr.DeviceScene(ren, lr)
    frame = SharedDeviceFrame(dist, rank, world, lr, a.width, a.height)
    stream = torch.cuda.Stream()  # not the default stream (handle 0 = NULL = blocking launch on the handle's own stream)
    torch.cuda.set_stream(stream)
    ms = []
    for _ in range(a.frames):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        frame.render(scene, ren, stream.cuda_stream)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
    ok = True
    if rank == 0:
        img = frame.download()
        single = scene.render_rgb8(ren.frame_params())
        ok = bool(np.array_equal(img, single)) and not frame.timed_out()
        from PIL import Image

        Image.fromarray(img).save(a.output)
        print(f"{a.width}x{a.height} on {world} GPU(s): frame ms (host clock around launch + sync) {['%.3f' % m for m in ms]}; "
              f"{'identical to the 1-GPU frame' if ok else 'MISMATCH'} -> {a.output}", flush=True)
    frame.close()
    scene.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
