"""What the L2 flush in front of a timed step costs the step itself (bench.py timing method, one GPU):
    python tools/exp_flush_gap.py [trace4k|trace8k]
variants: fill (torch fill_ of 256 MiB directly in front of the first event, bench.py's method), fill+sleep N us (a spin kernel
between the flush and the first event: the flush's write-backs drain before the timed region starts, the queue stays
non-empty so the render kernel still launches back to back), memset (cudaMemsetAsync as the flush), none (no flush)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import ray_rust_b200 as rr

cfg = sys.argv[1] if len(sys.argv) > 1 else "trace4k"
W, H = (3840, 2160) if cfg == "trace4k" else (7680, 4320)
ren = rr.default_scene(W, H)
scene = rr.DeviceScene(ren, 0)
p = ren.frame_params()
dev = torch.device("cuda", 0)
buf = torch.empty(H * W * 3, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(device=dev)  # NOT the default stream: its handle is 0, which the C ABI reads as "use the handle's own stream and block"
torch.cuda.set_stream(stream)
assert stream.cuda_stream != 0
sptr = C.c_void_p(stream.cuda_stream)
lib = rr.ffi.load()
clock_khz = torch.cuda.get_device_properties(0).clock_rate if hasattr(torch.cuda.get_device_properties(0), "clock_rate") else 1965000


def run(kind, sleep_us=0, n=60):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for i in range(n):
        if kind == "fill":
            flush.fill_(i & 0xFF)
        elif kind == "memset":
            rr.ffi.check(lib.rr_device_memset(0, C.c_void_p(flush.data_ptr()), i & 0xFF, 256 << 20))
        if sleep_us:
            torch.cuda._sleep(int(sleep_us * 1965))
        ev[i][0].record(stream)
        rr.ffi.check(lib.rr_render_rgb8_device(scene.handle, C.byref(p), C.c_void_p(buf.data_ptr()), 0, sptr))
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev[5:])
    return statistics.mean(t), t[len(t) // 2], t[0]


for _ in range(2):
    for kind, sl in (("fill", 0), ("fill", 5), ("fill", 20), ("fill", 50), ("memset", 0), ("none", 0), ("none", 20)):
        m, med, mn = run(kind, sl)
        print(f"{cfg} flush={kind:6s} sleep={sl:3d}us: mean {m:.4f} med {med:.4f} min {mn:.4f} ms", flush=True)
scene.close()
