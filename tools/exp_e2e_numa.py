"""Experiment (torchrun, N ranks): what bounds the N-GPU e2e path (each rank copies its row bands into ONE shared
page-locked host frame)? Variants: A current; B barrier alone; C private pinned buffer per rank (no sharing);
D shared frame whose band pages are bound to the NUMA node of the GPU that writes them (mbind before first touch)."""
import ctypes as C, mmap, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ray_rust_b200 as rr

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = rr.ffi.load()
W, H, B = 7680, 4320, 16
ren = rr.default_scene(W, H)
scene = rr.DeviceScene(ren, lr)
p = ren.frame_params(B, rank, world)
frame_bytes = W * H * 3
libc = C.CDLL(None, use_errno=True)

def numa_of_gpu(i):
    bdf = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    if bdf is None:
        import pynvml
        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(i)).busId
        bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
    bdf = bdf.lower()
    if len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    try:
        return int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
    except Exception as e:
        return -1

def timeit(fn, n=30):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3 / n

def shm_frame(tag, nodes=None):
    path = f"/dev/shm/rr_exp_{tag}_{os.environ.get('MASTER_PORT','0')}"
    if rank == 0:
        with open(path, "wb") as f: f.truncate(frame_bytes)
    dist.barrier()
    f = open(path, "r+b"); m = mmap.mmap(f.fileno(), frame_bytes)
    arr = np.frombuffer(m, dtype=np.uint8)
    if nodes is not None and rank == 0:
        band_bytes = B * W * 3
        nb = (H + B - 1) // B
        ok = 0
        for b in range(nb):
            node = nodes[b % world]
            if node < 0: continue
            mask = C.c_ulong(1 << node)
            lo = b * band_bytes; ln = min(band_bytes, frame_bytes - lo)
            lo_al = lo // 4096 * 4096
            r = libc.syscall(237, C.c_void_p(arr.ctypes.data + lo_al), C.c_ulong(ln + lo - lo_al), 2, C.byref(mask), C.c_ulong(64), 0)
            ok += (r == 0)
        print(f"mbind ok for {ok}/{nb} bands, errno {C.get_errno()}", flush=True)
    if rank == 0:
        arr[:] = 0   # first touch after the policy is set
    dist.barrier()
    rr.ffi.check(lib.rr_host_register(C.c_void_p(arr.ctypes.data), frame_bytes))
    return path, f, m, arr

res = {}
pa, fa, ma, arr_a = shm_frame("a")
ha = C.c_void_p(arr_a.ctypes.data)
def step_a():
    rr.ffi.check(lib.rr_render_rgb8_placed(scene.handle, C.byref(p), ha, 0)); dist.barrier()
def step_a_nobar():
    rr.ffi.check(lib.rr_render_rgb8_placed(scene.handle, C.byref(p), ha, 0))
res["A shared shm + barrier"] = timeit(step_a)
res["A' shared shm, no barrier"] = timeit(step_a_nobar)
res["B barrier alone"] = timeit(lambda: dist.barrier())
rows = rr.frame_rows(p)
priv = C.c_void_p()
rr.ffi.check(lib.rr_host_alloc(rows * W * 3, C.byref(priv)))
def step_c():
    rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), priv, 0))
res["C private pinned, no barrier"] = timeit(step_c)
if rank == 0:
    for k, v in res.items(): print(f"{k}: {v:.3f} ms", flush=True)
dist.barrier()
lib.rr_host_unregister(ha)
del arr_a
if rank == 0:
    os.unlink(pa)
dist.destroy_process_group()
os._exit(0)
