#!/usr/bin/env python
"""bench.py — headline benchmark of the per-pixel tracing path (BASELINE.json metric: Mrays/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" renders one whole frame of the workload:
  N=1  "default-4k-trace"  = BASELINE.json configs[1]: built-in scene, 3840x2160, ray-trace mode.
  N>1  "default-8k-trace-bands" = configs[4]: built-in scene, 7680x4320, interleaved 16-row bands over
       the N ranks (one process per GPU). Every rank's render kernel stores its rows straight into rank 0's
       frame over NVLink peer memory (CUDA IPC); the step ends with a 1-element NCCL all-reduce that orders
       completion. The NCCL-gather + un-interleave alternative is timed too and reported under "alt".
`value` = reference-equivalent rays (one ray = one scene-level raycast(); counted by the
instrumented kernel and checked against the oracle in tests/) per second of device time with the scene
resident in HBM; `e2e` = same through the C ABI into HOST memory (wall clock, includes the D2H of the
frame; for N>1 each rank copies its bands into one shared page-locked host frame over its own PCIe link).
The reference arm (--impl reference) times the CPU oracle (oracle/, the C++ restatement of the reference:
no Rust toolchain exists here) on all host cores.
"""
import argparse
import ctypes as C
import json
import mmap
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, march, glow, scene)
    "default-4k-trace": (3840, 2160, False, None, "default"),
    "default-8k-trace-bands": (7680, 4320, False, None, "default"),
    "default-8k-trace": (7680, 4320, False, None, "default"),
    "default-4k-march-glow": (3840, 2160, True, 1.0, "default"),
    "synthetic1024-4k-trace": (3840, 2160, False, None, "synthetic"),
    "default-640x480-trace": (640, 480, False, None, "default"),
}
BAND_ROWS = 16

# ALGORITHMIC flops per unit of work (1 flop = one f32 add/sub/mul/div/sqrt/min/max/neg)
FLOPS = dict(pixel_setup=71, sphere_test=19, floor_test=15, hit_point=6, sphere_normal=12, shading=54,
             refract=38, bounce=28, bg=27 + 6, quantise=9, sphere_dist=12, floor_dist=10, march_step=7, glow=5)


def algorithmic_flops(c, march):
    """flops(frame) = sum(counter x constant) with the counters of the reference's brute-force algorithm."""
    floor_tests = c["object_tests"] - c["sphere_tests"]
    hits = c["shadow"]  # one shading() call per hit, one shadow ray per shading() call
    f = FLOPS
    total = c["pixels"] * (f["pixel_setup"] + f["quantise"])
    if march:
        total += c["sphere_tests"] * f["sphere_dist"] + floor_tests * f["floor_dist"] + c["march_steps"] * f["march_step"]
        total += (c["primary"] + c["refract"]) * f["glow"]
    else:
        total += c["sphere_tests"] * f["sphere_test"] + floor_tests * f["floor_test"]
    total += hits * (f["hit_point"] + f["shading"] + f["bounce"]) + c["sphere_hits"] * f["sphere_normal"]
    total += c["refract"] * f["refract"] + c["bg_evals"] * f["bg"]
    return int(total)


def make_env(rr, name):
    w, h, march, glow, kind = WORKLOADS[name]
    if kind == "synthetic":
        return rr.synthetic_scene(w, h, use_raymarching=march, glow_effect=glow)
    return rr.default_scene(w, h, use_raymarching=march, glow_effect=glow)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_arm(ob, ren, threads, reps, params=None):
    """Time the oracle (CPU restatement of the reference path) on `threads` host threads."""
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        ob.render(ren, params=params, threads=threads, want_u8=True)
        dt = time.perf_counter() - t
        best = dt if best is None or dt < best else best
    return best


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ray_rust_b200 as rr
    from oracle import binding as ob

    name = args.workload or ("default-4k-trace" if args.gpus == 1 else "default-8k-trace-bands")
    ren = make_env(rr, name)
    threads = os.cpu_count() or 1
    counts = ob.render(ren, threads=threads, want_u8=False, want_counts=True)["counts"]
    rays = counts.rays()
    for _ in range(args.warmup):
        ob.render(ren, threads=threads)
    times = []
    for _ in range(args.steps):
        t = time.perf_counter()
        ob.render(ren, threads=threads)
        times.append(time.perf_counter() - t)
    ms = 1e3 * sum(times) / len(times)
    v = rays / (ms * 1e-3) / 1e6
    w, h = WORKLOADS[name][:2]
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "width": w, "height": h, "rays_per_frame": rays,
                   "note": "CPU oracle (C++ restatement of the reference's render(); the Rust binary cannot be built here), "
                           "render only, -t = host cores"},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full {w}x{h} frames, mean"},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--band-rows", type=int, default=None, help="rows per interleaved band of the multi-GPU shard (default 16)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps or 10
        args.warmup = 1 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = args.steps or 100
    args.warmup = 5 if args.warmup is None else max(3, args.warmup)

    import numpy as np
    import torch

    import ray_rust_b200 as rr
    from ray_rust_b200 import bands

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dist = None
    if world > 1:
        # NCCL_DEBUG=VERSION prints "NCCL version ..." on stdout, which must carry exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        # communicator creation prints "NCCL version ..." on stdout; keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device=torch.device("cuda", local_rank))
            dist.all_reduce(warm)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = rr.ffi.load()

    name = args.workload or ("default-4k-trace" if world == 1 else "default-8k-trace-bands")
    W, H, march, glow, _ = WORKLOADS[name]
    ren = make_env(rr, name)
    scene = rr.DeviceScene(ren, local_rank)
    sharded = world > 1
    band_rows = args.band_rows or BAND_ROWS
    p = ren.frame_params(band_rows, rank, world) if sharded else ren.frame_params()
    my_rows = rr.frame_rows(p)
    max_rows = bands.max_shard_rows(H, band_rows, world) if sharded else H
    shard_bytes = max_rows * W * 3
    frame_bytes = H * W * 3

    # ---- ray counts of the whole frame (reference-equivalent; instrumented kernel, untimed) ----
    _, cnt = scene.render_count(ren.frame_params(), want_image=False)
    counts = cnt.as_dict()
    rays = cnt.rays()
    flops = algorithmic_flops(counts, march)

    stream = torch.cuda.current_stream(dev)
    sptr = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- buffers -------------------------------------------------------------------------------
    frame_ptr = C.c_void_p()          # the one frame: local on rank 0, NVLink peer mapping elsewhere
    out = None
    if sharded:
        handle = (C.c_uint8 * 64)()
        # one allocation on rank 0: the frame, then the completion words (one uint32 per rank) and a status word
        flags_off = (frame_bytes + 255) // 256 * 256
        if rank == 0:
            rr.ffi.check(lib.rr_device_alloc(local_rank, flags_off + 512, C.byref(frame_ptr)))
            rr.ffi.check(lib.rr_device_memset(local_rank, C.c_void_p(frame_ptr.value + flags_off), 0, 512))
            rr.ffi.check(lib.rr_ipc_export(frame_ptr, handle))
        box = [bytes(handle)]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            hb = (C.c_uint8 * 64).from_buffer_copy(box[0])
            rr.ffi.check(lib.rr_ipc_open(local_rank, hb, C.byref(frame_ptr)))
        flags_ptr = C.c_void_p(frame_ptr.value + flags_off)            # completion words, one per rank
        status_ptr = C.c_void_p(frame_ptr.value + flags_off + 256)   # timeout status of the wait kernels
        arrive_ptr = C.c_void_p(frame_ptr.value + flags_off + 128)   # device-side barrier: arrival words, one per rank
        go_ptr = C.c_void_p(frame_ptr.value + flags_off + 192)       # ... and the release word
        epoch = [0]
        gate = [0]

        def device_barrier():
            """Aligns the ranks' streams on the device (flag words in rank 0's memory, written and polled over NVLink):
            every rank publishes its arrival, rank 0 waits for all of them and opens the gate, every rank waits for the
            gate. A NCCL barrier lets the ranks' streams resume tens of microseconds apart; this one a few. Untimed."""
            gate[0] += 1
            rr.ffi.check(lib.rr_fence_signal_device(local_rank, C.c_void_p(arrive_ptr.value + 4 * rank), gate[0], sptr))
            if rank == 0:
                rr.ffi.check(lib.rr_fence_wait_device(local_rank, arrive_ptr, world, gate[0], 5000, status_ptr, sptr))
                rr.ffi.check(lib.rr_fence_signal_device(local_rank, go_ptr, gate[0], sptr))
            rr.ffi.check(lib.rr_fence_wait_device(local_rank, go_ptr, 1, gate[0], 5000, status_ptr, sptr))
        tok = torch.zeros(1, dtype=torch.float32, device=dev)
        packed = torch.empty(shard_bytes, dtype=torch.uint8, device=dev)
        gathered = torch.empty(world * shard_bytes, dtype=torch.uint8, device=dev) if rank == 0 else None
        gframe = torch.empty(frame_bytes, dtype=torch.uint8, device=dev) if rank == 0 else None
    else:
        out = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)

    def step_main():
        """one frame, device resident. returns number of this repo's kernels launched"""
        if not sharded:
            scene.render_rgb8_device(p, out.data_ptr(), stream=stream.cuda_stream)
            return 1
        # fused: the render kernel stores its rows into rank 0's frame over NVLink AND publishes its completion word
        # there; rank 0's stream waits on the words. No collective in the step.
        epoch[0] += 1
        rr.ffi.check(lib.rr_render_rgb8_placed_signal_device(scene.handle, C.byref(p), frame_ptr, W * 3, flags_ptr, epoch[0], sptr))
        if rank == 0:
            rr.ffi.check(lib.rr_fence_wait_device(local_rank, flags_ptr, world, epoch[0], 5000, status_ptr, sptr))
            return 2
        return 1

    def step_allreduce_fence():
        rr.ffi.check(lib.rr_render_rgb8_placed_device(scene.handle, C.byref(p), frame_ptr, W * 3, sptr))
        dist.all_reduce(tok)  # completion fence: rank 0's stream passes it only after every rank's kernel
        return 1

    def step_gather():
        scene.render_rgb8_device(p, packed.data_ptr(), stream=stream.cuda_stream)
        glist = list(gathered.chunk(world)) if rank == 0 else None
        dist.gather(packed, glist, dst=0)
        if rank == 0:
            rr.ffi.check(lib.rr_bands_unpack_device(C.byref(p), C.c_void_p(gathered.data_ptr()), shard_bytes,
                                                    C.c_void_p(gframe.data_ptr()), sptr))
            return 2
        return 1

    def timed(step_fn, steps, sample_clocks):
        for _ in range(args.warmup):
            step_fn()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        launches = 0
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xFF)  # evict the previous frame from L2 (not timed)
            if dist is not None:
                dist.barrier()
                device_barrier()
            ev[i][0].record(stream)
            launches += step_fn()
            ev[i][1].record(stream)
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()) / steps, launches, clocks, wall

    ms_per_step, launches, clocks, wall_s = timed(step_main, args.steps, True)
    value = rays / (ms_per_step * 1e-3) / 1e6

    # kernel-only time of this rank's render launch (roofline numerator), same stream, same L2 hygiene
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 30))]
    kbuf = packed if sharded else out
    for a, b in kev:
        flush.fill_(1)
        a.record(stream)
        scene.render_rgb8_device(p, kbuf.data_ptr(), stream=stream.cuda_stream)
        b.record(stream)
    torch.cuda.synchronize(dev)
    kt = torch.tensor([sum(a.elapsed_time(b) for a, b in kev) / len(kev)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    kernel_ms = float(kt.item())

    alt = None
    frame_check = None
    if sharded:
        gms, _, _, _ = timed(step_gather, max(5, args.steps // 2), False)
        alt = {"method": "nccl gather to rank 0 + rr_bands_unpack_device", "ms_per_step": gms, "value": rays / (gms * 1e-3) / 1e6}
        ams, _, _, _ = timed(step_allreduce_fence, max(5, args.steps // 2), False)
        alt = [alt, {"method": "placed NVLink stores + 1-element NCCL all-reduce as the completion fence", "ms_per_step": ams,
                     "value": rays / (ams * 1e-3) / 1e6}]
        # The N-GPU frame must be byte-identical to the 1-GPU frame (and the two assembly methods must agree).
        # The fused completion signal is what orders the copy below: rank 0 clears the frame, everybody renders once,
        # and rank 0 copies the frame out ON ITS STREAM right behind the wait kernel, with no host-side barrier between.
        if rank == 0:
            rr.ffi.check(lib.rr_device_memset(local_rank, frame_ptr, 0, frame_bytes))
        barrier()
        step_main()
        peer = None
        if rank == 0:
            peer = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)
            whole = ren.frame_params()  # band_count = 1: the unpack kernel degenerates to a row-wise copy
            rr.ffi.check(lib.rr_bands_unpack_device(C.byref(whole), frame_ptr, frame_bytes, C.c_void_p(peer.data_ptr()), sptr))
        step_gather()
        barrier()
        if rank == 0:
            single = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)
            scene.render_rgb8_device(ren.frame_params(), single.data_ptr(), stream=stream.cuda_stream)
            torch.cuda.synchronize(dev)
            st = (C.c_uint32 * 1)()
            rr.ffi.check(lib.rr_device_read(local_rank, status_ptr, st, 4))
            ok = bool(torch.equal(peer, single)) and bool(torch.equal(gframe, single)) and st[0] == 0
            frame_check = "identical to the 1-GPU frame" if ok else ("MISMATCH" if st[0] == 0 else "FENCE TIMEOUT")
        barrier()

    # ---- e2e: the reference-facing call, frame delivered to page-locked HOST memory ------------
    e2e_steps = max(5, min(args.steps, 50))
    if not sharded:
        host = C.c_void_p()
        rr.ffi.check(lib.rr_host_alloc(max(1, frame_bytes), C.byref(host)))

        def e2e_step():
            rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), host, 0))
    else:
        shm_path = f"/dev/shm/rr_frame_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            with open(shm_path, "wb") as f:
                f.truncate(frame_bytes)
        dist.barrier()
        shm_f = open(shm_path, "r+b")
        shm = mmap.mmap(shm_f.fileno(), frame_bytes)
        host_arr = np.frombuffer(shm, dtype=np.uint8)
        host = C.c_void_p(host_arr.ctypes.data)
        rr.ffi.check(lib.rr_host_register(host, frame_bytes))

        def e2e_step():
            rr.ffi.check(lib.rr_render_rgb8_placed(scene.handle, C.byref(p), host, 0))
            dist.barrier()

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_s.item()) * 1e3 / e2e_steps
    e2e_value = rays / (e2e_ms * 1e-3) / 1e6
    # what bounds e2e: the frame has to cross PCIe once. Raw pinned D2H bandwidth of the same byte count, same box.
    pcie = None
    if not sharded:
        pin = torch.empty(frame_bytes, dtype=torch.uint8).pin_memory()
        src = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)
        for _ in range(3):
            pin.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(10):
            pin.copy_(src, non_blocking=True)
        c1.record(stream)
        torch.cuda.synchronize(dev)
        raw_ms = c0.elapsed_time(c1) / 10
        pcie = {"bound": "pcie-d2h", "bytes": frame_bytes, "raw_copy_ms": raw_ms, "raw_copy_gbs": frame_bytes / raw_ms / 1e6,
                "achieved_gbs": frame_bytes / e2e_ms / 1e6, "frac": raw_ms / e2e_ms}
        del pin, src
    e2e_check = None
    if sharded:
        if rank == 0:
            single = np.empty((H, W, 3), dtype=np.uint8)
            scene.render_rgb8(ren.frame_params(), out=single)
            e2e_check = "identical to the 1-GPU frame" if np.array_equal(host_arr.reshape(H, W, 3), single) else "MISMATCH"
        barrier()
        lib.rr_host_unregister(host)
        del host_arr
        shm.close()
        shm_f.close()
        dist.barrier()
        if rank == 0:
            os.unlink(shm_path)
    else:
        lib.rr_host_free(host)

    if rank != 0:
        if sharded:
            lib.rr_ipc_close(local_rank, frame_ptr)
        scene.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (the render kernel) -----------------------------------
    a, b = C.c_float(), C.c_float()
    rr.ffi.check(lib.rr_fp32_peak_tflops(local_rank, C.byref(a), C.byref(b)))
    peaks = measured_peaks()
    sm_max = (peaks or {}).get("sm_max_mhz", 1965.0)
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    derived_unfused = 128 * n_sm * sm_max * 1e6 / 1e12  # 1 flop/lane/clk: FMUL and FADD issue separately (-fmad=false)
    share = (my_rows / H) if sharded else 1.0  # one launch renders this rank's share of the frame
    achieved = flops * share / (kernel_ms * 1e-3) / 1e12
    fb_bytes = my_rows * W * 3
    hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
    roofline = {
        "bound": "fp32", "achieved": achieved, "peak": derived_unfused, "unit": "TFLOP/s", "frac": achieved / derived_unfused,
        "traffic": None,
        "kernel": "rr::march_kernel" if march else "rr::trace_kernel", "kernel_ms": kernel_ms,
        "algorithmic_flops_per_launch": int(flops * share),
        "peak_source": f"derived: 128 FP32 lanes x {n_sm} SMs x {sm_max:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz"
                       f"{'' if peaks else ' absent: fallback 1965'}) x 1 flop (unfused FMUL/FADD, -fmad=false for bit parity); "
                       "neither MEASURED_PEAKS.json nor the profiling guide carries an FP32 non-tensor figure",
        "measured_unfused_tflops": a.value, "measured_ffma_tflops": b.value,
        "frac_of_measured_unfused": achieved / a.value if a.value else None,
        "hbm": {"algorithmic_bytes_per_launch": fb_bytes, "achieved_gbs": fb_bytes / (kernel_ms * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "frac": fb_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"},
    }
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name, {})
    except Exception:
        pass
    roofline["traffic"] = prof.get("dram_bytes_per_launch")
    roofline["traffic_note"] = prof.get("note")
    culled = march or ren.flatten().desc.n_objects >= 25
    if culled:
        # The exact culls (BVH; in march mode the sqrt / bounding-sphere / glow skips) remove most of the reference's
        # brute-force arithmetic, so reference-equivalent flops / time says nothing about pipe utilisation and must not be
        # read as a roofline fraction. frac is computed from the flops the kernel EXECUTES (ncu count of the same launch,
        # committed under profiles/), the reference-equivalent rate is kept beside it.
        roofline["frac_reference_equivalent"] = roofline["frac"]
        roofline["achieved_reference_equivalent"] = roofline["achieved"]
        ex = prof.get("executed_flops_per_launch") if not sharded and (W, H) == (3840, 2160) else None
        roofline["achieved"] = ex / (kernel_ms * 1e-3) / 1e12 if ex else None
        roofline["frac"] = roofline["achieved"] / derived_unfused if ex else None
        roofline["frac_of_measured_unfused"] = roofline["achieved"] / a.value if ex and a.value else None
        roofline["executed_flops_per_launch"] = ex
        roofline["note"] = ("culled launch: achieved/frac = flops EXECUTED per launch (ncu, " + str(prof.get("executed_note")) +
                            ") / live kernel time; *_reference_equivalent = the reference's brute-force flop count / the same time")
    elif prof.get("executed_flops_per_launch") and not sharded and (W, H) == (3840, 2160):
        roofline["executed_flops_per_launch"] = prof["executed_flops_per_launch"]
        roofline["frac_executed"] = prof["executed_flops_per_launch"] / (kernel_ms * 1e-3) / 1e12 / derived_unfused

    # ---- CPU baseline beside it (oracle port on all host cores; bounded sample) -----------------
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import binding as ob

        threads = os.cpu_count() or 1
        reps = 3 if not march else 1
        if march or WORKLOADS[name][4] == "synthetic":
            sp = ren.frame_params(1, 5, 16)  # bounded sample: one interleaved 1/16 of the rows
            sub = ob.render(ren, params=sp, threads=threads, want_u8=False, want_counts=True)["counts"].rays()
            best = cpu_arm(ob, ren, threads, reps, params=sp)
            cpu = {"value": sub / best / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": f"rows 5::16 of the {W}x{H} frame ({sub} rays), best of {reps}"}
        else:
            best = cpu_arm(ob, ren, threads, reps)
            cpu = {"value": rays / best / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": f"full {W}x{H} frame ({rays} rays), best of {reps}", "frame_ms": best * 1e3}

    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "width": W, "height": H, "mode": "raymarch" if march else "raytrace",
                   "max_reflections": 3, "max_refractions": 10, "rays_per_frame": rays, "ray_classes": counts,
                   "l2": "flushed between timed steps (256 MiB write, untimed)",
                   "parallelism": f"row-bands{world}x{band_rows}, kernel stores rows AND its completion word into rank 0's memory over NVLink (CUDA IPC), rank 0 waits on the words; no collective"
                   if sharded else "1gpu",
                   "scene_resident": True},
        "frame_ms": ms_per_step,
        "kernel_ms": kernel_ms,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_frame": e2e_ms, "h2d_bytes_per_step": C.sizeof(rr.ffi.rr_frame_params),
                "d2h_bytes_per_step": my_rows * W * 3,
                "api": ("rr_render_rgb8_placed (C ABI): each rank's bands -> one shared page-locked host frame, + barrier"
                        if sharded else "rr_render_rgb8 (C ABI) -> pinned host RGB8 frame"),
                "check": e2e_check, "roofline": pcie},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "wall_s_timed_region": wall_s,
    }
    if alt:
        line["alt"] = alt
        line["frame_check"] = frame_check
    print(json.dumps(line), flush=True)
    if sharded:
        lib.rr_device_free(local_rank, frame_ptr)
    scene.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
