#!/usr/bin/env python
"""bench.py — headline benchmark of the per-pixel tracing path (BASELINE.json metric: Mrays/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-configs]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" renders one whole frame of the workload. The headline workload of the JSON line is
  N=1  "default-4k-trace"       = BASELINE.json configs[1]: built-in scene, 3840x2160, ray-trace mode;
  N>1  "default-8k-trace-bands" = configs[4]: built-in scene, 7680x4320, interleaved 16-row bands over the N ranks
       (one process per GPU). Every rank's render kernel stores its rows AND its completion word straight into rank
       0's frame over NVLink peer memory (CUDA IPC); rank 0's stream waits on the words. No collective in the step;
       the NCCL-gather + un-interleave and the all-reduce-fence alternatives are timed too and reported under "alt".
The same line carries every other BASELINE config under "configs" (same measurement, fewer steps): at N=1 the 8K
frame, the 4K ray-march + glow frame, the 1 024-sphere frame (through the YAML -s/-d round trip) and the 640x480
frame; at N>1 the band-sharded march and 1 024-sphere frames. Every N>1 entry also carries `same_workload_1gpu_ms`
(rank 0 renders the SAME frame alone, same run, same timing method) and `efficiency_same_workload`.

`value` = reference-equivalent rays (one ray = one scene-level raycast()/raymarch_single(); counted by
the instrumented kernel and checked against the oracle in tests/) per second of device time with the scene resident
in HBM; `e2e` = the same through the C ABI into HOST memory (wall clock, includes the D2H of the frame; for N>1 each
rank copies its bands into one shared page-locked host frame over its own PCIe link), checked against the
device-resident frame. The reference arm (--impl reference) times the CPU oracle (oracle/, the C++ restatement of the
reference: no Rust toolchain exists here) on all host cores.
"""
import argparse
import ctypes as C
import json
import mmap
import os
import statistics
import subprocess
import sys
import tempfile
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
# the other BASELINE configs measured beside the headline workload (name, max timed steps)
EXTRA_1GPU = [("default-8k-trace", 50), ("default-4k-march-glow", 20), ("synthetic1024-4k-trace", 30), ("default-640x480-trace", 50)]
EXTRA_NGPU = [("default-4k-march-glow", 15), ("synthetic1024-4k-trace", 20)]
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
    """The workload's RenderEnv. The synthetic scene goes through the YAML -s / -d round trip, as BASELINE configs[3] says."""
    w, h, march, glow, kind = WORKLOADS[name]
    if kind == "synthetic":
        ren = rr.synthetic_scene(w, h, use_raymarching=march, glow_effect=glow)
        back = rr.synthetic_scene(w, h, n_spheres=0, use_raymarching=march, glow_effect=glow)
        back.deserialize(ren.serialize())
        return back
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


def cpu_baseline(rr, ren, name, rays):
    """Bounded sample of the same workload on all host cores (oracle port)."""
    from oracle import binding as ob

    W, H, march, _, kind = WORKLOADS[name]
    threads = os.cpu_count() or 1
    reps = 3 if not march else 1
    if march or kind == "synthetic":
        sp = ren.frame_params(1, 5, 16)  # bounded sample: one interleaved 1/16 of the rows
        sub = ob.render(ren, params=sp, threads=threads, want_u8=False, want_counts=True)["counts"].rays()
        best = cpu_arm(ob, ren, threads, reps, params=sp)
        return {"value": sub / best / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                "sample": f"rows 5::16 of the {W}x{H} frame ({sub} rays), best of {reps}", "est_frame_ms": best * 1e3 * rays / sub}
    best = cpu_arm(ob, ren, threads, reps)
    return {"value": rays / best / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"full {W}x{H} frame ({rays} rays), best of {reps}", "frame_ms": best * 1e3}


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


class Ctx:
    pass


def setup(args):
    import numpy as np
    import torch

    import ray_rust_b200 as rr
    from ray_rust_b200 import bands

    c = Ctx()
    c.np, c.torch, c.rr, c.bands, c.args = np, torch, rr, bands, args
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != c.world and c.world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={c.world}")
    c.dist = None
    if c.world > 1:
        # NCCL_DEBUG=VERSION prints "NCCL version ..." on stdout, which must carry exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist

        torch.cuda.set_device(c.local_rank)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)  # communicator creation prints on stdout; keep stdout for the one JSON line
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", c.local_rank))
            warm = torch.zeros(1, device=torch.device("cuda", c.local_rank))
            dist.all_reduce(warm)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        c.dist = dist
    c.dev = torch.device("cuda", c.local_rank)
    torch.cuda.set_device(c.dev)
    c.lib = rr.ffi.load()
    # Everything timed runs on ONE explicit, non-default CUDA stream: the L2 flush, the events and the render launches. (The
    # default stream's handle is 0, which the C ABI reads as "no caller stream: use the handle's own stream and block" — the
    # launch would then run beside the flush on another stream and the events would bracket a host-side wait. The records
    # up to profiles/r2u_* were taken that way and read ~8 us too long per step; profiles/r2z_flush_gap.txt.)
    c.stream = torch.cuda.Stream(device=c.dev)
    torch.cuda.set_stream(c.stream)
    if not c.stream.cuda_stream:
        raise SystemExit("bench.py needs a non-default CUDA stream")
    c.sptr = C.c_void_p(c.stream.cuda_stream)
    c.flush = torch.empty(256 << 20, dtype=torch.uint8, device=c.dev)  # > 126 MB L2
    c.sharded = c.world > 1
    c.band_rows = args.band_rows or BAND_ROWS
    a, b = C.c_float(), C.c_float()
    rr.ffi.check(c.lib.rr_fp32_peak_tflops(c.local_rank, C.byref(a), C.byref(b)))
    c.meas_unfused, c.meas_ffma = a.value, b.value
    c.peaks = measured_peaks()
    return c


def barrier(c):
    if c.dist is not None:
        c.dist.barrier()
    c.torch.cuda.synchronize(c.dev)


def timed(c, step_fn, steps, warmup, sample_clocks, pre_step=None):
    """W untimed + K timed steps; every timed step starts behind an (untimed) L2 flush and, for N>1, a barrier + a
    device-side barrier; CUDA events on the launching stream; max over ranks."""
    torch = c.torch
    for _ in range(warmup):
        step_fn()
    barrier(c)
    sampler = ClockSampler(c.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches = 0
    barrier(c)
    t0 = time.perf_counter()
    for i in range(steps):
        c.flush.fill_(i & 0xFF)  # evict the previous frame from L2 (not timed)
        if pre_step:
            pre_step()
        ev[i][0].record(c.stream)
        launches += step_fn() or 0
        ev[i][1].record(c.stream)
    barrier(c)
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=c.dev)
    if c.dist is not None:
        c.dist.all_reduce(tot, op=c.dist.ReduceOp.MAX)
    return float(tot.item()) / steps, launches, clocks, wall


def measure(c, name, steps, warmup, primary):
    """One workload on the current topology (1 GPU, or row bands over the N ranks). Returns the result dict (rank 0)."""
    np, torch, rr, lib, dist = c.np, c.torch, c.rr, c.lib, c.dist
    rank, world, local_rank, dev, stream, sptr = c.rank, c.world, c.local_rank, c.dev, c.stream, c.sptr
    W, H, march, glow, kind = WORKLOADS[name]
    ren = make_env(rr, name)
    scene = rr.DeviceScene(ren, local_rank)
    sharded = c.sharded
    band_rows = c.band_rows
    p_eq = ren.frame_params(band_rows, rank, world) if sharded else ren.frame_params()  # equal shares (e2e, gather alternative)
    p = p_eq                                                                            # the device-frame step's shard (may get unequal spans below)
    whole = ren.frame_params()
    eq_rows = rr.frame_rows(p_eq)
    max_rows = c.bands.max_shard_rows(H, band_rows, world) if sharded else H
    shard_bytes = max_rows * W * 3
    frame_bytes = H * W * 3

    # ---- ray counts of the whole frame (reference-equivalent; instrumented kernel, untimed) ----
    _, cnt = scene.render_count(whole, want_image=False)
    counts = cnt.as_dict()
    rays = cnt.rays()
    flops = algorithmic_flops(counts, march)

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
            dist.barrier()
            gate[0] += 1
            rr.ffi.check(lib.rr_fence_signal_device(local_rank, C.c_void_p(arrive_ptr.value + 4 * rank), gate[0], sptr))
            if rank == 0:
                rr.ffi.check(lib.rr_fence_wait_device(local_rank, arrive_ptr, world, gate[0], 5000, status_ptr, sptr))
                rr.ffi.check(lib.rr_fence_signal_device(local_rank, go_ptr, gate[0], sptr))
            rr.ffi.check(lib.rr_fence_wait_device(local_rank, go_ptr, 1, gate[0], 5000, status_ptr, sptr))
        tok = torch.zeros(1, dtype=torch.float32, device=dev)
        packed = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)  # a shard in packed order (any share of the frame)
        gathered = torch.empty(world * shard_bytes, dtype=torch.uint8, device=dev) if rank == 0 and primary else None
        gframe = torch.empty(frame_bytes, dtype=torch.uint8, device=dev) if rank == 0 and primary else None
    else:
        device_barrier = None
        out = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)

    def step_main():
        """one frame, device resident. returns number of this repo's kernels launched"""
        if not sharded:
            scene.render_rgb8_device(p, out.data_ptr(), stream=stream.cuda_stream)
            return 1
        # fused: the render kernel stores its rows into rank 0's frame over NVLink AND publishes its completion word
        # there; rank 0's stream waits on the words. No collective in the step.
        epoch[0] += 1
        # (the call publishes into d_flags[band_index]; with unequal spans band_index is a slot, not the rank)
        rr.ffi.check(lib.rr_render_rgb8_placed_signal_device(scene.handle, C.byref(shard[0]), frame_ptr, W * 3,
                                                             C.c_void_p(flags_ptr.value + 4 * (rank - shard[0].band_index)), epoch[0], sptr))
        if rank == 0:
            rr.ffi.check(lib.rr_fence_wait_device(local_rank, flags_ptr, world, epoch[0], 5000, status_ptr, sptr))
            return 2
        return 1

    def step_allreduce_fence():
        rr.ffi.check(lib.rr_render_rgb8_placed_device(scene.handle, C.byref(p_eq), frame_ptr, W * 3, sptr))
        dist.all_reduce(tok)  # completion fence: rank 0's stream passes it only after every rank's kernel
        return 1

    def step_gather():
        scene.render_rgb8_device(p_eq, packed.data_ptr(), stream=stream.cuda_stream)
        glist = list(gathered.chunk(world)) if rank == 0 else None
        dist.gather(packed[:shard_bytes], glist, dst=0)
        if rank == 0:
            rr.ffi.check(lib.rr_bands_unpack_device(C.byref(p_eq), C.c_void_p(gathered.data_ptr()), shard_bytes,
                                                    C.c_void_p(gframe.data_ptr()), sptr))
            return 2
        return 1

    shard = [p]
    shares = None
    single = None
    same_1gpu_ms = None
    if sharded:
        # the SAME frame on one GPU in the same run (honest scaling base; also what the share chooser needs)
        single = torch.empty(frame_bytes, dtype=torch.uint8, device=dev) if rank == 0 else None
        if rank == 0:
            sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(5, min(steps, 20)))]
            for _ in range(3):
                scene.render_rgb8_device(whole, single.data_ptr(), stream=stream.cuda_stream)
            for a, b in sev:
                c.flush.fill_(2)
                a.record(stream)
                scene.render_rgb8_device(whole, single.data_ptr(), stream=stream.cuda_stream)
                b.record(stream)
            torch.cuda.synchronize(dev)
            same_1gpu_ms = sum(a.elapsed_time(b) for a, b in sev) / len(sev)
        barrier(c)

    def copy_rate(params):
        """Raw NVLink ingress of rank 0: every other rank copies as many bytes as its shard holds from local memory into
        rank 0's frame at the same time (plain cudaMemcpyAsync on peer memory), behind the same device-side barrier,
        timed like the step (events, max over ranks). Returns (ms, bytes into rank 0)."""
        n_bytes = rr.frame_rows(params) * W * 3
        off = (rank * (frame_bytes // world)) // 256 * 256  # anywhere inside the frame (the bytes are overwritten by the next render)
        off = min(off, frame_bytes - n_bytes)

        def step_copy():
            if rank != 0:  # rank 0 only receives
                rr.ffi.check(lib.rr_device_copy(local_rank, C.c_void_p(frame_ptr.value + off), C.c_void_p(packed.data_ptr()), n_bytes, sptr))
            return 0

        cms, _, _, _ = timed(c, step_copy, 10, 3, False, device_barrier)
        tot = torch.tensor([n_bytes if rank != 0 else 0], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        return cms, int(tot.item())

    if sharded and not march and not c.args.equal_shares:
        # Unequal shares. Every row rendered elsewhere enters rank 0 through ITS NVLink ports, rows it renders itself do
        # not; when that ingress, not the kernels, bounds the step (8 GPUs, sub-millisecond frames), rank 0 should own more
        # band slots than the others (rr_frame_params.band_span). Chosen here from two numbers measured in this run: the
        # 1-GPU time of the frame and the raw ingress rate with equal shards.
        cms, inbound = copy_rate(p_eq)
        pick = [1, 1, None]
        if rank == 0:
            rate = inbound / cms  # bytes per ms
            best = None
            for b in (1, 2, 3, 4):
                for a in range(b, 3 * b + 1):
                    period = a + (world - 1) * b
                    # (+ 0.030 ms: what a shard launch and the transfer path take beyond their proportional part, fitted to
                    #  the 8-GPU records profiles/r2t_bench_n8*.json: equal shares 0.148 ms, 4/3 0.1425, 5/4 0.1418)
                    t0 = same_1gpu_ms * a / period + 0.030        # rank 0's kernel
                    tr = same_1gpu_ms * b / period + 0.030
                    tx = frame_bytes * (period - a) / period / rate + 0.030  # what has to cross into rank 0
                    est = max(t0, tr, tx)
                    if best is None or est < best[0] - 1e-4:
                        best = (est, a, b)
            pick = [best[1], best[2], best[0]]
            if c.args.shares:
                pick = [int(c.args.shares.split(",")[0]), int(c.args.shares.split(",")[1]), None]
        dist.broadcast_object_list(pick, src=0)
        if pick[0] != pick[1]:
            spans, period = c.bands.weighted_spans(world, pick[0], pick[1])
            shard[0] = ren.frame_params(band_rows, spans[rank][0], period, spans[rank][1])
        shares = {"rank0_slots": pick[0], "other_slots": pick[1], "period_slots": pick[0] + (world - 1) * pick[1],
                  "rank0_share": pick[0] / (pick[0] + (world - 1) * pick[1]), "predicted_step_ms": pick[2],
                  "inputs": {"same_workload_1gpu_ms": same_1gpu_ms, "raw_ingress_ms_equal_shards": cms},
                  "note": "rr_frame_params.band_span: rank 0 (the frame's owner) renders more band slots per period than the other ranks"}
    p = shard[0]
    my_rows = rr.frame_rows(p)

    ms_per_step, launches, clocks, wall_s = timed(c, step_main, steps, warmup, primary, device_barrier)
    value = rays / (ms_per_step * 1e-3) / 1e6

    # kernel-only time of this rank's render launch (roofline numerator), same stream, same L2 hygiene
    ksteps = min(steps, 30)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ksteps)]
    kbuf = packed if sharded else out
    for a, b in kev:
        c.flush.fill_(1)
        a.record(stream)
        scene.render_rgb8_device(p, kbuf.data_ptr(), stream=stream.cuda_stream)
        b.record(stream)
    torch.cuda.synchronize(dev)
    per_rank_kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / len(kev)
    kt = torch.tensor([per_rank_kernel_ms], dtype=torch.float64, device=dev)
    kall = None
    if dist is not None:
        kall = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(kall, kt)
        kall = [float(x.item()) for x in kall]
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    kernel_ms = float(kt.item())

    # ---- N>1: the SAME frame on one GPU in the same run (honest scaling base), alternatives, frame check ----
    alt = None
    frame_check = None
    if sharded:
        if primary:
            gms, _, _, _ = timed(c, step_gather, max(5, steps // 2), warmup, False, device_barrier)
            ams, _, _, _ = timed(c, step_allreduce_fence, max(5, steps // 2), warmup, False, device_barrier)
            alt = [{"method": "nccl gather to rank 0 + rr_bands_unpack_device", "ms_per_step": gms, "value": rays / (gms * 1e-3) / 1e6},
                   {"method": "placed NVLink stores + 1-element NCCL all-reduce as the completion fence", "ms_per_step": ams,
                    "value": rays / (ams * 1e-3) / 1e6}]
        # The N-GPU frame must be byte-identical to the 1-GPU frame (and the assembly methods must agree). The fused
        # completion signal is what orders the copy below: rank 0 clears the frame, everybody renders once, and rank 0
        # copies the frame out ON ITS STREAM right behind the wait kernel, with no host-side barrier between.
        if rank == 0:
            rr.ffi.check(lib.rr_device_memset(local_rank, frame_ptr, 0, frame_bytes))
        barrier(c)
        step_main()
        peer = None
        if rank == 0:
            peer = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)
            rr.ffi.check(lib.rr_bands_unpack_device(C.byref(whole), frame_ptr, frame_bytes, C.c_void_p(peer.data_ptr()), sptr))
        if primary:
            step_gather()
        barrier(c)
        if rank == 0:
            torch.cuda.synchronize(dev)
            st = (C.c_uint32 * 1)()
            rr.ffi.check(lib.rr_device_read(local_rank, status_ptr, st, 4))
            ok = bool(torch.equal(peer, single)) and (not primary or bool(torch.equal(gframe, single))) and st[0] == 0
            frame_check = "identical to the 1-GPU frame" if ok else ("MISMATCH" if st[0] == 0 else "FENCE TIMEOUT")
        barrier(c)

    # ---- N>1: what bounds the device-frame step once the kernels are short: every row rendered elsewhere has to enter
    # rank 0 over ITS NVLink ports. Raw rate of that, measured here: every other rank copies as many bytes as its shard
    # holds from local memory into rank 0's frame at the same time (plain cudaMemcpyAsync on peer memory), behind the same
    # device-side barrier, timed like the step (events, max over ranks).
    nvlink = None
    if sharded and primary:
        cms, inbound = copy_rate(p)
        nvlink = {"bound": f"nvlink ingress of rank 0 ({world - 1} peers writing their shards at once, cudaMemcpyAsync)",
                  "bytes_into_rank0": inbound, "raw_copy_ms": cms, "raw_gbs": inbound / cms / 1e6,
                  "achieved_gbs": inbound / ms_per_step / 1e6, "frac": cms / ms_per_step,
                  "note": "the step cannot be shorter than max(raw_copy_ms, this frame's 1-GPU time x the largest share): the kernel's own stores ARE the transfer"}
        barrier(c)

    # ---- e2e: the reference-facing call, frame delivered to page-locked HOST memory ------------
    e2e_steps = max(5, min(steps, 50))
    if not sharded:
        host = C.c_void_p()
        rr.ffi.check(lib.rr_host_alloc(max(1, frame_bytes), C.byref(host)))

        def e2e_step():
            rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), host, 0))
    else:
        shm_path = f"/dev/shm/rr_frame_{os.environ.get('MASTER_PORT', '0')}_{name}"
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
            rr.ffi.check(lib.rr_render_rgb8_placed(scene.handle, C.byref(p_eq), host, 0))
            dist.barrier()

    for _ in range(3):
        e2e_step()
    barrier(c)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_s.item()) * 1e3 / e2e_steps
    e2e_value = rays / (e2e_ms * 1e-3) / 1e6
    # what bounds e2e: the frame has to cross PCIe once. Raw pinned D2H copy of the same bytes on the same box, measured
    # here: one GPU alone at N=1; at N>1 all ranks copy their share at the same time (aggregate rate of the box).
    copy_bytes = frame_bytes if not sharded else eq_rows * W * 3
    pin = torch.empty(max(1, copy_bytes), dtype=torch.uint8).pin_memory()
    src = torch.empty(max(1, copy_bytes), dtype=torch.uint8, device=dev)
    for _ in range(3):
        pin.copy_(src, non_blocking=True)
    barrier(c)
    t0 = time.perf_counter()
    for _ in range(10):
        pin.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    raw_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(raw_t, op=dist.ReduceOp.MAX)
    raw_ms = float(raw_t.item()) * 1e3 / 10
    pcie = {"bound": "pcie-d2h" if not sharded else f"pcie-d2h, {world} GPUs copying their bands at once (aggregate)",
            "bytes": frame_bytes, "raw_copy_ms": raw_ms, "raw_copy_gbs": frame_bytes / raw_ms / 1e6,
            "achieved_gbs": frame_bytes / e2e_ms / 1e6, "frac": raw_ms / e2e_ms}
    del pin, src

    # the frame the timed e2e calls delivered must be the device-resident frame
    e2e_check = None
    e2e_pageable = None
    if sharded:
        if rank == 0:
            ref = single.cpu().numpy()
            e2e_check = "identical to the 1-GPU frame" if np.array_equal(host_arr, ref) else "MISMATCH"
        barrier(c)
        lib.rr_host_unregister(host)
        del host_arr
        shm.close()
        shm_f.close()
        dist.barrier()
        if rank == 0:
            os.unlink(shm_path)
    else:
        got = np.frombuffer(C.string_at(host, frame_bytes), dtype=np.uint8)
        torch.cuda.synchronize(dev)
        e2e_check = "identical to the device-resident frame" if np.array_equal(got, out.cpu().numpy()) else "MISMATCH"
        lib.rr_host_free(host)
        # the same call with a plain (pageable) caller buffer: what a caller gets who ignores rr_host_alloc
        pg = np.empty(frame_bytes, dtype=np.uint8)
        pg[:] = 0  # touch the pages
        for _ in range(2):
            rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), pg.ctypes.data_as(C.c_void_p), 0))
        psteps = max(3, min(e2e_steps, 10))
        t0 = time.perf_counter()
        for _ in range(psteps):
            rr.ffi.check(lib.rr_render_rgb8(scene.handle, C.byref(p), pg.ctypes.data_as(C.c_void_p), 0))
        pms = (time.perf_counter() - t0) * 1e3 / psteps
        e2e_pageable = {"ms_per_frame": pms, "value": rays / (pms * 1e-3) / 1e6, "vs_pinned": pms / e2e_ms,
                        "check": "identical" if np.array_equal(pg, got) else "MISMATCH",
                        "note": "pageable destination: the CUDA driver stages the D2H copy through its own pinned buffers "
                                "(synchronous, ~half the pinned rate); every caller shipped here (CLI, web, render_frames, "
                                "render()) uses rr_host_alloc frames instead"}

    if rank != 0:
        if sharded:
            lib.rr_ipc_close(local_rank, frame_ptr)
        scene.close()
        return None

    # ---- roofline of the dominant kernel (the render kernel) -----------------------------------
    peaks = c.peaks
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
        "measured_unfused_tflops": c.meas_unfused, "measured_ffma_tflops": c.meas_ffma,
        "frac_of_measured_unfused": achieved / c.meas_unfused if c.meas_unfused else None,
        "hbm": {"algorithmic_bytes_per_launch": fb_bytes, "achieved_gbs": fb_bytes / (kernel_ms * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "frac": fb_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"},
    }
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name, {})
    except Exception:
        pass
    roofline["traffic"] = prof.get("dram_bytes_per_launch") if not sharded else None
    roofline["traffic_note"] = prof.get("note") if not sharded else None
    culled = march or ren.flatten().desc.n_objects >= 25
    ex = prof.get("executed_flops_per_launch") if not sharded else None
    if culled:
        # The exact culls (BVH; in march mode the sqrt / bounding-sphere / glow skips) remove most of the reference's
        # brute-force arithmetic, so reference-equivalent flops / time says nothing about pipe utilisation and must not be
        # read as a roofline fraction. frac is computed from the flops the kernel EXECUTES (ncu count of the same launch,
        # committed under profiles/), the reference-equivalent rate is kept beside it.
        roofline["frac_reference_equivalent"] = roofline["frac"]
        roofline["achieved_reference_equivalent"] = roofline["achieved"]
        roofline["achieved"] = ex / (kernel_ms * 1e-3) / 1e12 if ex else None
        roofline["frac"] = roofline["achieved"] / derived_unfused if ex else None
        roofline["frac_of_measured_unfused"] = roofline["achieved"] / c.meas_unfused if ex and c.meas_unfused else None
        roofline["executed_flops_per_launch"] = ex
        roofline["note"] = ("culled launch: achieved/frac = flops EXECUTED per launch (ncu, " + str(prof.get("executed_note")) +
                            ") / live kernel time; *_reference_equivalent = the reference's brute-force flop count / the same time")
    elif ex:
        roofline["executed_flops_per_launch"] = ex
        roofline["frac_executed"] = ex / (kernel_ms * 1e-3) / 1e12 / derived_unfused

    res = {
        "workload": name if not sharded or name.endswith("-bands") else name + "-bands",
        "width": W, "height": H, "mode": "raymarch" if march else "raytrace", "rays_per_frame": rays,
        "steps": steps, "ms_per_step": ms_per_step, "value": value, "unit": "Mrays/s", "kernel_ms": kernel_ms,
        "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_frame": e2e_ms, "h2d_bytes_per_step": C.sizeof(rr.ffi.rr_frame_params),
                "d2h_bytes_per_step": (eq_rows if sharded else my_rows) * W * 3,
                "api": ("rr_render_rgb8_placed (C ABI): each rank's bands -> one shared page-locked host frame, + barrier"
                        if sharded else "rr_render_rgb8 (C ABI) -> pinned host RGB8 frame"),
                "check": e2e_check, "roofline": pcie},
        "roofline": roofline,
        "_counts": counts, "_clocks": clocks, "_wall_s": wall_s,
    }
    if e2e_pageable:
        res["e2e_pageable"] = e2e_pageable
    if sharded:
        res["same_workload_1gpu_ms"] = same_1gpu_ms
        res["efficiency_same_workload"] = same_1gpu_ms / (world * ms_per_step)
        res["per_rank_kernel_ms"] = kall
        res["ideal_kernel_ms"] = same_1gpu_ms / world
        res["frame_check"] = frame_check
        if shares:
            res["band_shares"] = shares
        if nvlink:
            res["nvlink_roofline"] = nvlink
            res["step_lower_bound_ms"] = max(nvlink["raw_copy_ms"], same_1gpu_ms * (shares["rank0_share"] if shares else 1.0 / world))
            res["efficiency_vs_bound"] = res["step_lower_bound_ms"] / ms_per_step
        if alt:
            res["alt"] = alt
        lib.rr_device_free(local_rank, frame_ptr)
    elif not c.args.no_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline(rr, ren, name, rays)
    scene.close()
    return res


def host_path(c):
    """What the reference's own "Rendering time" consists of once the render takes < 1 ms (main.rs:316-348 times render() +
    image::save_buffer): the PNG encode alone, the whole CLI, and render_frames (render.rs:926-989) in frames/s."""
    np, rr = c.np, c.rr
    out = {}
    hostlib = os.path.join(ROOT, "ray-rust_b200", "host", "libray_rust_host.so")
    cli = os.path.join(ROOT, "ray-rust_b200", "host", "ray-rust")
    if not (os.path.exists(hostlib) and os.path.exists(cli)):
        return {"unavailable": "C++ host layer not built"}
    hl = C.CDLL(hostlib)
    hl.rrh_png_encode_ms.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    hl.rrh_env_new.restype = C.c_void_p
    hl.rrh_env_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64]
    hl.rrh_env_free.argtypes = [C.c_void_p]
    hl.rrh_env_deserialize.argtypes = [C.c_void_p, C.c_char_p]
    hl.rrh_render_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_int]
    hl.rrh_last_error.restype = C.c_char_p
    cores = os.cpu_count() or 1
    for tag, (w, h) in (("4k", (3840, 2160)), ("8k", (7680, 4320))):
        ren = rr.default_scene(w, h)
        scene = rr.DeviceScene(ren, c.local_rank)
        img = scene.render_rgb8(ren.frame_params())
        scene.close()
        ms, nb = C.c_double(), C.c_uint64()
        if hl.rrh_png_encode_ms(img.ctypes.data_as(C.c_void_p), w, h, 3, C.byref(ms), C.byref(nb)) == 0:
            out[f"png_encode_ms_{tag}"] = ms.value
            out[f"png_bytes_{tag}"] = int(nb.value)
    out["png_encode_note"] = (f"encode_png_rgb8 (host/rr_png.cpp): filter Sub + zlib level 1 over parallel row stripes, {cores} host cores, "
                              "best of 3; image::save_buffer in the reference")
    # whole CLI, as a user runs it: process start, CUDA context, scene upload, render, PNG encode, file write
    with tempfile.TemporaryDirectory() as td:
        for tag, (w, h) in (("4k", (3840, 2160)),):
            best, inner = None, None
            for _ in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([cli, str(w), str(h), "-o", os.path.join(td, "x.png"), "--gpu", str(c.local_rank)], capture_output=True, text=True)
                dt = (time.perf_counter() - t0) * 1e3
                if r.returncode == 0 and (best is None or dt < best):
                    best = dt
                    for ln in r.stdout.splitlines():
                        if ln.startswith("Rendering time:"):
                            inner = float(ln.split(":")[1]) * 1e3
            out[f"cli_wall_ms_{tag}"] = best
            out[f"cli_rendering_time_ms_{tag}"] = inner
    out["cli_note"] = ("`ray-rust 3840 2160 -o x.png`: wall = whole process incl. CUDA context creation; rendering_time = what the CLI "
                       "prints, i.e. the reference's timed region main.rs:316-348 (scene upload + render + D2H + PNG encode + write)")
    # render_frames: 24 frames of a camera move, 1080p and 4K, frames only (CRC) and with the PNG encode the CLI does
    import yaml

    for tag, (w, h) in (("1080p", (1920, 1080)), ("4k", (3840, 2160))):
        ren = rr.default_scene(w, h)
        doc = yaml.safe_load(ren.serialize())
        doc["camera_motion"] = [
            {"camera": {"position": {"x": 60.0, "y": -120.0, "z": -280.0}, "pyr": dict(doc["camera"]["pyr"])},
             "velocity": {"x": 10.0, "y": 0.0, "z": 5.0}, "camera_target": {"x": 0.0, "y": -30.0, "z": 172.0}, "duration": 6.0},
            {"camera": {"position": {"x": 120.0, "y": -60.0, "z": -240.0}, "pyr": dict(doc["camera"]["pyr"])},
             "velocity": {"x": 0.0, "y": 0.0, "z": 0.0}, "camera_target": None, "duration": 6.0},
        ]
        env = hl.rrh_env_new(0, w, h, 0, 0, 0.0, 0, 0)
        if not env or hl.rrh_env_deserialize(env, yaml.safe_dump(doc).encode()) != 0:
            continue
        for mode, key in ((0, "frames_per_s"), (1, "frames_per_s_with_png_encode")):
            sec = C.c_double()
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)  # render_frames prints the reference's progress lines; stdout carries exactly one JSON line
            try:
                hl.rrh_render_frames(env, mode, 1, C.byref(sec), None, 0)  # warm-up (scene upload, pinned frames)
                n = hl.rrh_render_frames(env, mode, 1, C.byref(sec), None, 0)
            finally:
                os.dup2(saved, 1)
                os.close(saved)
            if n > 0:
                out[f"render_frames_{tag}_{key}"] = n / sec.value
                out[f"render_frames_{tag}_frames"] = n
            else:
                out[f"render_frames_{tag}_error"] = hl.rrh_last_error().decode("utf-8", "replace")
        hl.rrh_env_free(env)
    out["render_frames_note"] = ("render_frames (render.rs:926-989) on ONE GPU through rr_render_rgb8_async: two page-locked frames in flight, "
                                 "frame_proc = CRC-32 of the frame / in-memory PNG encode (what the CLI does before writing)")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="only the headline workload (no other BASELINE configs, no host-path timings)")
    ap.add_argument("--band-rows", type=int, default=None, help="rows per interleaved band of the multi-GPU shard (default 16)")
    ap.add_argument("--equal-shares", action="store_true", help="N>1: every rank owns one band slot per period (no unequal band spans)")
    ap.add_argument("--shares", default=None, help="N>1: force 'A,B' band slots per period for rank 0 / every other rank (default: chosen from measurements)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps or 10
        args.warmup = 1 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = args.steps or 100
    args.warmup = 5 if args.warmup is None else max(3, args.warmup)

    c = setup(args)
    name = args.workload or ("default-4k-trace" if c.world == 1 else "default-8k-trace-bands")
    main_res = measure(c, name, args.steps, args.warmup, True)
    configs = {}
    if not args.workload and not args.no_configs:
        for nm, cap in (EXTRA_NGPU if c.sharded else EXTRA_1GPU):
            r = measure(c, nm, max(5, min(args.steps, cap)), args.warmup, False)
            if r is not None:
                configs[r["workload"]] = {k: v for k, v in r.items() if not k.startswith("_") or k == "_counts"}
                configs[r["workload"]]["ray_classes"] = configs[r["workload"]].pop("_counts")
    hp = None
    if c.rank == 0 and not c.sharded and not args.workload and not args.no_configs:
        try:
            hp = host_path(c)
        except Exception as e:  # noqa: BLE001  (never lose the bench line to the auxiliary timings)
            hp = {"error": repr(e)}

    if c.rank == 0:
        r = main_res
        W, H = r["width"], r["height"]
        line = {
            "metric": "Mrays/s", "value": r["value"], "unit": "Mrays/s", "n_gpus": c.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if c.sharded else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "width": W, "height": H, "mode": r["mode"],
                       "max_reflections": 3, "max_refractions": 10, "rays_per_frame": r["rays_per_frame"], "ray_classes": r["_counts"],
                       "l2": "flushed between timed steps (256 MiB write, untimed)",
                       "stream": "one non-default CUDA stream carries the flush, the events and the render launches",
                       "parallelism": f"row-bands{c.world}x{c.band_rows}, kernel stores rows AND its completion word into rank 0's memory over NVLink (CUDA IPC), rank 0 waits on the words; no collective"
                       if c.sharded else "1gpu",
                       "scene_resident": True},
            "frame_ms": r["ms_per_step"],
            "kernel_ms": r["kernel_ms"],
            "clocks": r["_clocks"],
            "e2e": r["e2e"],
            "gpu_launches": r["gpu_launches"],
            "roofline": r["roofline"],
            "cpu_baseline": r.get("cpu_baseline"),
            "wall_s_timed_region": r["_wall_s"],
        }
        for k in ("e2e_pageable", "same_workload_1gpu_ms", "efficiency_same_workload", "per_rank_kernel_ms", "ideal_kernel_ms", "frame_check",
                  "band_shares", "nvlink_roofline", "step_lower_bound_ms", "efficiency_vs_bound", "alt"):
            if k in r:
                line[k] = r[k]
        if configs:
            line["configs"] = configs
        if hp:
            line["host_path"] = hp
        print(json.dumps(line), flush=True)
    if c.dist is not None:
        c.dist.barrier()
        c.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
