"""world_size-2 (and 3) CPU tests of the N>1 path with the gloo backend: each rank renders its
interleaved row bands (with the oracle here — there is no GPU in this container), bands are gathered
to rank 0 and un-interleaved; the result must be byte-identical to the single-rank frame."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, band_rows, w, h, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ray_rust_b200 as rr
    from oracle import binding as ob
    from ray_rust_b200 import bands

    ren = rr.default_scene(w, h)
    p = ren.frame_params(band_rows, rank, world)
    mine = ob.render(ren, params=p)["u8"]
    assert mine.shape[0] == len(bands.shard_rows(h, band_rows, rank, world)) == rr.frame_rows(p)
    pad = bands.max_shard_rows(h, band_rows, world)
    buf = torch.zeros((pad, w, 3), dtype=torch.uint8)
    buf[: mine.shape[0]] = torch.from_numpy(mine)
    glist = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, glist, dst=0)
    # max-over-ranks reduction of a per-rank timing, as bench.py does
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    if rank == 0:
        frame = bands.unpack(torch.stack(glist).numpy(), h, band_rows, world)
        np.save(out_path, frame)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,band_rows", [(2, 16), (3, 5)])
def test_gloo_band_gather_matches_single_rank(tmp_path, world, band_rows):
    sys.path.insert(0, ROOT)
    import ray_rust_b200 as rr
    from oracle import binding as ob

    w, h = 96, 70
    out = str(tmp_path / "frame.npy")
    port = 29600 + world * 7 + band_rows
    mp.spawn(_worker, args=(world, port, band_rows, w, h, out), nprocs=world, join=True)
    full = ob.render(rr.default_scene(w, h))["u8"]
    assert np.array_equal(np.load(out), full)


def test_band_helpers():
    sys.path.insert(0, ROOT)
    from ray_rust_b200 import bands

    for yres, br, world in [(4320, 16, 8), (70, 8, 3), (5, 16, 4), (33, 1, 2)]:
        seen = np.concatenate([bands.shard_rows(yres, br, r, world) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(yres))
    assert bands.max_shard_rows(4320, 16, 8) == 34 * 16
