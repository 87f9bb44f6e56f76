"""Row-band sharding of a frame over ranks: band b = iy // band_rows belongs to rank b % world.

Host-side helpers shared by bench.py and the multi-rank tests; the device-side equivalents are
rr_frame_params.band_* (rendering one shard) and rr_bands_unpack_device (un-interleaving a gather).
"""
import numpy as np


def shard_rows(yres, band_rows, rank, world):
    """Image rows rendered by `rank`, in the packed order the kernels write them."""
    if world <= 1:
        return np.arange(yres)
    iy = np.arange(yres)
    return iy[(iy // band_rows) % world == rank]


def weighted_spans(world, rank0_slots, other_slots):
    """Unequal shares (rr_frame_params.band_span): a period of rank0_slots + (world - 1) * other_slots band slots, of which
    rank 0 owns the first rank0_slots and every other rank other_slots. Returns [(band_index, band_span)] per rank and the
    period (= band_count). The owner of a multi-GPU device frame receives all other ranks' rows over its NVLink ports;
    rows it renders itself do not cross them, so when that ingress bounds the step, rank 0 should render more."""
    spans, at = [], 0
    for r in range(world):
        n = rank0_slots if r == 0 else other_slots
        spans.append((at, n))
        at += n
    return spans, at


def span_rows(yres, band_rows, band_index, band_span, band_count):
    """Image rows of the shard that owns slots [band_index, band_index + band_span) of every period of band_count bands."""
    iy = np.arange(yres)
    slot = (iy // band_rows) % band_count
    return iy[(slot >= band_index) & (slot < band_index + band_span)]


def max_shard_rows(yres, band_rows, world):
    return max(len(shard_rows(yres, band_rows, r, world)) for r in range(world))


def unpack(gathered, yres, band_rows, world):
    """gathered[rank, local_row, ...] (padded to max_shard_rows) -> frame[iy, ...] (numpy mirror of the device kernel)."""
    out = np.empty((yres,) + gathered.shape[2:], dtype=gathered.dtype)
    for r in range(world):
        rows = shard_rows(yres, band_rows, r, world)
        out[rows] = gathered[r, :len(rows)]
    return out
