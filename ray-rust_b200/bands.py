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


def max_shard_rows(yres, band_rows, world):
    return max(len(shard_rows(yres, band_rows, r, world)) for r in range(world))


def unpack(gathered, yres, band_rows, world):
    """gathered[rank, local_row, ...] (padded to max_shard_rows) -> frame[iy, ...] (numpy mirror of the device kernel)."""
    out = np.empty((yres,) + gathered.shape[2:], dtype=gathered.dtype)
    for r in range(world):
        rows = shard_rows(yres, band_rows, r, world)
        out[rows] = gathered[r, :len(rows)]
    return out
