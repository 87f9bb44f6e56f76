"""One frame over the GPUs of one box (one process per GPU), as a reusable host-side object.

The reference gathers finished rows to the caller thread over an mpsc channel (render.rs:846-886); here rank 0 owns the
frame in its device memory, every rank's render kernel stores its interleaved row bands straight into it over NVLink and
publishes a completion word there, and rank 0's stream waits on the words. `torch.distributed` is used only
to hand the CUDA IPC handle around; no collective is involved in a frame.

    frame = SharedDeviceFrame(dist, rank, world, local_rank, width, height)      # collective: every rank calls it
    frame.render(scene, ren, stream)        # every rank; on rank 0 the stream then owns the complete frame
    img = frame.download() if rank == 0 else None
    frame.close()
"""
import ctypes as C

import numpy as np

from . import ffi

BAND_ROWS = 16


class SharedDeviceFrame:
    def __init__(self, dist, rank, world, local_rank, width, height, band_rows=BAND_ROWS, timeout_ms=5000):
        self.lib = ffi.load()
        self.dist, self.rank, self.world, self.local_rank = dist, rank, world, local_rank
        self.width, self.height, self.band_rows, self.timeout_ms = int(width), int(height), int(band_rows), int(timeout_ms)
        self.frame_bytes = self.width * self.height * 3
        self._words_off = (self.frame_bytes + 255) // 256 * 256    # completion words (one per rank), then a status word
        self._ptr = C.c_void_p()
        self.epoch = 0
        handle = (C.c_uint8 * 64)()
        if rank == 0:
            ffi.check(self.lib.rr_device_alloc(local_rank, self._words_off + 512, C.byref(self._ptr)))
            ffi.check(self.lib.rr_device_memset(local_rank, C.c_void_p(self._ptr.value + self._words_off), 0, 512))
            ffi.check(self.lib.rr_ipc_export(self._ptr, handle))
        box = [bytes(handle)]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                ffi.check(self.lib.rr_ipc_open(local_rank, (C.c_uint8 * 64).from_buffer_copy(box[0]), C.byref(self._ptr)))
        self.flags = C.c_void_p(self._ptr.value + self._words_off)
        self.status = C.c_void_p(self._ptr.value + self._words_off + 256)

    @property
    def device_ptr(self):
        """Device address of the RGB8 frame (row-major, width * 3 bytes per row): local on rank 0, a peer mapping elsewhere."""
        return self._ptr.value

    def params(self, ren):
        """This rank's shard of the frame: rr_frame_params with the interleaved row bands filled in."""
        return ren.frame_params(self.band_rows, self.rank, self.world) if self.world > 1 else ren.frame_params()

    def render(self, scene, ren, stream=0):
        """Queue this rank's bands on `stream` (a cudaStream_t value); on rank 0 also the wait for every rank's completion
        word. Returns the frame's epoch."""
        self.epoch += 1
        p = self.params(ren)
        sp = C.c_void_p(stream) if stream else None
        ffi.check(self.lib.rr_render_rgb8_placed_signal_device(scene.handle, C.byref(p), self._ptr, self.width * 3, self.flags,
                                                               self.epoch, sp))
        if self.rank == 0:
            ffi.check(self.lib.rr_fence_wait_device(self.local_rank, self.flags, self.world, self.epoch, self.timeout_ms,
                                                    self.status, sp))
        return self.epoch

    def timed_out(self):
        """Rank 0: did any wait give up (a rank that never rendered)? Synchronises the device."""
        st = (C.c_uint32 * 1)()
        ffi.check(self.lib.rr_device_read(self.local_rank, self.status, st, 4))
        return st[0] != 0

    def download(self):
        """Rank 0: the frame as a (height, width, 3) uint8 array (synchronous copy; call after the stream has finished)."""
        out = np.empty((self.height, self.width, 3), dtype=np.uint8)
        ffi.check(self.lib.rr_device_read(self.local_rank, self._ptr, out.ctypes.data_as(C.c_void_p), self.frame_bytes))
        return out

    def close(self):
        if self._ptr.value is None:
            return
        if self.rank == 0:
            if self.world > 1:
                self.dist.barrier()   # peers unmap first
            self.lib.rr_device_free(self.local_rank, self._ptr)
        else:
            self.lib.rr_ipc_close(self.local_rank, self._ptr)
            self.dist.barrier()
        self._ptr = C.c_void_p()
