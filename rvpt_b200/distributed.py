"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over
NVLink/NVSwitch), the frame sharded by 16x16 pixel tile, one all-gather of the
per-tile result per frame.

The path has exactly one exchange step — assembling the image — and no
data-path collective before it: every rank seeds its RNG from global (x, y, W,
frame) (util.glsl:35-36) and owns the running mean of its tiles. The gather is
"in place": each rank's kernels write their tiles straight into the rank's slot
of the gather buffer (rvpt_b200_set_external_tiles), so there is no pack pass.

The gather of frame f overlaps the rendering of frame f+1: two gather buffers
alternate, the collective runs asynchronously on NCCL's stream and the untile
on a side stream; the render stream only waits (on the device) for the gather
that last used the buffer it is about to overwrite.
"""
from __future__ import annotations

import numpy as np

from . import tiles


class FrameGather:
    """Owns two gather buffers ([nranks][n_local_padded*256] rgba8 as int32)
    and the assembled raster image."""

    def __init__(self, engine, dist, torch, device, untile_on_all_ranks: bool = False):
        self.engine, self.dist, self.torch = engine, dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        ti = engine.tile_info()
        assert ti.nranks == self.world and ti.rank == self.rank
        self.slot_elems = ti.n_local_tiles_padded * tiles.TILE_PIXELS
        self.gathered = [torch.zeros(self.world * self.slot_elems, dtype=torch.int32, device=device)
                         for _ in range(2)]
        self.my_slot = [g[self.rank * self.slot_elems:(self.rank + 1) * self.slot_elems]
                        for g in self.gathered]
        self.raster = torch.zeros(engine.height * engine.width, dtype=torch.int32, device=device)
        self.untiles = self.rank == 0 or untile_on_all_ranks
        self.side = torch.cuda.Stream(device=device)
        self.pending = [None, None]   # (work, untile-done event) of the last gather per buffer
        self.parity = 0
        # every wait below is enqueued on torch's current stream, so the engine must launch its
        # kernels there too — on its own stream the collective could read a slot the frame kernel
        # is still writing, and the next frame could overwrite a slot a gather is still reading
        self.stream = torch.cuda.current_stream(device)
        engine.set_stream(self.stream.cuda_stream)
        engine.set_external_tiles(None, self.my_slot[0].data_ptr())

    def _check_stream(self) -> None:
        cur = self.torch.cuda.current_stream()
        if cur.cuda_stream != self.stream.cuda_stream:
            raise RuntimeError("FrameGather is bound to the stream that was current when it was created; "
                               "call it under that stream (torch.cuda.stream(...))")

    def begin_frame(self) -> None:
        """Call before render_frame: points the kernels at the buffer of this
        frame and makes the render stream wait for its previous gather."""
        self._check_stream()
        b = self.parity
        pend = self.pending[b]
        if pend is not None:
            work, done = pend
            work.wait()                       # render stream waits for the NCCL kernel
            if done is not None:
                self.torch.cuda.current_stream().wait_event(done)
            self.pending[b] = None
        self.engine.set_external_tiles(None, self.my_slot[b].data_ptr())

    def end_frame(self) -> None:
        """Call after render_frame: ONE collective for the frame, asynchronous;
        the scatter into the raster image follows it on the side stream."""
        self._check_stream()
        b = self.parity
        work = self.dist.all_gather_into_tensor(self.gathered[b], self.my_slot[b], async_op=True)
        done = None
        if self.untiles:
            with self.torch.cuda.stream(self.side):
                work.wait()                   # side stream waits for the NCCL kernel
                self.engine.untile(self.gathered[b].data_ptr(), self.raster.data_ptr(), 4,
                                   self.side.cuda_stream)
                done = self.torch.cuda.Event()
                done.record(self.side)
        self.pending[b] = (work, done)
        self.parity ^= 1

    def gather(self) -> None:
        """Synchronous-in-stream variant: gather + untile of the frame just
        rendered, ordered before anything launched later on the render stream."""
        self.end_frame()
        self.flush()

    def flush(self) -> None:
        """Makes the render stream wait for every gather still in flight."""
        for b in range(2):
            pend = self.pending[b]
            if pend is not None:
                work, done = pend
                work.wait()
                if done is not None:
                    self.torch.cuda.current_stream().wait_event(done)
                self.pending[b] = None

    def image(self) -> np.ndarray:
        """HxWx4 uint8 (valid on ranks that untile)."""
        self.flush()
        self.torch.cuda.synchronize()
        a = self.raster.cpu().numpy().view(np.uint8)
        return a.reshape(self.engine.height, self.engine.width, 4)


class PeerOutput:
    """Fused gather over NVLink peer memory (the default N > 1 path).

    Rank 0 owns the raster image; every other rank maps it through CUDA IPC and
    its frame kernels store finished rgba8 pixels directly into it while they
    compute — no collective, no pack, no untile per frame. NCCL only carries the
    one-off handle exchange and the barrier that ends a step."""

    def __init__(self, engine, dist, torch, device):
        self.engine, self.dist, self.torch = engine, dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        # both images of the display rank's double-buffered result (see read_image_async)
        for second in (False, True):
            handle = torch.zeros(64, dtype=torch.uint8, device=device)
            if self.rank == 0:
                handle.copy_(torch.frombuffer(bytearray(engine.export_output(second)), dtype=torch.uint8))
            dist.broadcast(handle, src=0)
            if self.rank != 0:
                engine.attach_output(bytes(handle.cpu().numpy().tobytes()), second)
        dist.barrier()

    def finish(self) -> None:
        """Every rank's frames have landed in rank 0's image after this."""
        self.torch.cuda.synchronize()
        self.dist.barrier()

    def read_image_async(self, out_pinned: np.ndarray) -> None:
        """End of a step with the read-back overlapping the next step: every rank's pixels are in
        the display rank's current image (stream completion + barrier); the display rank starts
        copying it to `out_pinned` and every rank switches to the other image for what it renders
        next. engine.wait_output() on rank 0 (before the buffer is reused) completes the copy."""
        if self.rank == 0:
            # the copy started one step ago read the image every rank is about to switch back to:
            # it must have landed before the barrier lets them (it overlapped this step's frames)
            self.engine.wait_output()
        self.finish()
        if self.rank == 0:
            self.engine.read_output_rgba8_async(out_pinned)
        else:
            self.engine.flip_output()

    def image(self) -> np.ndarray:
        self.finish()
        if self.rank != 0:
            return None
        return self.engine.read_output_rgba8()
