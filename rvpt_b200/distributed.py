"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over
NVLink/NVSwitch), the frame sharded by 16x16 pixel tile, one all-gather of the
per-tile result per frame.

The path has exactly one exchange step — assembling the image — and no
data-path collective before it: every rank seeds its RNG from global (x, y, W,
frame) (util.glsl:35-36) and owns the running mean of its tiles. The gather is
"in place": each rank's kernels write their tiles straight into the rank's slot
of the gather buffer (rvpt_b200_set_external_tiles), so there is no pack pass.
"""
from __future__ import annotations

import numpy as np

from . import tiles


class FrameGather:
    """Owns the gather buffer ([nranks][n_local_padded*256] rgba8 as int32) and
    the assembled raster image on every rank."""

    def __init__(self, engine, dist, torch, device):
        self.engine, self.dist, self.torch = engine, dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        ti = engine.tile_info()
        assert ti.nranks == self.world and ti.rank == self.rank
        self.slot_elems = ti.n_local_tiles_padded * tiles.TILE_PIXELS
        self.gathered = torch.zeros(self.world * self.slot_elems, dtype=torch.int32, device=device)
        self.raster = torch.zeros(engine.height * engine.width, dtype=torch.int32, device=device)
        self.my_slot = self.gathered[self.rank * self.slot_elems:(self.rank + 1) * self.slot_elems]
        engine.set_external_tiles(None, self.my_slot.data_ptr())

    def gather(self, untile_on_all_ranks: bool = False) -> None:
        """One collective per frame; rank 0 (or every rank) scatters the
        gathered tiles into the raster image with rvpt_b200_untile()."""
        self.dist.all_gather_into_tensor(self.gathered, self.my_slot)
        if self.rank == 0 or untile_on_all_ranks:
            self.engine.untile(self.gathered.data_ptr(), self.raster.data_ptr(), 4)

    def image(self) -> np.ndarray:
        """HxWx4 uint8 (valid on ranks that untile)."""
        self.torch.cuda.synchronize()
        a = self.raster.cpu().numpy().view(np.uint8)
        return a.reshape(self.engine.height, self.engine.width, 4)
