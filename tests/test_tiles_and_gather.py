"""The N > 1 path on CPU: tile partition arithmetic and the per-frame gather
(gloo, world_size 2 and 3). No radiance is computed by the product here — the
ranks' tile buffers are cut out of an oracle-rendered image, all-gathered and
re-assembled with the same layout rules the kernels and rvpt_b200_untile use."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PINNED_POSE


def test_slot_maps_partition_the_image(rv):
    from rvpt_b200 import tiles
    for (w, h) in ((208, 120), (16, 16), (33, 17), (1920, 1080)):
        for nranks in (1, 2, 3, 8):
            seen = np.zeros(w * h, int)
            padded = set()
            for r in range(nranks):
                idx = tiles.slot_to_raster(w, h, r, nranks)
                owned, pad = tiles.local_tile_counts(w, h, r, nranks)
                padded.add(pad)
                assert len(idx) == pad * 256
                ok = idx >= 0
                np.add.at(seen, idx[ok], 1)
                # slots past the owned tiles are padding
                assert (idx[owned * 256:] == -1).all()
            assert (seen == 1).all(), "every pixel belongs to exactly one rank"
            assert len(padded) == 1, "all ranks pad to the same tile count"


def test_tile_roundtrip_numpy(rv):
    from rvpt_b200 import tiles
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, size=(90, 200, 4), dtype=np.uint8)
    for nranks in (1, 2, 4, 5):
        parts = np.stack([tiles.tiles_from_raster(img, r, nranks) for r in range(nranks)])
        back = tiles.raster_from_gathered(parts, 200, 90, nranks)
        assert np.array_equal(back, img)


def test_warp_blocks_are_contiguous(rv):
    """Each warp's 8x4 pixel block is 32 consecutive slots (coalesced 512 B of
    float4 accumulation per warp)."""
    from rvpt_b200 import tiles
    px, py = tiles.in_tile_offsets()
    for w in range(8):
        bx, by = px[w * 32:(w + 1) * 32], py[w * 32:(w + 1) * 32]
        assert bx.max() - bx.min() == 7 and by.max() - by.min() == 3
        assert len(set(zip(bx.tolist(), by.tolist()))) == 32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, image_path, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rvpt_b200 import tiles
    img = np.load(image_path)
    # what this rank's kernels would have written into its slot of the gather buffer
    mine = torch.from_numpy(tiles.tiles_from_raster(img, rank, world).view(np.int32).reshape(-1))
    gathered = torch.zeros(world * mine.numel(), dtype=torch.int32)
    dist.all_gather_into_tensor(gathered, mine)  # the ONE collective per frame
    if rank == 0:
        g = gathered.numpy().view(np.uint8).reshape(world, -1, 4)
        np.save(out_path, tiles.raster_from_gathered(g, w, h, world))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_gather_reassembles_the_frame(rv, oracle_mod, builtin, tmp_path, world):
    w, h = 208, 120
    cam = rv.camera_data(translation=PINNED_POSE, aspect=w / h)
    ora = oracle_mod.OracleRenderer(w, h, builtin.triangles, builtin.materials, builtin.nodes)
    for f in range(2):
        ora.render_frame(rv.default_settings(frame=f), cam)
    image_path, out_path = tmp_path / "img.npy", tmp_path / "out.npy"
    np.save(image_path, ora.result)
    mp.spawn(_worker, args=(world, _free_port(), w, h, str(image_path), str(out_path)),
             nprocs=world, join=True)
    assert np.array_equal(np.load(out_path), ora.result)
