"""CPU, world_size 2, gloo: the bucketed gradient averaging used for data-parallel training and the halo-tiled
inference partitioning (no GPU, no native library)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rrdbnet_oracle as orc
from sr_gan_fd_b200 import tile


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sr_gan_fd_b200.dist import GradBucketReducer
        torch.manual_seed(100 + rank)
        flat = torch.randn(1000)
        mine = flat.clone()
        red = GradBucketReducer(average=True, min_bucket_numel=150)
        # buckets arrive tail-first, in descending address order, like b200sr_backward announces them
        for off, cnt in [(800, 200), (700, 100), (600, 100), (100, 500), (0, 100)]:
            red.bucket_ready(flat, off, cnt)
        red.finish(flat)
        gathered = [torch.empty(1000) for _ in range(world)]
        dist.all_gather(gathered, mine)
        expect = sum(gathered) / world
        ok = torch.allclose(flat, expect, atol=1e-6)
        ret[rank] = (bool(ok), red.buckets_seen)

        # DP equivalence of the math: mean-L1 over the global batch == average of per-rank mean-L1 gradients
        params = orc.init_params(seed=0, channels=8, growth=4, num_blocks=1, upscale_factor=2)
        g = torch.Generator().manual_seed(7)
        lr = torch.rand(4, 3, 8, 8, generator=g)
        gt = torch.rand(4, 3, 16, 16, generator=g)
        _, _, g_all = orc.rrdbnet_l1_step(params, lr, gt)
        _, _, g_loc = orc.rrdbnet_l1_step(params, lr[rank * 2:rank * 2 + 2], gt[rank * 2:rank * 2 + 2])
        flat_loc = torch.cat([v.flatten() for v in g_loc.values()])
        red2 = GradBucketReducer(average=True)
        n = flat_loc.numel()
        red2.bucket_ready(flat_loc, n // 2, n - n // 2)
        red2.bucket_ready(flat_loc, 0, n // 2)
        red2.finish(flat_loc)
        flat_all = torch.cat([v.flatten() for v in g_all.values()])
        ret[rank + world] = orc.rel_l2(flat_loc, flat_all)
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        ok, seen = ret[r]
        assert ok
        assert seen == [(800, 200), (700, 100), (600, 100), (100, 500), (0, 100)]
        assert ret[r + world] < 1e-5


def test_band_planning():
    bands = tile.plan_bands(1024, 8, 16)
    assert [b[1] - b[0] for b in bands] == [128] * 8
    assert bands[0][2] == 0 and bands[-1][3] == 0 and all(b[2] == 16 for b in bands[1:])
    assert sum(b[1] - b[0] for b in tile.plan_bands(123, 5, 8)) == 123
    owned = [tile.bands_for_rank(bands, r, 3) for r in range(3)]
    assert sorted(sum(owned, [])) == list(range(8))
    assert abs(tile.redundant_fraction(1024, 8, 16) - 14 * 16 / 1024) < 1e-12
    with pytest.raises(ValueError):
        tile.plan_bands(4, 5, 1)


def test_tiled_equals_whole_frame_with_oracle_net():
    """Tile driver logic with the fp32 oracle as the network: halo >= receptive field is exact, small halo is close."""
    params = orc.init_params(seed=1, channels=8, growth=4, num_blocks=1, upscale_factor=2)
    params = orc.in_range_fixture(params)
    net = lambda t: orc.rrdbnet_forward(params, t)
    lr = torch.rand(1, 3, 48, 20)
    whole = net(lr)
    outs = []
    for rank in range(3):
        out, (r0, r1) = tile.tiled_forward(net, lr, 2, num_bands=3, halo=24, rank=rank, world_size=3)
        outs.append(out[:, :, r0:r1])
    stitched = torch.cat(outs, 2)
    assert stitched.shape == whole.shape
    assert orc.rel_l2(stitched, whole) < 1e-6       # 1 RRDB: receptive radius 15+3 LR px < 24
    out0, _ = tile.tiled_forward(net, lr, 2, num_bands=4, halo=0)
    assert orc.rel_l2(out0, whole) > 1e-4            # no halo visibly differs
