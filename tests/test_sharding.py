"""Multi-GPU host logic on CPU: partition arithmetic, and the band gather over a
real 2-process gloo group (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from motionestimation_b200 import sharding, frames


def test_pair_slices_tile_the_batch():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [sharding.pair_slice(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.pair_slice(4, 2, 2)


def test_row_costs_match_pixel_compares():
    for (W, H, B, R) in [(1920, 1080, 16, 32), (352, 288, 8, 12), (37, 29, 5, 3)]:
        assert sum(sharding.row_costs(W, H, B, R)) == frames.pixel_compares(W, H, B, R)


def test_band_rows_tile_and_balance():
    W, H, B, R = 3840, 2160, 16, 64
    nby = -(-H // B)
    w = sharding.row_costs(W, H, B, R)
    for world in (1, 2, 4, 8):
        spans = [sharding.band_rows(nby, world, r, w) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == nby
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        cost = [sum(w[b:e]) for b, e in spans]
        assert max(cost) <= 1.15 * (sum(w) / world)          # balanced within 15 %
        plain = [sharding.band_rows(nby, world, r) for r in range(world)]
        assert plain[0][0] == 0 and plain[-1][1] == nby
    # fewer rows than ranks: some bands are empty, the tiling still holds
    spans = [sharding.band_rows(2, 4, r) for r in range(4)]
    assert sum(e - b for b, e in spans) == 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, nby, nbx, npairs, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(npairs * nby * nbx, dtype=torch.int32).reshape(npairs, nby * nbx)
        w = [1.0 + (i % 3) for i in range(nby)]                # uneven costs -> uneven bands
        b0, b1 = sharding.band_rows(nby, world, rank, w)
        local = full[:, b0 * nbx: b1 * nbx].contiguous()
        got = sharding.gather_bands(local, (b0, b1), nby, nbx)
        ok = bool(torch.equal(got, full))
        # pair sharding: every rank's slice, all-gathered, is the whole batch in order
        p0, p1 = sharding.pair_slice(npairs, world, rank)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([p1 - p0]))
        ok = ok and sum(int(s) for s in sizes) == npairs
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nby,nbx,npairs", [(9, 5, 3), (2, 7, 1), (135, 4, 2)])
def test_gather_bands_two_ranks_gloo(nby, nbx, npairs):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_gather_worker, args=(world, port, nby, nbx, npairs, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


@pytest.mark.gpu
def test_search_banded_single_rank_nccl():
    """search_banded over a 1-rank NCCL group equals the plain full search."""
    import motionestimation_b200 as me
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        W, H, B, R = 352, 288, 16, 32
        cur8, ref8 = me.foreman(2), me.foreman(1)
        cur, ref = torch.from_numpy(cur8).cuda(), torch.from_numpy(ref8).cuda()
        with me.Estimator(W, H, B, R) as est:
            res = sharding.search_banded(est, cur, ref, W, W * H, 1)
            exp = est.search_u8(cur8, ref8)
        torch.cuda.synchronize()
        assert np.array_equal(res["mvx"].cpu().numpy(), exp["mvx"])
        assert np.array_equal(res["mvy"].cpu().numpy(), exp["mvy"])
        assert np.array_equal(res["ssd"].cpu().numpy().view(np.uint32), exp["ssd"])
    finally:
        dist.destroy_process_group()
