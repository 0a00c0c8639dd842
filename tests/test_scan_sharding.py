"""Multi-GPU host logic on CPU: view partitioning and the single end-of-scan gather, world_size 2 over gloo."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_views_is_a_partition():
    from aarmvsnet_b200 import scan
    for n in (1, 7, 48, 49):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                seen += list(scan.shard_views(n, world, r))
            assert seen == list(range(n))
            sizes = [len(scan.shard_views(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == scan.max_share(n, world)
    assert [scan.max_share(49, g) for g in (2, 4, 8)] == [25, 13, 7]          # SURVEY.md §8e scaling ceiling
    with pytest.raises(ValueError):
        scan.shard_views(4, 2, 2)


def test_shard_units_over_several_scans():
    from aarmvsnet_b200 import scan
    counts = [49, 49, 5]
    for world in (1, 2, 8):
        seen = []
        for r in range(world):
            seen += scan.shard_units(counts, world, r)
        assert seen == [(s, v) for s, n in enumerate(counts) for v in range(n)]          # a partition, in order
    sizes = [len(scan.shard_units([49, 49], 8, r)) for r in range(8)]
    assert max(sizes) == 13 and min(sizes) == 12                    # per-scan sharding would leave 14 on the slowest rank
    assert scan.shard_units([3], 2, 1) == [(0, 2)]


def _worker(rank, world, port, n_views, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aarmvsnet_b200 import scan

    def forward_view(v):          # stands in for EMVSNet on this rank's GPU: recognisable per-view maps
        return torch.full((2, 4, 6), float(v)) + torch.arange(2.0).view(2, 1, 1) * 0.5

    full = scan.run_scan(forward_view, n_views)
    ok = full.shape == (n_views, 2, 4, 6) and all(torch.equal(full[v], forward_view(v)) for v in range(n_views))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [5, 8])
def test_scan_gather_world2_gloo(n_views):
    world = 2
    port = 29500 + (os.getpid() % 2000) + n_views
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_views, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aarmvsnet_b200.training import average_gradients
    flat = torch.arange(10.0) * (rank + 1)                  # rank 0: g, rank 1: 2g -> mean 1.5 g
    average_gradients(flat)
    out[rank] = bool(torch.allclose(flat, torch.arange(10.0) * 1.5))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_average_world2_gloo():
    """The one collective of the data-parallel training step (aarmvsnet_b200.training.average_gradients): an all-reduce of
    the flat gradient buffer, here over gloo with two CPU processes."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
    from aarmvsnet_b200.training import average_gradients
    t = torch.ones(3)
    assert average_gradients(t) is t and torch.equal(t, torch.ones(3))      # single process: untouched
