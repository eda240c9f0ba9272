"""CPU: the N > 1 path (pair sharding + result gather) with world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cotr_b200.inference.sharding import pair_range


def test_pair_ranges_partition_the_batch():
    for n in (1, 2, 7, 8, 64):
        for world in (1, 2, 3, 8):
            blocks = [pair_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [e - s for s, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cotr_b200.inference.sharding import forward_sharded
    from oracle.fake_model import FakeCOTR
    g = torch.Generator().manual_seed(0)
    img = torch.randn(n_pairs, 3, 256, 512, generator=g)
    q = torch.rand(n_pairs, 33, 2, generator=g)
    model = FakeCOTR()
    out = forward_sharded(model, img, q)
    ref = FakeCOTR()(img, q)['pred_corrs']
    results[rank] = (bool(torch.equal(out, ref)), [c[0][0] for c in model.calls])
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [5, 2, 1])
def test_sharded_forward_equals_unsharded_gloo(n_pairs):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_pairs, results), nprocs=world, join=True)
        results = dict(results)
    assert all(results[r][0] for r in range(world))                     # bitwise equal to the unsharded forward
    assert sum(sum(results[r][1]) for r in range(world)) == n_pairs     # every pair processed exactly once
