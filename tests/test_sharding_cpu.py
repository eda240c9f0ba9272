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


def _engine_worker(rank, world, port, results):
    """SPMD engines over ShardedCOTR: every rank runs the same scheduler, each model call is split over the ranks."""
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import contextlib
    import io
    from cotr_b200.inference.sharding import ShardedCOTR
    from cotr_b200.inference.sparse_engine import FasterSparseEngine, SparseEngine
    from cotr_b200.inference.inference_helper import cotr_flow
    from cotr_b200.utils.utils import fix_randomness
    from oracle.fake_model import FakeCOTR, synthetic_image
    img_a = synthetic_image(1, 300, 400)
    img_b = synthetic_image(2, 360, 288)
    q = np.random.RandomState(5).uniform([5, 5], [395, 295], size=(21, 2))
    zooms = np.linspace(0.5, 0.0625, 4)
    ok = []
    n_local = 0
    with contextlib.redirect_stdout(io.StringIO()):
        for cls, kw in ((SparseEngine, {}), (FasterSparseEngine, {"max_load": 8})):
            out = []
            for sharded in (True, False):
                fix_randomness(0)
                fake = FakeCOTR()
                model = ShardedCOTR(fake) if sharded else fake
                eng = cls(model, 5, mode='tile', **kw)
                out.append(eng.cotr_corr_multiscale(img_a, img_b, zooms, 1, max_corrs=21, queries_a=q.copy(), force=True))
                if sharded:
                    n_local += sum(c[0][0] for c in fake.calls)
                else:
                    n_total = sum(c[0][0] for c in fake.calls)
            ok.append(bool(np.array_equal(out[0], out[1])) and out[0].shape[0] > 0)
        # single context, many queries: the dense pass splits its 131 072 queries over the ranks
        flow = [cotr_flow(m, img_a[:256, :256], img_b[:256, :256]) for m in (ShardedCOTR(FakeCOTR()), FakeCOTR())]
        ok.append(all(np.array_equal(a, b) for a, b in zip(flow[0], flow[1])))
    results[rank] = (ok, n_local, n_total)
    dist.destroy_process_group()


def test_sharded_engines_equal_unsharded_gloo():
    """BASELINE.json configs[4] host logic: SparseEngine / FasterSparseEngine / cotr_flow over ShardedCOTR with world
    size 2 return exactly what they return on one rank, and the contexts really are divided between the ranks."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_engine_worker, args=(world, port, results), nprocs=world, join=True)
        results = dict(results)
    for r in range(world):
        assert all(results[r][0]), results[r][0]
    # every context of the (last) engine run was processed once across the ranks, and no rank did all of them
    assert 0 < results[0][1] and 0 < results[1][1]


def _exchange_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cotr_b200.inference.sharding import open_exchange
    # no CUDA device here: every rank must come back with the same answer ("no peer transport"), without hanging or raising
    results[rank] = open_exchange(8192, torch.device("cpu")) is None and open_exchange(8200, torch.device("cpu")) is None
    dist.destroy_process_group()


def test_peer_exchange_declines_collectively_without_cuda():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_exchange_worker, args=(world, port, results), nprocs=world, join=True)
        assert dict(results) == {0: True, 1: True}


def test_engine_bench_compares_survivor_sets():
    """tools/engine_bench.compare_runs: overlap and differences of two runs whose survivors differ near the cut."""
    import numpy as np
    from tools.engine_bench import compare_runs
    a = np.array([[1., 1., 5., 5.], [2., 2., 6., 6.], [3., 3., 7., 7.]])
    b = np.array([[2., 2., 6., 6.5], [1., 1., 5., 5.], [9., 9., 0., 0.]])
    r = compare_runs(a, b)
    assert r["same_source_points"] is False and r["common_source_points"] == 2 and r["of"] == 3
    assert r["max_abs_diff_px"] == 0.5 and r["median_abs_diff_px"] == 0.25
    assert compare_runs(a, a.copy())["same_source_points"] is True


def test_async_gather_is_transparent_on_one_rank():
    """Without a process group (one GPU) `AsyncGather` hands the block straight back: no streams, no exchange."""
    from cotr_b200.inference.sharding import AsyncGather
    g = AsyncGather((1, 4, 2), torch.device("cpu"))
    assert g.backend == "local" and g.exchange is None and g.side is None
    pred = torch.arange(8, dtype=torch.float32).view(1, 4, 2)
    g.submit(pred)
    assert g.wait() is pred and g.check() == 0
    g.close()
