"""GPU: the peer-memory result exchange (csrc/peer_exchange.cu, include/cotr_b200.h `cotr_exchange_*`).

Integer / byte work: every comparison is bit-exact.  The single-GPU tests run two ranks inside one process on the same
device (`cotr_exchange_connect_local`), which exercises the kernels, the slot / flag protocol and the failure reports;
the two-GPU tests (skipped on a one-GPU box) add real NVLink peer stores and the cross-process IPC mapping that
`bench.py --gpus N` uses."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ranks(world, block_bytes, slots, devices):
    from cotr_b200 import capi
    exs = [capi.NativeExchange(devices[r], r, world, block_bytes, slots) for r in range(world)]
    for e in exs:
        e.connect_local(exs)
    return exs


def _blocks(world, n_floats, step, devices):
    return [torch.arange(n_floats, dtype=torch.float32, device=f"cuda:{devices[r]}") + 1000.0 * r + 0.25 * step for r in range(world)]


def _run_protocol(devices):
    world, n = len(devices), 2048                     # 8 KB blocks: the headline shape's (1,1024,2) fp32
    exs = _ranks(world, 4 * n, 4, devices)
    streams = [torch.cuda.Stream(device=f"cuda:{d}") for d in devices]
    for step in range(1, 11):                         # laps the 4 slots more than twice
        blocks = _blocks(world, n, step, devices)
        torch.cuda.synchronize()
        seqs = [exs[r].push(blocks[r], stream=streams[r]) for r in range(world)]
        assert seqs == [step] * world
        outs = [torch.zeros(world * n, dtype=torch.float32, device=f"cuda:{devices[r]}") for r in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            exs[r].wait(step, out=outs[r], stream=streams[r])
        for d in devices:
            torch.cuda.synchronize(d)
        want = torch.cat([b.cpu() for b in blocks])
        for r in range(world):
            assert torch.equal(outs[r].cpu(), want), (step, r)
            assert exs[r].status() == 0
    for e in exs:
        e.close()


def test_two_ranks_on_one_device(built_lib):
    _run_protocol([0, 0])


def test_eight_ranks_on_one_device(built_lib):
    _run_protocol([0] * 8)


def test_ragged_and_large_blocks(built_lib):
    """Blocks of different sizes per rank (contexts split unevenly), larger than one push chunk, an empty one included."""
    world, cap = 3, 4 * 300000
    exs = _ranks(world, cap, 2, [0] * world)
    sizes = [300000, 0, 70004]                         # floats; 16-byte multiples
    blocks = [torch.randn(max(s, 4), device="cuda")[:s].contiguous() if s else torch.empty(0, device="cuda") for s in sizes]
    for r in range(world):
        exs[r].push(blocks[r])
    out = torch.zeros(sum(sizes), device="cuda")
    exs[2].wait(1, out=out, bytes_per_rank=[4 * s for s in sizes])
    torch.cuda.synchronize()
    assert torch.equal(out, torch.cat(blocks)) and exs[2].status() == 0
    for e in exs:
        e.close()


def test_lapped_reader_and_missing_peer_are_reported(built_lib):
    from cotr_b200 import capi
    exs = _ranks(2, 64, 2, [0, 0])
    b = torch.ones(16, device="cuda")
    assert exs[0].push(b) == 1
    for _ in range(3):                                 # rank 1 runs two laps ahead of rank 0's wait
        exs[1].push(b)
    out = torch.zeros(32, device="cuda")
    exs[0].wait(1, out=out)
    torch.cuda.synchronize()
    assert exs[0].status() == 2
    with pytest.raises(RuntimeError, match="overwritten"):
        exs[1].wait(1, out=out)                        # rank 1 itself pushed past step 1: refused on the host
    with pytest.raises(RuntimeError, match="never pushed"):
        exs[0].wait(5, out=out)
    with pytest.raises(RuntimeError, match="capacity"):
        exs[0].push(torch.ones(32, device="cuda"))
    for e in exs:
        e.close()
    # a peer that never publishes: bounded wait (~3 s), status 1, no hang
    exs = _ranks(2, 64, 2, [0, 0])
    exs[0].push(b)
    exs[0].wait(1, out=out)
    torch.cuda.synchronize()
    assert exs[0].status() == 1
    for e in exs:
        e.close()
    assert capi.lib().cotr_exchange_status(None) == -1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_on_two_devices(built_lib):
    _run_protocol([0, 1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_async_gather_across_processes(built_lib, tmp_path):
    """Two processes (torchrun, nccl): AsyncGather picks the peer transport, pipelined submits and push / wait pairs give
    exactly what an NCCL all-gather gives."""
    worker = os.path.join(REPO, "tests", "exchange_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", worker, str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=REPO)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for rank in range(2):
        assert open(tmp_path / f"rank{rank}.txt").read().strip() == "peer ok"
