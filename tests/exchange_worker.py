"""Worker of tests/test_exchange_gpu.py::test_async_gather_across_processes (one process per GPU under torchrun)."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main(out_dir):
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend="nccl", device_id=dev)
    from cotr_b200.inference.sharding import AsyncGather
    for backend in ("peer", "nccl", "auto"):
        gather = AsyncGather((1, 1024, 2), dev, backend=backend)
        assert gather.backend == ("nccl" if backend == "nccl" else "peer"), gather.backend
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        want = torch.empty((world, 1024, 2), device=dev)
        # pipelined: many submits, one wait (what bench.py's timed loop does)
        for step in range(37):
            pred = torch.randn((1, 1024, 2), device=dev, generator=gen)
            gather.submit(pred)
        got = gather.wait()
        gather.check()
        dist.all_gather_into_tensor(want, pred)
        assert torch.equal(got, want), f"{backend}: pipelined submits"
        # lock step: submit / wait pairs (what the end-to-end call does), with a slow rank
        for step in range(12):
            if rank == step % world:
                torch.cuda._sleep(20_000_000)             # ~10 ms of device time on this rank only
            pred = torch.randn((1, 1024, 2), device=dev, generator=gen)
            gather.submit(pred)
            got = gather.wait().clone()
            dist.all_gather_into_tensor(want, pred)
            assert torch.equal(got, want), f"{backend}: lock step {step}"
        gather.check()
        gather.close()
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("peer ok\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
