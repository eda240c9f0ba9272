#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_batch_exchange.sh N TAG   - peer-memory result exchange: tests + headline bench at N ranks
N=${1:-2}; TAG=${2:-r02_exchange}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_exchange_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -c 1500 gpurun_out/${TAG}_bench_n$N.json; tail -5 gpurun_out/${TAG}_bench_n$N.err
