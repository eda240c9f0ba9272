#!/bin/bash
# N GPUs of one node (gpurun --gpus N): headline bench + engine config 5 under torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --config 5 --steps 1 > gpurun_out/r02_config5_n$N.json 2> gpurun_out/r02_config5_n$N.err
for f in bench_n$N config5_n$N; do echo "== $f"; tail -c 1800 gpurun_out/r02_$f.json; tail -3 gpurun_out/r02_$f.err; done
