#!/bin/bash
# final single-GPU records of the round: headline bench, reference arm, engine configs, smoke
mkdir -p gpurun_out
T=r02_b21
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 900 python bench.py --config 3 --steps 1 > gpurun_out/${T}_config3.json 2> gpurun_out/${T}_config3.err
timeout 600 python bench.py --config 5 --steps 1 > gpurun_out/${T}_config5.json 2> gpurun_out/${T}_config5.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
for f in bench_n1 bench_ref config3 config5; do echo "== $f"; tail -c 2500 gpurun_out/${T}_$f.json; tail -3 gpurun_out/${T}_$f.err; done; tail -2 gpurun_out/${T}_smoke.log
