#!/bin/bash
# ncu evidence of the round: launch list of the bench command + full captures of the dominant kernels (one GPU)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_ncu_launches_bench_n1.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 400 -c 6 -o gpurun_out/r02_ncu_gemm python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r02_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 40 -c 2 -o gpurun_out/r02_ncu_attn python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r02_ncu_attn.log 2>&1
ls -la gpurun_out/r02_ncu*; tail -2 gpurun_out/r02_ncu_bench.log
