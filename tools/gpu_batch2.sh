#!/bin/bash
# round-2 GPU batch 2: deferred LayerNorm (no LayerNorm launches) - parity + A/B against the round-1 schedule
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "deferred" 2>&1 | tail -15 > gpurun_out/r02_b2_kernels.log
python tools/bringup.py model_tc > gpurun_out/r02_b2_model.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_b2_pytest.log
python tools/ab_libs.py run r01 current > gpurun_out/r02_b2_ab.log 2>&1
tail -5 gpurun_out/r02_b2_kernels.log; tail -8 gpurun_out/r02_b2_model.log; tail -4 gpurun_out/r02_b2_pytest.log; cat gpurun_out/r02_b2_ab.log
