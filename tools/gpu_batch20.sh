#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/bringup.py model_tc > gpurun_out/r02_b20_model.log 2>&1
timeout 300 python tools/ab_libs.py run base19 current > gpurun_out/r02_b20_ab.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_b20_pytest.log
tail -3 gpurun_out/r02_b20_model.log; cat gpurun_out/r02_b20_ab.log; tail -4 gpurun_out/r02_b20_pytest.log
