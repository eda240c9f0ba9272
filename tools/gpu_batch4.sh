#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_b4_pytest.log
python tools/ab_libs.py run r01 current > gpurun_out/r02_b4_ab.log 2>&1
python tools/bringup.py gemm_timeline > gpurun_out/r02_b4_gemm_timeline.log 2>&1
python tools/bringup.py launch_profile > gpurun_out/r02_b4_launch_profile.log 2>&1
tail -4 gpurun_out/r02_b4_pytest.log; cat gpurun_out/r02_b4_ab.log; grep -E "a_ln|res_ln" gpurun_out/r02_b4_gemm_timeline.log; tail -3 gpurun_out/r02_b4_launch_profile.log
