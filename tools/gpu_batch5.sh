#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_b5_pytest.log
python tools/ab_libs.py run r01 current ld_ca no_order > gpurun_out/r02_b5_ab.log 2>&1
python tools/bringup.py gemm_timeline > gpurun_out/r02_b5_gemm_timeline.log 2>&1
COTR_PROFILE_B=8 python tools/bringup.py launch_profile > gpurun_out/r02_b5_launch_profile_b8.log 2>&1
tail -6 gpurun_out/r02_b5_pytest.log; cat gpurun_out/r02_b5_ab.log; grep -E "a_ln|res_ln" gpurun_out/r02_b5_gemm_timeline.log; tail -3 gpurun_out/r02_b5_launch_profile_b8.log
