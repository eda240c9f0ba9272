#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_libs.py run base13 current > gpurun_out/r02_b14_ab.log 2>&1
COTR_PROFILE_B=32 COTR_PROFILE_Q=1 timeout 300 python tools/bringup.py launch_profile > gpurun_out/r02_b14_launch_profile_b32.log 2>&1
cat gpurun_out/r02_b14_ab.log; tail -3 gpurun_out/r02_b14_launch_profile_b32.log
