#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_libs.py run r01 current ldnc > gpurun_out/r02_b11_ab.log 2>&1
timeout 300 python tools/ab_timing.py "0:default,524288:deferred-ln,786432:deferred-ln+dataflow" > gpurun_out/r02_b11_schedules_ab.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_b11_pytest.log
COTR_TRACE_N=140 COTR_TRACE_SLOTS=46,47,53,54 timeout 200 python tools/bringup.py forward_trace > gpurun_out/r02_b11_forward_trace.log 2>&1
cat gpurun_out/r02_b11_ab.log gpurun_out/r02_b11_schedules_ab.log; tail -4 gpurun_out/r02_b11_pytest.log; grep "cycle stamps" gpurun_out/r02_b11_forward_trace.log | cut -c1-600
