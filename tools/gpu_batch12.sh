#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_timing.py "0:auto,65536:explicit-ln,524288:deferred-ln" > gpurun_out/r02_b12_schedules_ab.log 2>&1
COTR_TRACE_N=140 COTR_TRACE_SLOTS=44,45,50 timeout 200 python tools/bringup.py forward_trace > gpurun_out/r02_b12_forward_trace.log 2>&1
COTR_TRACE_VARIANT=786432 COTR_TRACE_N=140 COTR_TRACE_SLOTS=44,45,49 timeout 200 python tools/bringup.py forward_trace > gpurun_out/r02_b12_forward_trace_dataflow.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_b12_pytest.log
cat gpurun_out/r02_b12_schedules_ab.log; tail -4 gpurun_out/r02_b12_pytest.log; grep "cycle stamps" gpurun_out/r02_b12_forward_trace.log | cut -c1-500;  grep -E "launch  (44|45|46):" gpurun_out/r02_b12_forward_trace.log | cut -c1-170; echo DATAFLOW; grep "cycle stamps" gpurun_out/r02_b12_forward_trace_dataflow.log | cut -c1-500; grep -E "launch  (44|45|46):" gpurun_out/r02_b12_forward_trace_dataflow.log | cut -c1-170
