#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_b9_pytest.log
timeout 300 python tools/ab_timing.py "0:dataflow,262144:hardware-wait" > gpurun_out/r02_b9_dataflow_ab.log 2>&1
timeout 300 python tools/ab_libs.py run r01 current > gpurun_out/r02_b9_ab.log 2>&1
timeout 200 python tools/bringup.py gemm_timeline > gpurun_out/r02_b9_gemm_timeline.log 2>&1
COTR_TRACE_N=112 timeout 200 python tools/bringup.py forward_trace > gpurun_out/r02_b9_forward_trace.log 2>&1
tail -4 gpurun_out/r02_b9_pytest.log; cat gpurun_out/r02_b9_dataflow_ab.log gpurun_out/r02_b9_ab.log; grep -E "a_ln|res_ln" gpurun_out/r02_b9_gemm_timeline.log
