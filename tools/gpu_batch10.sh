#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_b10_pytest.log
timeout 300 python tools/ab_timing.py "0:default,524288:deferred-ln,786432:deferred-ln+dataflow" > gpurun_out/r02_b10_schedules_ab.log 2>&1
timeout 300 python tools/ab_libs.py run r01 current > gpurun_out/r02_b10_ab.log 2>&1
tail -4 gpurun_out/r02_b10_pytest.log; cat gpurun_out/r02_b10_schedules_ab.log gpurun_out/r02_b10_ab.log
