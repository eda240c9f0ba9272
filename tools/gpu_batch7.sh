#!/bin/bash
mkdir -p gpurun_out
python tools/ab_libs.py run current attn1w all1w onebar > gpurun_out/r02_b7_ab.log 2>&1
for v in attn1w all1w onebar; do echo "== $v"; COTR_B200_LIB=$PWD/cotr_b200/lib/ab/$v.so python tools/bringup.py model_tc 2>&1 | tail -2; done > gpurun_out/r02_b7_numerics.log
cat gpurun_out/r02_b7_ab.log gpurun_out/r02_b7_numerics.log
python -m pytest tests/test_engine_gpu.py -x -q -k "squad" 2>&1 | tail -40 > gpurun_out/r02_b7_squad.log
grep -E "^E " gpurun_out/r02_b7_squad.log | head -20
