#!/bin/bash
# round-2 GPU batch 1: test-suite, 2-CTAs/SM A/B, compute-sanitizer passes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_b1_pytest.log
python tools/ab_libs.py run current two_ctas > gpurun_out/r02_b1_ab.log 2>&1
COTR_B200_LIB=$PWD/cotr_b200/lib/ab/two_ctas.so python tools/bringup.py model_tc > gpurun_out/r02_b1_two_ctas_model.log 2>&1
S=/usr/local/cuda/bin/compute-sanitizer
timeout 400 $S --tool memcheck python tools/bringup.py model_tc > gpurun_out/r02_san_memcheck_model.log 2>&1
timeout 400 $S --tool racecheck python tools/bringup.py gemm_tc_v0 > gpurun_out/r02_san_racecheck_gemm.log 2>&1
timeout 300 $S --tool racecheck python tools/bringup.py attn_tc > gpurun_out/r02_san_racecheck_attn.log 2>&1
timeout 400 $S --tool synccheck python tools/bringup.py gemm_tc_v0 > gpurun_out/r02_san_synccheck_gemm.log 2>&1
timeout 300 $S --tool synccheck python tools/bringup.py attn_tc > gpurun_out/r02_san_synccheck_attn.log 2>&1
tail -3 gpurun_out/r02_b1_pytest.log; cat gpurun_out/r02_b1_ab.log; tail -4 gpurun_out/r02_b1_two_ctas_model.log
for f in gpurun_out/r02_san_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|WORST|pred vs" $f | tail -3; done
