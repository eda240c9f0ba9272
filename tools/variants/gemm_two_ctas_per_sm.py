"""Experiment: let two narrow-tile GEMM CTAs share an SM, so that under programmatic dependent launch the whole next grid
is resident (prologue done, weights in flight, sitting in griddepcontrol.wait) before its predecessor ends - today a
CTA needs the SM to itself (166 KB shared memory, 512 TMEM columns, 164 registers), so most CTAs of the next grid
start only when a predecessor CTA exits (1-2.5 us of start skew per launch in profiles/r01_forward_trace.md).
For BN <= 32: 2 pipeline stages (80 KB + partial rows), 3 main accumulators (256 TMEM columns), 128 registers.
Check precision (tools/bringup.py model_tc) as well as time."""
import os
import sys

p = os.path.join(sys.argv[1], "gemm_tc.cu")
s = open(p).read()


def rep(a, b):
    global s
    assert s.count(a) == 1, a
    s = s.replace(a, b)


rep("    static constexpr int kStages = kStagesRaw > 4 ? 4 : kStagesRaw;",
    "    static constexpr int kStages = BN <= 32 ? 2 : (kStagesRaw > 4 ? 4 : kStagesRaw);")
rep("    static constexpr int kMain = BN >= 256 ? 1 : (BN >= 128 ? 2 : (BN == 64 ? 3 : 4));",
    "    static constexpr int kMain = BN >= 256 ? 1 : (BN >= 128 ? 2 : 3);")
rep("__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(",
    "__global__ void __launch_bounds__(kThreads, BN <= 32 ? 2 : 1) gemm_tc_kernel(")
open(p, "w").write(s)
