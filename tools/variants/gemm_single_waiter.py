"""Experiment: one thread per GEMM CTA in griddepcontrol.wait, everybody else behind an mbarrier (see attention_single_waiter.py)."""
import os, sys
p = os.path.join(sys.argv[1], "gemm_tc.cu")
s = open(p).read()
def rep(a, b, cnt=1):
    global s
    assert s.count(a) == cnt, (a, s.count(a))
    s = s.replace(a, b)
rep("    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * C::kStages + 3);",
    "    uint64_t* dep_ready = bars + 3 * C::kStages + 3;\n    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * C::kStages + 4);")
rep("        mbar_init(vec_full, 64);\n", "        mbar_init(vec_full, 64);\n        mbar_init(dep_ready, 1);\n")
rep("        if (threadIdx.x == 0) pdl_launch_dependents();\n        pdl_wait();\n        if (threadIdx.x == 0) COTR_TS(2);",
    "        if (threadIdx.x == 0) { pdl_launch_dependents(); pdl_wait(); mbar_arrive(dep_ready); }\n        mbar_wait(dep_ready, 0);\n        if (threadIdx.x == 0) COTR_TS(2);")
rep("        pdl_wait();\n        auto stage_stats = [&]", "        mbar_wait(dep_ready, 0);\n        auto stage_stats = [&]")
rep("        if (warp >= 4) pdl_wait();           // residual / add operands come from the previous kernels",
    "        if (warp >= 4) mbar_wait(dep_ready, 0);           // residual / add operands come from the previous kernels")
open(p, "w").write(s)
