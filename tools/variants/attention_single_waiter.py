"""Experiment: in the attention kernel only ONE thread executes griddepcontrol.wait; the other warps learn about the
completed dependency through an mbarrier (hypothesis: warps parked in griddepcontrol.wait are released one after the
other, ~0.7-1.2 us apart - 8 waiting warps would explain the ~6 us every attention launch loses, profiles/r01_forward_trace.md)."""
import os, sys
p = os.path.join(sys.argv[1], "attention_tc.cu")
s = open(p).read()
def rep(a, b):
    global s
    assert s.count(a) == 1, a
    s = s.replace(a, b)
rep("    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);", "    uint64_t* dep_ready = bars + 8;\n    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);")
rep("        mbar_init(&p_empty[1], 1);\n", "        mbar_init(&p_empty[1], 1);\n        mbar_init(dep_ready, 1);\n")
rep("        if (t == 0) pdl_launch_dependents();             // the next kernel may start its prologue on idle SMs\n        pdl_wait();                                      // prologue above overlaps the previous kernel\n",
    "        mbar_wait(dep_ready, 0);                         // the MMA warp's lane 0 is the only thread in griddepcontrol.wait\n")
rep("        if (lane == 0) {\n            constexpr uint32_t idesc_s = make_idesc_f16_f32(128, 256);",
    "        if (lane == 0) {\n            pdl_launch_dependents();\n            pdl_wait();\n            mbar_arrive(dep_ready);\n            constexpr uint32_t idesc_s = make_idesc_f16_f32(128, 256);")
open(p, "w").write(s)
