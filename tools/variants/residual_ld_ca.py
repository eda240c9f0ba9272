"""Experiment: residual operand / elementwise activation loads with plain ld.global (L1-allocating, coherent after
griddepcontrol.wait) instead of ld.global.cg.  (The .nc flavour of round 1 is unsafe under PDL, see split16.cuh.)"""
import os, sys
for f, pairs in (("gemm_tc.cu", [("o.res_hi[j] = __ldcg(", "o.res_hi[j] = __ldca("), ("o.res_lo[j] = __ldcg(", "o.res_lo[j] = __ldca(")]),
                 ("split16.cuh", [("const uint4 h = __ldcg(", "const uint4 h = __ldca("), ("const uint4 l = __ldcg(", "const uint4 l = __ldca(")])):
    p = os.path.join(sys.argv[1], f)
    s = open(p).read()
    for a, b in pairs:
        assert s.count(a) == 1, a
        s = s.replace(a, b)
    open(p, "w").write(s)
