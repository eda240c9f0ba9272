"""Timing experiment ONLY (unsafe under PDL, see split16.cuh): activation / residual loads through ld.global.nc again."""
import os, sys
for f, pairs in (("gemm_tc.cu", [("o.res_hi[j] = __ldcg(", "o.res_hi[j] = __ldg("), ("o.res_lo[j] = __ldcg(", "o.res_lo[j] = __ldg(")]),
                 ("split16.cuh", [("const uint4 h = __ldcg(", "const uint4 h = __ldg("), ("const uint4 l = __ldcg(", "const uint4 l = __ldg(")])):
    p = os.path.join(sys.argv[1], f)
    s = open(p).read()
    for a, b in pairs:
        assert s.count(a) == 1, a
        s = s.replace(a, b)
    open(p, "w").write(s)
