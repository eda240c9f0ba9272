"""Timing experiment: staging + S MMAs only; the softmax warps skip both passes and hand zero-work P chunks to the
MMA warp (P buffers are left as they are).  Separates the cost of the staging / S phase from the softmax / PV loop."""
import os
import sys

p = os.path.join(sys.argv[1], "attention_tc.cu")
s = open(p).read()
a = "        for (int c = 0; c < kTokens; c += 128) {"
assert s.count(a) == 1
s = s.replace(a, "        for (int c = 0; c < 0; c += 128) {")
a = "                    v[j] = fast_exp2(fmaf(__uint_as_float(r[h][j]), kLog2e, -mxs));"
assert s.count(a) == 1
s = s.replace(a, "                    v[j] = 0.001f;")
open(p, "w").write(s)
