"""Timing experiment (wrong results by design): the attention kernel does nothing after its dependency wait - no staging,
no MMA, no softmax, no stores.  What 12 of these still cost is launch + prologue + wait, not the body."""
import os
import sys

p = os.path.join(sys.argv[1], "attention_tc.cu")
s = open(p).read()
a = "        if (dflow) mbar_wait(dep_ready, 0); else pdl_wait();      // prologue above overlaps the previous kernel\n"
assert s.count(a) == 1
s = s.replace(a, a + "        if (p.nq < 0) {   // never: the whole body is skipped\n")
a = "    } else {\n        // ================= MMA issuer"
assert s.count(a) == 1
s = s.replace(a, "        }\n    } else if (p.nq < 0) {\n        // ================= MMA issuer")
open(p, "w").write(s)
