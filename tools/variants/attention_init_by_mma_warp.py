"""Experiment for the open item of profiles/r01_forward_trace.md: the thread that initialises the mbarriers passes
griddepcontrol.wait ~5 us late.  Here the barriers are initialised by lane 0 of the MMA warp (warp 8), which never
executes griddepcontrol.wait."""
import os
import sys

p = os.path.join(sys.argv[1], "attention_tc.cu")
s = open(p).read()
a = "    if (threadIdx.x == 0) {\n        mbar_init(qk_full, kSoftmaxThreads);"
assert s.count(a) == 1
s = s.replace(a, "    if (threadIdx.x == kSoftmaxThreads) {\n        mbar_init(qk_full, kSoftmaxThreads);")
open(p, "w").write(s)
