"""Timing experiment (wrong results by design): K and V^T are not staged (their cp.async loops are skipped, the barriers
still complete); Q staging, MMAs, softmax and stores run as usual on whatever shared memory holds."""
import os
import sys

p = os.path.join(sys.argv[1], "attention_tc.cu")
s = open(p).read()
a = "            for (int i = 0; i < kTokens / 128; ++i) {"
assert s.count(a) == 1
s = s.replace(a, "            for (int i = 0; i < 0; ++i) {")
a = "            for (int i = 0; i < 8; ++i) {\n                const int u = t + kSoftmaxThreads * i;"
assert s.count(a) == 1
s = s.replace(a, "            for (int i = 0; i < 0; ++i) {\n                const int u = t + kSoftmaxThreads * i;")
open(p, "w").write(s)
