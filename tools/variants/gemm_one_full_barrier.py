"""Experiment: one 'stage full' mbarrier per pipeline stage for both operands (128 cp.async arrivals + the TMA thread's
arrive.expect_tx) instead of two - the MMA thread pays one try_wait (~90 cycles) less per 64-wide K chunk."""
import os, sys
p = os.path.join(sys.argv[1], "gemm_tc.cu")
s = open(p).read()
def rep(a, b, cnt=1):
    global s
    assert s.count(a) == cnt, (a, s.count(a))
    s = s.replace(a, b)
rep("            mbar_init(&full_a[s], 128);\n            mbar_init(&full_b[s], 1);", "            mbar_init(&full_a[s], 129);\n            mbar_init(&full_b[s], 1);")
rep("                mbar_arrive_expect_tx(&full_b[s], 2u * C::kBPlane);", "                mbar_arrive_expect_tx(&full_a[s], 2u * C::kBPlane);")
rep("                tma_bulk_g2s(b_dst, src, C::kBPlane, &full_b[s]);", "                tma_bulk_g2s(b_dst, src, C::kBPlane, &full_a[s]);")
rep("                tma_bulk_g2s(b_dst + C::kBPlane, src + (size_t)npad * 128, C::kBPlane, &full_b[s]);", "                tma_bulk_g2s(b_dst + C::kBPlane, src + (size_t)npad * 128, C::kBPlane, &full_a[s]);")
rep("                mbar_wait(&full_b[s], ph);\n", "")
open(p, "w").write(s)
