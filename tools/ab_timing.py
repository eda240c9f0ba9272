"""A/B timing of the forward (B=1, Q=1024; B=8) under the library's bring-up switches, in one process.
variant bits: 256 = no programmatic dependent launch, 512 = no split-K, 1024.. = tile-width thresholds
(launch_gemm_tc in csrc/gemm_tc.cu)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from cotr_b200 import capi
from cotr_b200.models import build_model
from cotr_b200.utils import synthetic as fixtures

sd = fixtures.make_state_dict(0)
model = build_model(None)
model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
model = model.cuda().eval()
nat = model.native()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def measure(B, Q, n=30):
    img, q = fixtures.make_inputs(1, B, Q)
    img = torch.from_numpy(img).cuda(); q = torch.from_numpy(q).cuda()
    for _ in range(4):
        model(img, q)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); model(img, q); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


VARIANTS = [(0, "default"), (512, "no split-K"), (256, "no pdl")]
if len(sys.argv) > 1:       # e.g. "0:default,1024:thr64=48,4096:thr128=48"
    VARIANTS = [(int(x.split(":")[0]), x.split(":")[1]) for x in sys.argv[1].split(",")]
for rep in range(2):
    for variant, name in VARIANTS:
        capi.lib().cotr_debug_set_variant(variant)
        nat.set_gemm_path(1); nat.set_gemm_path(0)          # drops the cached graphs
        print(f"  rep {rep} {name:18s}: B=1 Q=1024 {measure(1, 1024):.3f} ms | B=8 Q=1024 {measure(8, 1024, 10):.3f} ms | B=1 Q=16384 {measure(1, 16384, 10):.3f} ms", flush=True)
capi.lib().cotr_debug_set_variant(0)
