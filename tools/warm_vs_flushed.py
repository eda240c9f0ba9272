import os, sys
sys.path.insert(0, os.getcwd())
import torch
from cotr_b200.models import build_model
from cotr_b200.utils import synthetic as fixtures
sd = fixtures.make_state_dict(0)
model = build_model(None)
model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
model = model.cuda().eval()
img, q = fixtures.make_inputs(1, 1, 1024)
img = torch.from_numpy(img).cuda(); q = torch.from_numpy(q).cuda()
for _ in range(5):
    model(img, q)
torch.cuda.synchronize()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for mode in ("warm", "flushed"):
    ts = []
    for _ in range(40):
        if mode == "flushed":
            flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); model(img, q); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(mode, "median %.4f ms" % ts[len(ts) // 2], "min %.4f" % ts[0], flush=True)
