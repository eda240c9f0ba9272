"""Engine-level timing on the GPU (BASELINE.json configs[2]-style): SparseEngine / FasterSparseEngine with forced
queries and 4 zoom levels on a synthetic pair, native model, device-side vs host-side crop/resize/normalise.

    python tools/engine_bench.py [n_queries]
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch

from cotr_b200.inference.sparse_engine import FasterSparseEngine, SparseEngine
from cotr_b200.models import build_model
from cotr_b200.utils.utils import fix_randomness
from cotr_b200.utils import synthetic as fixtures
from cotr_b200.utils.synthetic import synthetic_image

n_queries = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sd = fixtures.make_state_dict(0)
model = build_model(None)
model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
model = model.cuda().eval()
img_a = synthetic_image(51, 1024, 1024)
img_b = synthetic_image(52, 1024, 1024)
rs = np.random.RandomState(1)
queries = np.stack([rs.uniform(10, 1010, n_queries), rs.uniform(10, 1010, n_queries)], axis=1)
zooms = np.linspace(0.5, 0.0625, 4)


def run(engine_cls, on_device, **kw):
    fix_randomness(0)
    eng = engine_cls(model, 32, mode='tile', device_preprocess=on_device, **kw)
    sys.stdout = open(os.devnull, "w")
    try:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        corrs = eng.cotr_corr_multiscale(img_a, img_b, zooms, 1, max_corrs=n_queries, queries_a=queries.copy(), force=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        sys.stdout = sys.__stdout__
    return corrs, dt


run(SparseEngine, True)       # warm-up: graphs, workspace, coefficient tables
for cls, kw in ((SparseEngine, {}), (FasterSparseEngine, {})):
    base = None
    for on_device in (True, False):
        corrs, dt = run(cls, on_device, **kw)
        tag = "device pixels" if on_device else "host PIL pixels"
        same = "" if base is None else f"  identical to device path: {np.array_equal(base, corrs)}"
        base = corrs if base is None else base
        print(f"{cls.__name__:18s} {tag:16s}: {len(corrs)} correspondences from {n_queries} queries x 4 zoom levels in {dt:.2f} s "
              f"({n_queries * 4 / dt:.0f} query-steps/s){same}", flush=True)
