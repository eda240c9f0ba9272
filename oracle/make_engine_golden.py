"""Generate tests/golden/engine_*.npz by running the REAL reference engines (authoring container only).

    python -m oracle.make_engine_golden

The reference's SparseEngine / FasterSparseEngine / cotr_flow / cotr_corr_base are imported unmodified through
oracle/ref_shim.py and driven by oracle/fake_model.py on seeded synthetic images; results and the model call log are
stored.  tests/test_engine_cpu.py replays the same scenarios through cotr_b200.inference and demands identical output.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import ref_shim  # noqa: E402
from oracle.fake_model import FakeCOTR, synthetic_image  # noqa: E402

ZOOMS = np.linspace(0.5, 0.0625, 4)


def _plain(o):
    """Numeric ndarray; task identifiers are None on the random-sampling path -> -1."""
    a = np.asarray(o)
    if a.dtype == object:
        a = np.array([-1 if v is None else v for v in a.ravel()], dtype=np.float64).reshape(a.shape)
    return a


def scenarios(SparseEngine, FasterSparseEngine, cotr_flow, cotr_corr_base, fix_randomness):
    """name -> callable returning (list of result arrays, FakeCOTR)."""
    img_a = synthetic_image(11, 300, 400)      # non-square: two overlapping tiles each in 'tile' mode
    img_b = synthetic_image(12, 360, 288)
    sq_a = synthetic_image(13, 320, 320)
    sq_b = synthetic_image(14, 256, 256)
    rs = np.random.RandomState(5)
    q_a = np.stack([rs.uniform(5, 395, 40), rs.uniform(5, 295, 40)], axis=1)
    q_sq = np.stack([rs.uniform(5, 315, 30), rs.uniform(5, 315, 30)], axis=1)

    def run(fn):
        def wrapped():
            fix_randomness(0)
            model = FakeCOTR()
            out = fn(model)
            out = list(out) if isinstance(out, (tuple, list)) else [out]
            return [_plain(o) for o in out], model
        return wrapped

    return {
        "flow_tile": run(lambda m: [x for i, x in enumerate(cotr_flow(m, img_a, img_b)) if i in (0, 1, 3, 4)]),
        "corr_base": run(lambda m: cotr_corr_base(m, img_a, img_b, q_a.copy())),
        "sparse_tile_random": run(lambda m: SparseEngine(m, 32, mode='tile').cotr_corr_multiscale_with_cycle_consistency(
            img_a, img_b, ZOOMS, 1, max_corrs=20, queries_a=None, return_idx=True, return_cycle_error=True)),
        "sparse_stretch_forced": run(lambda m: SparseEngine(m, 8, mode='stretching').cotr_corr_multiscale(
            img_a, img_b, ZOOMS, 3, max_corrs=40, queries_a=q_a.copy(), force=True, return_idx=True)),
        "sparse_square_queries": run(lambda m: SparseEngine(m, 16, mode='tile').cotr_corr_multiscale(
            sq_a, sq_b, ZOOMS, 2, max_corrs=25, queries_a=q_sq.copy(), force=False, return_idx=True)),
        "sparse_known_scale": run(lambda m: SparseEngine(m, 16, mode='tile').cotr_corr_multiscale(
            img_a, img_b, np.linspace(0.25, 0.0625, 2), 1, max_corrs=40, queries_a=q_a.copy(), force=True, areas=[1.0, 0.7])),
        "faster_tile_forced": run(lambda m: FasterSparseEngine(m, 4, mode='tile', max_load=16).cotr_corr_multiscale(
            img_a, img_b, ZOOMS, 1, max_corrs=40, queries_a=q_a.copy(), force=True, return_idx=True)),
        "faster_cycle": run(lambda m: FasterSparseEngine(m, 8, mode='tile').cotr_corr_multiscale_with_cycle_consistency(
            sq_a, sq_b, ZOOMS, 1, max_corrs=15, queries_a=None)),
    }


def call_log_array(model):
    rows = []
    for img_shape, q_shape, s_img, s_q, s_pred in model.calls:
        rows.append([img_shape[0], q_shape[1], s_img, s_q, s_pred])
    return np.array(rows, dtype=np.float64)


def main():
    ref_shim.import_reference()
    from COTR.inference.sparse_engine import SparseEngine, FasterSparseEngine
    from COTR.inference.inference_helper import cotr_flow, cotr_corr_base
    from COTR.utils.utils import fix_randomness
    out_dir = os.path.join(REPO, "tests", "golden")
    for name, fn in scenarios(SparseEngine, FasterSparseEngine, cotr_flow, cotr_corr_base, fix_randomness).items():
        results, model = fn()
        payload = {f"out{i}": r for i, r in enumerate(results)}
        if name == "flow_tile":     # dense maps are large: keep a strided sample + full-precision checksums
            payload = {}
            for i, r in enumerate(results):
                payload[f"sum{i}"] = np.array(r.astype(np.float64).sum())
                payload[f"abs{i}"] = np.array(np.abs(r.astype(np.float64)).sum())
                payload[f"sample{i}"] = r[::17, ::13].copy()
        payload["calls"] = call_log_array(model)
        np.savez_compressed(os.path.join(out_dir, f"engine_{name}.npz"), **payload)
        print(name, [r.shape for r in results], "model calls:", len(model.calls))


if __name__ == "__main__":
    main()
