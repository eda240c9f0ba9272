"""CPU: host-side zoom-in engine parity.  tests/golden/engine_*.npz were produced by the REAL reference engines
(oracle/make_engine_golden.py) driven by oracle/fake_model.py; the rewrite in cotr_b200.inference, driven by the same
callable on the same seeded inputs, must make the same model calls and return the same correspondences."""
import os

import numpy as np
import pytest

from cotr_b200.inference.inference_helper import cotr_corr_base, cotr_flow, get_patch_centered_at, to_square_patches
from cotr_b200.inference.sparse_engine import FasterSparseEngine, SparseEngine
from cotr_b200.utils.utils import fix_randomness
from oracle.make_engine_golden import call_log_array, scenarios

SCENARIOS = scenarios(SparseEngine, FasterSparseEngine, cotr_flow, cotr_corr_base, fix_randomness)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_engine_matches_reference_engine(golden_dir, name, capsys):
    gold = np.load(os.path.join(golden_dir, f"engine_{name}.npz"))
    results, model = SCENARIOS[name]()
    calls = call_log_array(model)
    # same number of network invocations, same batch / query shapes, same image + query + prediction checksums
    assert calls.shape == gold["calls"].shape
    assert np.array_equal(calls[:, :2], gold["calls"][:, :2])
    np.testing.assert_allclose(calls[:, 2:], gold["calls"][:, 2:], rtol=1e-12, atol=1e-9)
    if name == "flow_tile":
        for i, r in enumerate(results):
            np.testing.assert_allclose(r.astype(np.float64).sum(), gold[f"sum{i}"], rtol=1e-12)
            np.testing.assert_allclose(np.abs(r.astype(np.float64)).sum(), gold[f"abs{i}"], rtol=1e-12)
            np.testing.assert_array_equal(r[::17, ::13], gold[f"sample{i}"])
    else:
        for i, r in enumerate(results):
            assert r.shape == gold[f"out{i}"].shape
            np.testing.assert_array_equal(r, gold[f"out{i}"])


def test_faster_engine_strands_ungrouped_tasks(golden_dir):
    """Reference quirk: tasks never grouped at an earlier zoom are silently dropped."""
    gold = np.load(os.path.join(golden_dir, "engine_faster_tile_forced.npz"))
    assert gold["out0"].shape[0] < 40


def test_rescue_stranded_flag_finishes_every_task(capsys):
    """Opt-in fix of that quirk: with rescue_stranded=True every forced query comes back, and the tasks the grouped
    phase did finish are answered exactly as without the flag (the extra work only touches the stranded ones)."""
    from oracle.fake_model import FakeCOTR, synthetic_image
    img_a = synthetic_image(1, 300, 400)
    img_b = synthetic_image(2, 360, 288)
    q = np.random.RandomState(5).uniform([5, 5], [395, 295], size=(40, 2))
    zooms = np.linspace(0.5, 0.0625, 4)
    out = {}
    for flag in (False, True):
        fix_randomness(0)
        eng = FasterSparseEngine(FakeCOTR(), 4, mode='tile', max_load=16, rescue_stranded=flag)
        out[flag] = eng.cotr_corr_multiscale(img_a, img_b, zooms, 1, max_corrs=40, queries_a=q.copy(), force=True, return_idx=True)
    (corr0, idx0), (corr1, idx1) = out[False], out[True]
    assert len(idx0) < 40 and len(idx1) == 40 and sorted(idx1.tolist()) == list(range(40))
    pos = {int(i): k for k, i in enumerate(idx1)}
    for k, i in enumerate(idx0):
        np.testing.assert_array_equal(corr0[k], corr1[pos[int(i)]])


def test_patch_geometry_edge_cases():
    img_shape = (300, 400, 3)
    # clamped at the top-left: shifted, not shrunk (inference_helper.py:88-98)
    p = get_patch_centered_at(None, [3.2, 2.9], scale=0.5, return_content=False, img_shape=img_shape)
    assert (p.x, p.y, p.w, p.h) == (0, 0, 150, 150) and (p.ow, p.oh) == (400, 300)
    # clamped at the bottom-right
    p = get_patch_centered_at(None, [399.0, 299.0], scale=0.5, return_content=False, img_shape=img_shape)
    assert (p.x + p.w, p.y + p.h) == (400, 300)
    # size is forced even, scale clipped to [0, 1]
    p = get_patch_centered_at(None, [200, 150], scale=7.0, return_content=False, img_shape=img_shape)
    assert (p.w, p.h) == (300, 300)
    p = get_patch_centered_at(None, [200, 150], scale=0.3033, return_content=False, img_shape=img_shape)
    assert p.w % 2 == 0 and p.w == int((300 * 0.3033 // 2) * 2)


def test_square_patch_tiling_rules():
    assert len(to_square_patches(np.zeros((64, 64, 3), np.uint8))) == 1
    with pytest.warns(UserWarning):
        tiles = to_square_patches(np.zeros((60, 100, 3), np.uint8))
    assert [(t.x, t.y, t.w, t.h) for t in tiles] == [(0, 0, 60, 60), (40, 0, 60, 60)]
    with pytest.raises(NotImplementedError):
        to_square_patches(np.zeros((60, 130, 3), np.uint8))


def test_nan_prediction_raises_like_the_reference():
    import torch
    from oracle.fake_model import FakeCOTR

    class NaNModel(FakeCOTR):
        def forward(self, img, queries):
            out = super().forward(img, queries)
            if queries.shape[1] == 1:
                out['pred_corrs'][0, 0, 0] = float('nan')
            return out

    img = np.zeros((128, 128, 3), np.uint8)
    eng = SparseEngine(NaNModel(), 4, mode='tile')
    with pytest.raises(ValueError, match='NaN in prediction'):       # sparse_engine.py:54-55
        eng.cotr_corr_multiscale(img, img, np.linspace(0.5, 0.25, 2), 1, max_corrs=4,
                                 queries_a=np.array([[30.0, 40.0], [60.0, 70.0]]), force=True)
