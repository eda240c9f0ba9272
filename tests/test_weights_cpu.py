"""CPU: checkpoint loading at the drop-in boundary - `safe_load_weights` (reference: COTR/utils/utils.py:164-193) on the
381-entry schema: strict load, the DataParallel 'module.' prefix in either direction, and the shape-matched partial load."""
import numpy as np
import torch

from cotr_b200.models import build_model
from cotr_b200.utils.utils import safe_load_weights
from oracle import fixtures


def _sd(seed=0):
    return {k: torch.from_numpy(v) for k, v in fixtures.make_state_dict(seed).items()}


def _same(model, sd):
    own = model.state_dict()
    return all(torch.equal(own[k], v) for k, v in sd.items())


def test_strict_load(capsys):
    model = build_model(None)
    sd = _sd(1)
    safe_load_weights(model, sd)
    assert 'weights safely loaded' in capsys.readouterr().out
    assert _same(model, sd)


def test_dataparallel_prefix_is_stripped(capsys):
    model = build_model(None)
    sd = _sd(2)
    safe_load_weights(model, {'module.' + k: v for k, v in sd.items()})
    assert _same(model, sd)


def test_partial_load_keeps_shape_matched_entries(capsys):
    model = build_model(None)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    sd = _sd(3)
    dropped = 'corr_embed.layers.2.weight'
    reshaped = 'input_proj.bias'
    partial = {k: v for k, v in sd.items() if k != dropped}
    partial[reshaped] = torch.zeros(7)                       # wrong shape: must be skipped, not crash
    partial['not.in.the.model'] = torch.zeros(3)
    safe_load_weights(model, partial)
    out = capsys.readouterr().out
    assert 'PARTIALLY' in out and dropped in out
    own = model.state_dict()
    assert torch.equal(own[dropped], before[dropped]) and torch.equal(own[reshaped], before[reshaped])
    key = 'transformer.encoder.layers.0.linear1.weight'
    assert torch.equal(own[key], sd[key])


def test_checkpoint_file_round_trip(tmp_path):
    """The reference's checkpoint file is a torch.save dict with 'model_state_dict' (cotr_trainer.py:76-81)."""
    src = build_model(None)
    src.load_state_dict(_sd(4))
    path = tmp_path / 'checkpoint.pth.tar'
    torch.save({'epoch': 1, 'iteration': 2, 'optim_state_dict': {}, 'model_state_dict': src.state_dict()}, path)
    dst = build_model(None)
    safe_load_weights(dst, torch.load(path, map_location='cpu')['model_state_dict'])
    assert _same(dst, src.state_dict())
    assert len(dst.state_dict()) == 381 and not any('num_batches_tracked' in k for k in dst.state_dict())
