"""CPU: the drop-in boundary - C-ABI symbols, state_dict schema, loud failure without a GPU."""
import argparse
import ctypes
import os
import re

import pytest
import torch

from oracle import fixtures

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "cotr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cotr_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    names = _declared_symbols()
    assert "cotr_create" in names and "cotr_forward" in names and len(names) >= 15
    handle = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/cotr_b200.h but not exported"


def test_ctypes_binding_covers_header(built_lib):
    from cotr_b200 import capi
    assert sorted(capi._PROTOTYPES) == _declared_symbols()
    assert capi.lib().cotr_version().startswith(b"cotr_b200")
    assert capi.lib().cotr_workspace_bytes(1, 1024) > 0


def _opt(**over):
    ns = argparse.Namespace(backbone="resnet50", hidden_dim=256, dilation=False, dropout=0.1, nheads=8, layer="layer3",
                            enc_layers=6, dec_layers=6, position_embedding="lin_sine", dim_feedforward=1024)
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


def test_state_dict_schema_matches_reference():
    from cotr_b200.models import build_model
    model = build_model(_opt())
    sd = model.state_dict()
    sch = dict(fixtures.schema())
    assert set(sd) == set(sch)
    for k, shape in sch.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    # FrozenBN entries are buffers, everything else a parameter (backbone.py:31-34)
    buffers = {n for n, _ in model.named_buffers()}
    assert all((".bn" in k or "downsample.1" in k) == (k in buffers) for k in sch)
    assert sum(p.numel() for p in model.parameters()) == 18387906
    # attributes the reference's callers read (train_cotr.py:49-55, cotr_model.py:20)
    assert model.transformer.d_model == 256
    for attr in ("corr_embed", "query_proj", "input_proj", "backbone"):
        assert hasattr(model, attr)


def test_strict_load_and_num_batches_tracked_is_dropped():
    from cotr_b200.models import build_model
    model = build_model(_opt())
    sd = {k: torch.from_numpy(v) for k, v in fixtures.make_state_dict(0).items()}
    sd["backbone.0.body.bn1.num_batches_tracked"] = torch.tensor(7)       # backbone.py:38-40
    model.load_state_dict(sd, strict=True)
    assert torch.equal(model.state_dict()["input_proj.bias"], sd["input_proj.bias"])


def test_unsupported_configuration_is_rejected():
    from cotr_b200.models import build_model
    with pytest.raises(NotImplementedError):
        build_model(_opt(layer="layer4", dim_feedforward=2048))


def test_forward_without_gpu_fails_loudly():
    from cotr_b200.models import build_model
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    model = build_model(_opt())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 3, 256, 512), torch.zeros(1, 4, 2))


def test_wrong_canvas_size_asserts_like_the_reference():
    from cotr_b200.models import build_model
    model = build_model(_opt())
    with pytest.raises(AssertionError):          # backbone.py:80
        model(torch.zeros(1, 3, 256, 256), torch.zeros(1, 4, 2))
