"""Import the REAL reference (ubc-vision/COTR) from /root/reference on CPU (TEST INFRASTRUCTURE).

Only usable in the authoring container: /root/reference does not exist on the
GPU box, so nothing that runs there may call this.  It is used by
oracle/make_golden.py to pin oracle/cotr_oracle.py and the host-side engine
against outputs of the reference itself.

Shims: stub modules for absent optional deps, explicit
PIL.Image import, torchvision resnet50 without the weight download, np.int,
and a CWD that holds ./out and ./tb_out.
"""
import argparse
import os
import sys
import types

REF_ROOT = os.environ.get("COTR_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "COTR", "models"))


def import_reference():
    """Returns the reference's top-level `COTR` package with the shims applied.  Changes CWD."""
    import numpy as np
    import PIL.Image  # noqa: F401  (the reference relies on matplotlib importing it)
    if not hasattr(np, "int"):
        np.int = int
    for name in ("matplotlib", "matplotlib.pyplot", "imageio", "tables", "vispy"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import torchvision
    if not getattr(torchvision.models.resnet50, "_cotr_shim", False):
        orig = torchvision.models.resnet50

        def resnet50_no_download(**kw):
            kw.pop("pretrained", None)
            return orig(weights=None, **kw)
        resnet50_no_download._cotr_shim = True
        torchvision.models.resnet50 = resnet50_no_download
    sys.dont_write_bytecode = True
    os.chdir(REF_ROOT)
    # our own repo also ships a `COTR` alias package: make sure the reference wins here
    for k in [k for k in sys.modules if k == "COTR" or k.startswith("COTR.")]:
        del sys.modules[k]
    if REF_ROOT in sys.path:
        sys.path.remove(REF_ROOT)
    sys.path.insert(0, REF_ROOT)
    # The reference's COTR directory has no __init__.py (namespace package), so a regular package of the same name
    # anywhere on sys.path - our alias package - would win.  Pin the package to the reference tree explicitly.
    pkg = types.ModuleType("COTR")
    pkg.__path__ = [os.path.join(REF_ROOT, "COTR")]
    sys.modules["COTR"] = pkg
    import COTR.models  # noqa: F401
    assert os.path.realpath(sys.modules["COTR.models"].__file__).startswith(os.path.realpath(REF_ROOT))
    return pkg


def default_opt():
    """The argparse namespace every reference demo builds (demo_single_pair.py:48-62)."""
    return argparse.Namespace(backbone="resnet50", hidden_dim=256, dilation=False, dropout=0.1, nheads=8,
                              layer="layer3", enc_layers=6, dec_layers=6, position_embedding="lin_sine",
                              dim_feedforward=1024)


def build_reference_model(state_dict_np):
    """Reference `build_model(opt)` on CPU with the given numpy weights loaded strictly."""
    import torch
    import_reference()
    from COTR.models import build_model
    model = build_model(default_opt())
    sd = {k: torch.from_numpy(v.copy()) for k, v in state_dict_np.items()}
    model.load_state_dict(sd, strict=True)
    return model.eval()
