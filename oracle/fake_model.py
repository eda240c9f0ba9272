"""A cheap deterministic stand-in for the COTR network, used to pin the HOST-side engine (TEST INFRASTRUCTURE).

The zoom-in loop is discontinuous in the network output (integer crop corners, accept/reject thresholds, argsort), so
engine-level parity is tested by driving the reference engine and the rewrite with the SAME callable and demanding
identical results.  This callable mimics the network's contract (`model(img, queries)['pred_corrs']`, a parameter
that names the device) and depends on both the queries and the image content, so wrong crops, wrong normalisation or
a different call order all change the outcome.  Every call is logged (shapes + checksums).
"""
import numpy as np
import torch
from torch import nn


class FakeCOTR(nn.Module):
    def __init__(self):
        super().__init__()
        self.anchor = nn.Parameter(torch.zeros(1), requires_grad=False)
        self.calls = []

    @torch.no_grad()
    def forward(self, img, queries):
        img = img.double()
        q = queries.double()
        assert img.shape[-2:] == (256, 512)
        left = img[..., :256].mean(dim=(1, 2, 3))
        right = img[..., 256:].mean(dim=(1, 2, 3))
        tilt = 0.004 * torch.tanh(left - right)[:, None]                      # image dependent, small
        x, y = q[..., 0], q[..., 1]
        to_right = x < 0.5
        px = torch.where(to_right, x + 0.5, x - 0.5) + tilt * torch.sin(6.0 * y)
        py = y + tilt * torch.cos(5.0 * x) * 0.5
        pred = torch.stack([px, py], dim=-1).float()
        self.calls.append((tuple(img.shape), tuple(q.shape), float(img.sum()), float(q.sum()), float(pred.double().sum())))
        return {'pred_corrs': pred}


from cotr_b200.utils.synthetic import synthetic_image  # noqa: E402,F401  (kept under its old name for the tests)
