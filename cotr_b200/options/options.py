"""Model flags of the hot path (reference: COTR/options/options.py:41-51).

`from COTR.options.options import *` in the reference demos also pulls in `sys`, `os`, `general_config`
(demo_single_pair.py:51,56,64), so they are re-exported here.
"""
import sys  # noqa: F401
import argparse  # noqa: F401
import json  # noqa: F401
import os  # noqa: F401

from .options_utils import str2bool
from . import options_utils  # noqa: F401
from ..global_configs import general_config, dataset_config  # noqa: F401


def set_COTR_arguments(parser):
    group = parser.add_argument_group('COTR model')
    group.add_argument('--backbone', type=str, default='resnet50')
    group.add_argument('--hidden_dim', type=int, default=256)
    group.add_argument('--dilation', type=str2bool, default=False)
    group.add_argument('--dropout', type=float, default=0.1)
    group.add_argument('--nheads', type=int, default=8)
    group.add_argument('--layer', type=str, default='layer3', help='which layer from resnet')
    group.add_argument('--enc_layers', type=int, default=6)
    group.add_argument('--dec_layers', type=int, default=6)
    group.add_argument('--position_embedding', type=str, default='lin_sine', help='sine wave type')
