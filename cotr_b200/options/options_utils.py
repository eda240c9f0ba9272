"""argparse helpers used by the demos (reference: COTR/options/options_utils.py:14-38)."""
import os  # noqa: F401  (the reference's demos rely on `from ...options_utils import *` re-exporting these)
import sys  # noqa: F401

from ..utils import utils
from ..global_configs import general_config, dataset_config  # noqa: F401


def str2bool(v: str) -> bool:
    return v.lower() in ('true', '1', 'yes', 'y', 't')


def print_opt(opt):
    lines = [name.rjust(25, ' ') + '  ' + str(getattr(opt, name)) for name in sorted(vars(opt))]
    utils.print_notification(lines, 'OPTIONS')
