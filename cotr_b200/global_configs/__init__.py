"""Site configuration (reference: COTR/global_configs/__init__.py + commons.json).

The reference asserts that ./out and ./tb_out exist relative to the CWD at import time; that is a training-time
convenience, so here the directories are only named, not required.
"""
general_config = {"out": "./out", "tb_out": "./tb_out"}
dataset_config = {}
