"""Placeholder for the reference's COTR/utils/debug_utils.py (an IPython breakpoint helper the demos import)."""


def embed_breakpoint(*_a, **_k):   # pragma: no cover
    raise RuntimeError("interactive breakpoints are not part of the inference hot path")
