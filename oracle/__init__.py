"""Test oracle for the COTR hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product path
(``cotr_b200``) never does: it fails loudly when the CUDA library is missing.

Parity status: the reference (ubc-vision/COTR @ 5c9363f) ships no tests, no
golden vectors and no known-answer fixtures for this path ("parity unpinned"
by the reference's own suite).  The restatement in
``oracle/cotr_oracle.py`` is therefore pinned against OUTPUTS OF THE REFERENCE
ITSELF, imported from ``/root/reference`` in the authoring container by
``oracle/make_golden.py`` (shim recipe: ``oracle/ref_shim.py``); the resulting
vectors are committed under ``tests/golden/`` together with that script.
"""
