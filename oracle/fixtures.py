"""Seeded synthetic weights and inputs for parity tests (TEST INFRASTRUCTURE).

No trained checkpoint is available offline, so parity
is checked on deterministic synthetic weights.  Everything is drawn from
``numpy.random.RandomState`` (bit-stable across numpy versions and machines),
so the authoring container and the GPU box build identical tensors.

Conditioning follows SURVEY.md appendix E.3: xavier-random transformer weights
make the output nearly constant over queries (a 1e-3 check would be vacuous),
over-peaked attention makes the reference's own fp32 differ from fp64.  The
default gains put the fixture in the usable window; tests assert both
``std_over_queries(pred) >> 1e-3`` and ``|fp32 - fp64| << 1e-3``.
"""
from cotr_b200.utils.synthetic import D, FF, make_inputs, make_state_dict, schema  # noqa: F401

# The generators live in the product package (cotr_b200/utils/synthetic.py) so that bench.py's native arm and the tools
# import nothing under oracle/; this module keeps the names the tests and the golden scripts use.
