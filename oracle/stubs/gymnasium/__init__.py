"""Test-only stand-in for `gymnasium` (absent from this image).

Only what the reference env touches at import/run time is provided
(reference: src/environment/yard.py:3,34-36 and src/environment/graph_layout.py:2,53).
This is oracle scaffolding: it is never imported by the product package.
"""
from . import spaces  # noqa: F401
