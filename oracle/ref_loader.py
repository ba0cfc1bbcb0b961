"""Load the UNMODIFIED reference environment from /root/reference (oracle scaffolding).

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is imported by the product package
`student_mechanism_design_b200`; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may use it, and this particular
module is only usable in the build container, because `/root/reference` does not exist
on the GPU box.  It is used by `oracle/gen_golden.py` to produce `tests/golden/*.npz`
and by the `needs_reference` tests that pin the restatement in `sy_oracle.py`.

Recipe: stub `gymnasium` / `pettingzoo` (absent here),
replace `environment.visualization` (matplotlib is absent) by a no-op visualiser and
hand the env a logger whose methods do nothing.  The env sources themselves
(yard.py, reward_calculator.py, pathfinding.py, action_mask.py, graph_layout.py,
belief_module.py) are executed as they are.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SY_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "environment", "yard.py"))


class NullLogger:
    """No-op replacement for src/logger.py:Logger (flatters the reference: no I/O)."""

    def log(self, *args, **kwargs):
        pass

    def log_scalar(self, *args, **kwargs):
        pass

    def log_weights(self, *args, **kwargs):
        pass


class _NullVisualizer:
    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


VIS_OFF = {
    "visualize_game": False,
    "visualize_heatmap": False,
    "save_visualization": False,
    "save_dir": "unused",
}

_loaded = None


def load_reference():
    """Return a namespace with the reference's env classes/functions (imported once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    src = os.path.join(REFERENCE_ROOT, "src")
    for p in (src, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [_STUBS, src]
    vis = types.ModuleType("environment.visualization")
    vis.GameVisualizer = _NullVisualizer
    sys.modules["environment.visualization"] = vis
    # `environment/__init__.py` imports graph_generator and belief_module too; both import fine.
    from environment.yard import CustomEnvironment  # type: ignore
    from environment.action_mask import compute_action_mask  # type: ignore
    from environment.belief_module import ParticleBeliefTracker  # type: ignore
    from environment.graph_layout import ConnectedGraph  # type: ignore
    from environment.pathfinding import Pathfinder  # type: ignore

    _loaded = types.SimpleNamespace(
        CustomEnvironment=CustomEnvironment,
        compute_action_mask=compute_action_mask,
        ParticleBeliefTracker=ParticleBeliefTracker,
        ConnectedGraph=ConnectedGraph,
        Pathfinder=Pathfinder,
    )
    return _loaded


def make_reference_env(P, money, weights, N, E, seed=None):
    """Construct the reference env; `seed` seeds both global streams it consumes
    (python `random` and `numpy.random`: graph_layout.py:17,31,64,74, yard.py:113)."""
    import random

    import numpy as np

    ref = load_reference()
    if seed is not None:
        random.seed(seed)
        np.random.seed(seed)
    return ref.CustomEnvironment(P, money, weights, NullLogger(), 0, N, E, dict(VIS_OFF))
