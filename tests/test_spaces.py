"""`action_space` / `observation_space` of the compatibility view against the reference's own
declaration (yard.py:482-554).  CPU only: the spaces are pure functions of the shape."""
import numpy as np
import pytest

from student_mechanism_design_b200 import compat


def _norm(space):
    """(kind, attributes) of one of OUR spaces (gymnasium's or the stand-ins: same attribute names)"""
    name = type(space).__name__.lstrip("_")
    if name == "Box":
        return ("Box", float(np.min(space.low)), float(np.max(space.high)), tuple(space.shape), np.dtype(space.dtype))
    if name == "Discrete":
        return ("Discrete", int(space.n), int(getattr(space, "start", 0)))
    if name == "MultiDiscrete":
        return ("MultiDiscrete", tuple(int(x) for x in np.asarray(space.nvec).ravel()))
    if name == "MultiBinary":
        return ("MultiBinary", int(space.n))
    raise AssertionError(name)


def test_observation_space_known_answer():
    """yard.py:500-554 written out for the 15-node / 2-police smoke configuration"""
    sp = compat.reference_observation_space(15, 2, 20, 10)
    got = {k: _norm(sp[k]) for k in sp.keys()}
    assert got == {
        "adjacency_matrix": ("Box", 0.0, 1.0, (15, 15), np.dtype(np.int64)),
        "node_features": ("Box", 0.0, 1.0, (15, 3), np.dtype(np.int64)),
        "edge_index": ("Box", 0.0, 15.0, (2, 20), np.dtype(np.int32)),
        "edge_features": ("Box", 0.0, 5.0, (20,), np.dtype(np.int32)),
        "MrX_pos": ("Discrete", 15, 0),
        "Polices_pos": ("MultiDiscrete", (15, 15)),
        "Currency": ("MultiDiscrete", (11, 11)),
        "action_mask": ("MultiBinary", 15),
        "agent_position": ("Discrete", 15, 0),
        "agent_budget": ("Box", 0.0, 1000.0, (1,), np.dtype(np.float32)),
    }
    assert _norm(compat.reference_action_space(15)) == ("Discrete", 15, 0)
    ext = compat.reference_observation_space(15, 2, 20, 10, belief=True, reveal=True)
    assert _norm(ext["belief_map"]) == ("Box", 0.0, 1.0, (15,), np.dtype(np.float32))
    assert _norm(ext["MrX_revealed"]) == ("Discrete", 16, -1)


@pytest.mark.needs_reference
@pytest.mark.parametrize("N,E,P,money", [(15, 20, 2, 10), (30, 50, 5, 7)])
def test_observation_space_matches_the_unmodified_reference(N, E, P, money):
    """the reference env's own observation_space / action_space objects (built through the gymnasium stub, which records
    the constructor arguments) declare exactly what compat declares"""
    import ref_loader

    w = {k: 0.1 for k in ("Police_distance", "Police_group", "Police_position", "Police_time", "Mrx_closest", "Mrx_average",
                          "Mrx_position", "Mrx_time", "Police_coverage", "Police_proximity", "Police_overlap_penalty")}
    env = ref_loader.make_reference_env(P, money, w, N, E, seed=3)
    ref = env.observation_space("MrX").args[0]
    ours = compat.reference_observation_space(N, P, env.actual_num_edges, money)
    assert set(ref) == set(ours.keys())
    for k, r in ref.items():
        kind = type(r).__name__
        o = _norm(ours[k])
        assert o[0] == kind, k
        if kind == "Box":
            assert o[1:] == (float(r.kwargs["low"]), float(r.kwargs["high"]), tuple(r.kwargs["shape"]), np.dtype(r.kwargs["dtype"])), k
        elif kind == "Discrete":
            assert o[1] == int(r.n), k
        elif kind == "MultiDiscrete":
            assert o[1] == tuple(int(x) for x in r.args[0]), k
        elif kind == "MultiBinary":
            assert o[1] == int(r.args[0]), k
    assert compat.reference_action_space(N).n == env.action_space("MrX").n == N
