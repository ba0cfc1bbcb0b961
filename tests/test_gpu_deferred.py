"""Software-pipelined stepping (sy_step_deferred / sy_flush_observations / sy_step_lagged_kernel) against the oracle.

A deferred step runs the dynamics of yard.py:144-269 exactly like sy_step, but leaves action_mask / node_features /
belief_map of the new state pending; the next deferred step writes them in the same launch as its own dynamics (the
lagged kernel: observation roles of step k and dynamics warps of step k + 1 in one CTA per 32-env tile).  Checked here:
state, rewards (float64 bits), flags, statistics after EVERY deferred call; the dense tensors describe the state before
the call (one step behind) and the current state after a flush; plain steps, partial resets and the two-launch
fall-back mixed in; BASELINE config 3 at its benchmarked size; ragged batches with mixed-graph tiles; no belief map;
15 police (MAXA = 16); the fp32 reward mode."""
import numpy as np
import pytest

import sy_oracle as so
import sy_oracle_c as oc
from test_gpu_fullsize import BELIEF_TOL, _compare, _same, _threads, torch_cuda  # noqa: F401

pytestmark = pytest.mark.gpu


def _pair(N, E, P, money, B, seed, *, toll, belief, reveal, graphs=1, auto_reset=True, reward_mode="fp64", lagged=True,
          max_timestep=None):
    import student_mechanism_design_b200 as pkg

    pool = pkg.generate_graph_pool(graphs, N, E, seed=0)
    kw = {} if max_timestep is None else {"max_timestep": max_timestep}
    env = pkg.BatchedScotlandYardEnv(B, P, money, graphs=pool, seed=seed, auto_reset=auto_reset, tolls=toll, belief=belief,
                                     belief_ce=belief, reveal_interval=reveal, keep_reward64=True, reward_mode=reward_mode, **kw)
    env.set_option("lagged_kernel", "on" if lagged else "off")
    cfg = so.OracleConfig(num_police=P, agent_money=money, toll=toll, belief=belief, reveal_interval=reveal,
                          reward_mode=reward_mode, **kw)
    gid = (np.arange(B) // 32) % graphs
    ob = oc.CBatch(cfg, pool, B, seed=seed, auto_reset=auto_reset, threads=_threads(), graph_id=gid.astype(np.int32))
    return env, ob


def _dense_snapshot(ob, belief):
    return ob._mask.copy(), ob._nf.copy(), (ob._belief.copy() if belief else None)


def _dense_equal(env, snap, belief, tag):
    mask, nf, bel = snap
    _same(env.action_mask, mask, (tag, "action_mask"))
    _same(env.node_features, nf, (tag, "node_features"))
    if belief:
        err = np.abs(env.belief_map.cpu().numpy().astype(np.float64) - bel).max()
        assert err <= BELIEF_TOL, (tag, "belief_map", err)


def _deferred_rollout(env, ob, steps, belief, *, dense_every=1, plain_every=0, auto_reset=True):
    env.reset()
    _compare(env, ob, None, belief, "reset")
    assert not env.observations_pending
    for s in range(steps):
        acts = env.sample_actions(step_counter=s)
        a_h = acts.cpu().numpy()
        assert np.array_equal(a_h, ob.sample_actions(s)), ("sampler", s)
        before = _dense_snapshot(ob, belief)
        want = ob.step(a_h)
        if plain_every and s % plain_every == plain_every - 1:
            env.step(acts)  # a plain step in the middle of a deferred run: flushes, steps, observes
            assert not env.observations_pending
            _compare(env, ob, want, belief, ("plain step", s))
            continue
        env.step_deferred(acts)
        assert env.observations_pending
        _compare(env, ob, want, belief and False, ("deferred step", s), dense=False)
        if s % dense_every == 0:  # the dense tensors trail the state by exactly one step
            _dense_equal(env, before, belief, ("one step behind", s))
    env.flush_observations()
    assert not env.observations_pending
    _compare(env, ob, None, belief, "after flush")
    env.flush_observations()  # idempotent
    _compare(env, ob, None, belief, "after second flush")
    st = env.stats()
    if auto_reset:  # (without it finished envs freeze: they neither step nor start new episodes)
        assert st["env_steps"] == env.num_envs * steps and st["episodes"] == int(np.asarray(ob._episode).sum())


def test_config3_full_size_deferred(torch_cuda):
    """BASELINE config 3 at its benchmarked size, 25 deferred steps: every tile of the 2 048-CTA lagged grid."""
    env, ob = _pair(200, 400, 6, 20, 65536, 1, toll=1, belief=True, reveal=5)
    _deferred_rollout(env, ob, 25, True, dense_every=4)
    env.close()


@pytest.mark.parametrize("lagged", [True, False])
def test_ragged_pool_deferred_with_plain_steps(torch_cuda, lagged):
    """B not a multiple of 32, a pool of 5 graphs whose blocks make the last tiles mixed, plain steps interleaved; also
    with the lagged kernel switched off (pending observations and dynamics as two launches)."""
    env, ob = _pair(200, 400, 6, 20, 2048 + 19, 5, toll=1, belief=True, reveal=5, graphs=5, lagged=lagged)
    _deferred_rollout(env, ob, 30, True, plain_every=7)
    env.close()


@pytest.mark.parametrize("shape", ["c2", "p15", "fp32", "tiny", "no_reset"])
def test_deferred_shapes(torch_cuda, shape):
    """No belief map (the lagged kernel without belief warps), 15 police (MAXA = 16, the largest dynamics tile), the
    fp32 reward arithmetic, a batch smaller than one tile, and frozen envs (auto_reset off, short episodes)."""
    if shape == "c2":
        env, ob = _pair(50, 110, 3, 10, 1024, 2, toll=0, belief=False, reveal=5)
        _deferred_rollout(env, ob, 60, False)
    elif shape == "p15":
        env, ob = _pair(120, 300, 15, 12, 777, 3, toll=1, belief=True, reveal=4)
        _deferred_rollout(env, ob, 20, True)
    elif shape == "fp32":
        env, ob = _pair(60, 130, 4, 10, 1500, 4, toll=0, belief=True, reveal=3, reward_mode="fp32")
        _deferred_rollout(env, ob, 25, True)
    elif shape == "tiny":
        env, ob = _pair(15, 20, 2, 10, 5, 6, toll=0, belief=True, reveal=0)
        _deferred_rollout(env, ob, 15, True)
    else:
        env, ob = _pair(30, 60, 2, 6, 300, 8, toll=1, belief=True, reveal=2, auto_reset=False, max_timestep=6)
        _deferred_rollout(env, ob, 14, True, auto_reset=False)
    env.close()


def test_statistics_of_deferred_steps_equal_plain_steps(torch_cuda):
    """Every statistic -- incl. the belief cross-entropy at reveal steps, which the OBSERVATION kernel accumulates -- is the
    same whether the steps run plainly, deferred with a final flush (the flush has no SyOut: it inherits the statistics
    flag of the step it completes), or deferred with a partial reset in between."""
    envs = []
    for _ in range(3):
        e, _ob = _pair(40, 80, 3, 8, 500, 17, toll=1, belief=True, reveal=3)
        e.reset()
        envs.append(e)
    plain, deferred, mixed = envs
    for s in range(20):
        acts = plain.sample_actions(step_counter=s)
        plain.step(acts)
        deferred.step_deferred(acts)
        mixed.step_deferred(acts) if s % 5 else mixed.step(acts)
    deferred.flush_observations()
    mixed.flush_observations()
    want = plain.stats()
    assert want["reveals"] > 0
    for e in (deferred, mixed):
        assert e.stats() == want
        assert e.belief_map.cpu().numpy().tobytes() == plain.belief_map.cpu().numpy().tobytes()
    for e in envs:
        e.close()


def test_partial_reset_with_pending_observations(torch_cuda):
    """reset(reset_mask) while observations are pending: the pending belief propagation of the envs that are NOT reset
    must not be lost (sy_reset flushes first).  Python oracle (it has the partial reset), frozen envs without auto-reset."""
    import os

    from conftest import GOLDEN
    from test_gpu_parity import _compare_out, _compare_state, _make_pair

    import student_mechanism_design_b200 as pkg

    t = np.load(os.path.join(GOLDEN, "tables.npz"))
    tables = (t["exp_neg"], t["coverage"])
    c = dict(N=20, E=34, P=3, money=7, G=3, B=150, kw=dict(belief=True, reveal_interval=2, tolls=1), mode="fp64")
    env, ob = _make_pair(pkg, c, tables, auto_reset=False, max_timestep=9)
    env.reset()
    N, A = c["N"], c["P"] + 1
    s = 0
    for rnd in range(3):
        for _ in range(7):
            acts = env.sample_actions(step_counter=s)
            want = ob.step(acts.cpu().numpy())
            env.step_deferred(acts)
            _compare_out(env, want, c, ("out", s))
            s += 1
        assert env.observations_pending
        m = np.asarray(ob.done)
        assert m.any() and not m.all()
        env.reset(reset_mask=m)
        for b in np.nonzero(m)[0]:
            ob.episode[b] += 1
            ob.envs[b].reset(ob.graphs[ob.graph_id[b]], so.philox_start_positions(ob.seed, b, ob.episode[b], N, A))
            ob.done[b] = False
        assert not env.observations_pending
        _compare_state(env, ob, c, ("after partial reset", rnd))
    env.close()


def test_rollout_random_is_pipelined_and_flushed(torch_cuda):
    """sy_rollout_random steps deferred and flushes once at the end: state AND observations equal the oracle's after the
    call, and the library launched one kernel per step (+ the first sampler and the flush)."""
    import student_mechanism_design_b200 as pkg

    B, T = 16384 + 5, 9
    env, ob = _pair(200, 400, 6, 20, B, 13, toll=1, belief=True, reveal=5)  # (lagged_kernel on)
    env.reset()
    torch_cuda.cuda.synchronize()
    lib = pkg.load_library()
    n0 = lib.sy_launch_count()
    env.rollout_random(T, step_counter=0)
    torch_cuda.cuda.synchronize()
    launches = lib.sy_launch_count() - n0
    want = None
    for s in range(T):
        want = ob.step(ob.sample_actions(s))
    _compare(env, ob, want, True, "after rollout_random")
    assert launches == T + 2, launches  # sampler, dynamics, (T - 1) lagged launches, flush
    env.close()


@pytest.mark.parametrize("shape,one_launch", [("c2", True), ("c2", False), ("c3_one_cta_per_sm", True), ("c3_one_wave", False), ("p15", True)])
def test_small_batch_rollout_graph_is_pipelined_by_default(torch_cuda, shape, one_launch):
    """Batches of at most one wave of the lagged kernel: capture_rollout / sy_rollout_random_dev run pipelined by default
    (lagged_kernel = auto) -- the whole segment as ONE launch (sy_rollout_lagged_kernel) up to one CTA per SM, one lagged
    launch per step above that or with rollout_kernel = off -- and the replayed CUDA graph leaves state AND observations
    equal to the oracle's (BASELINE config 2 as benchmarked, c3-shaped batches of 148 and 296 tiles minus a ragged tail,
    15 police on a graph pool)."""
    torch = torch_cuda
    if shape == "p15":
        env, ob = _pair(120, 230, 15, 12, 777, 3, toll=1, belief=True, reveal=4, graphs=3)
        belief = True
    elif shape == "c2":
        env, ob = _pair(50, 110, 3, 10, 1024, 2, toll=0, belief=False, reveal=5)
        belief = False
    else:  # 148 tiles: still ONE launch per rollout; 296 tiles: one lagged launch per step even with the option on
        tiles = 148 if shape == "c3_one_cta_per_sm" else 296
        env, ob = _pair(200, 400, 6, 20, tiles * 32 - 7, 3, toll=1, belief=True, reveal=5)
        belief = True
    env.set_option("lagged_kernel", "auto")
    env.set_option("rollout_kernel", "off" if shape == "c2" and not one_launch else "on")
    env.reset()
    seg, replays = 6, 5
    lib = __import__("student_mechanism_design_b200").load_library()
    graph, counter = env.capture_rollout(seg)  # warm-up segment, then capture
    torch.cuda.synchronize()
    n0 = lib.sy_launch_count()
    env.rollout_random(seg, step_counter=seg)  # the same call outside a graph, for the launch count
    torch.cuda.synchronize()
    # one launch: sampler + the rollout kernel; per step: sampler, dynamics, (seg - 1) lagged launches, flush
    assert lib.sy_launch_count() - n0 == (2 if one_launch else seg + 2)
    counter.fill_(2 * seg)
    for _ in range(replays):
        graph.replay()
    torch.cuda.synchronize()
    total = seg * (2 + replays)
    want = None
    for s in range(total):
        want = ob.step(ob.sample_actions(s))
    _compare(env, ob, want, belief, "after pipelined graph replays")
    assert not env.observations_pending
    env.close()
