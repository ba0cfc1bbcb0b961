"""Stand-ins for the parts of torchrl / tensordict the adapter touches (neither is installed in the build image):
`TensorDict` (nested tuple keys, batch-size check), `EnvBase` (reset / step / rollout / rand_action with the spec
attributes torchrl's base class requires), the spec classes (Composite / Bounded / Unbounded / Categorical / Binary:
shape, dtype, `is_in`, `rand`) and `check_env_specs`: every key a reset / step emits must be declared, with the declared
shape and dtype and inside its bounds, and every declared key must be emitted -- the checks of torchrl's own
`check_env_specs` that do not depend on torchrl internals."""
from types import SimpleNamespace

import torch


class TensorDict:
    def __init__(self, source=None, batch_size=None, device=None):
        self._d = {}
        self.batch_size = torch.Size(batch_size or [])
        self.device = device
        for k, v in (source or {}).items():
            self.set(k, v)

    @staticmethod
    def _key(k):
        return (k,) if isinstance(k, str) else tuple(k)

    def set(self, key, value):
        assert value.shape[: len(self.batch_size)] == self.batch_size, (key, value.shape, self.batch_size)
        self._d[self._key(key)] = value
        return self

    def get(self, key, default=...):
        k = self._key(key)
        if k in self._d:
            return self._d[k]
        if default is ...:
            raise KeyError(key)
        return default

    def keys(self):
        return list(self._d)

    def __getitem__(self, key):
        return self.get(key)


# ---------------------------------------------------------------------------------------------- specs
class _Leaf:
    def __init__(self, shape, dtype, device=None):
        self.shape, self.dtype, self.device = torch.Size(shape), dtype, device

    def _shape_dtype_ok(self, t):
        return tuple(t.shape) == tuple(self.shape) and t.dtype == self.dtype

    def is_in(self, t):
        return self._shape_dtype_ok(t)


class Unbounded(_Leaf):
    def __init__(self, shape, dtype=torch.float32, device=None):
        super().__init__(shape, dtype, device)

    def rand(self):
        return torch.randn(self.shape).to(self.dtype)


class Bounded(_Leaf):
    def __init__(self, low, high, shape, dtype=torch.float32, device=None):
        super().__init__(shape, dtype, device)
        self.low, self.high = low, high

    def is_in(self, t):
        return self._shape_dtype_ok(t) and bool((t >= self.low).all()) and bool((t <= self.high).all())

    def rand(self):
        return (torch.rand(self.shape) * (self.high - self.low) + self.low).to(self.dtype)


class Categorical(_Leaf):
    def __init__(self, n, shape, dtype=torch.int64, device=None):
        super().__init__(shape, dtype, device)
        self.n = n

    def is_in(self, t):
        return self._shape_dtype_ok(t) and bool((t >= 0).all()) and bool((t < self.n).all())

    def rand(self):
        return torch.randint(0, self.n, tuple(self.shape), dtype=self.dtype)


class Binary(_Leaf):
    def __init__(self, n, shape, dtype=torch.bool, device=None):
        super().__init__(shape, dtype, device)
        self.n = n
        assert self.shape[-1] == n, "torchrl: the last dimension of a Binary spec must equal n"

    def is_in(self, t):
        return self._shape_dtype_ok(t) and bool(((t == 0) | (t == 1)).all())

    def rand(self):
        return torch.rand(self.shape) < 0.5


class Composite:
    def __init__(self, source=None, shape=None, device=None):
        self._d = dict(source or {})
        self.shape, self.device = torch.Size(shape or []), device
        for k, v in self._d.items():  # torchrl: every entry shares the Composite's leading (batch) dims
            assert tuple(v.shape[: len(self.shape)]) == tuple(self.shape), (k, v.shape, self.shape)

    def leaves(self, prefix=()):
        for k, v in self._d.items():
            if isinstance(v, Composite):
                yield from v.leaves(prefix + (k,))
            else:
                yield prefix + (k,), v

    def keys(self, include_nested=True, leaves_only=True):
        return [k for k, _ in self.leaves()]

    def __getitem__(self, key):
        key = (key,) if isinstance(key, str) else tuple(key)
        v = self
        for k in key:
            v = v._d[k]
        return v

    def rand(self):
        return TensorDict({k: v.rand() for k, v in self.leaves()}, batch_size=self.shape)


SPECS = SimpleNamespace(Composite=Composite, Bounded=Bounded, Unbounded=Unbounded, Categorical=Categorical, Binary=Binary)


# ---------------------------------------------------------------------------------------------- env base
class EnvBase:
    """the surface of torchrl.envs.EnvBase the adapter and a collector use"""

    def __init__(self, device=None, batch_size=None):
        self.device, self.batch_size = device, batch_size
        self.observation_spec = self.action_spec = self.reward_spec = self.done_spec = None

    def reset(self, tensordict=None, **kwargs):
        return self._reset(tensordict, **kwargs)

    def step(self, tensordict):
        nxt = self._step(tensordict)
        tensordict.set("next_marker", torch.zeros(self.batch_size))
        return nxt

    def set_seed(self, seed):
        return self._set_seed(seed)

    def rand_action(self, tensordict=None):
        td = tensordict if tensordict is not None else TensorDict({}, batch_size=self.batch_size)
        for k, v in self.action_spec.leaves():
            td.set(k, v.rand())
        return td

    def rollout(self, max_steps, policy=None):
        """reset, then `max_steps` of policy (or random actions from the action spec) -> step; returns the step outputs"""
        td = self.reset()
        out = []
        for _ in range(max_steps):
            td = policy(td) if policy is not None else self.rand_action(td)
            td = self.step(td)
            out.append(td)
        return out


def check_env_specs(env, steps=3):
    """approximation of torchrl.envs.utils.check_env_specs: reset + random steps; every emitted key must be declared
    with the emitted shape / dtype / bounds, and every declared key must be emitted"""
    assert env.observation_spec is not None and env.action_spec is not None, "specs are not set"
    assert env.reward_spec is not None and env.done_spec is not None, "specs are not set"

    def check(td, specs, what):
        declared = {}
        for sp in specs:
            declared.update(dict(sp.leaves()))
        emitted = {k: td.get(k) for k in td.keys() if k != ("next_marker",)}
        missing = sorted(set(declared) - set(emitted))
        assert not missing, f"{what}: declared but not emitted: {missing}"
        extra = sorted(k for k in set(emitted) - set(declared) if k[-1] != "action")
        assert not extra, f"{what}: emitted but not declared: {extra}"
        for k, spec in declared.items():
            t = emitted[k]
            assert tuple(t.shape) == tuple(spec.shape), (what, k, tuple(t.shape), tuple(spec.shape))
            assert t.dtype == spec.dtype, (what, k, t.dtype, spec.dtype)
            assert spec.is_in(t), (what, k, "value outside the declared spec")

    td = env.reset()
    check(td, [env.observation_spec, env.done_spec], "reset")
    for _ in range(steps):
        td_in = env.rand_action(td)
        for k, spec in env.action_spec.leaves():
            assert spec.is_in(td_in.get(k)), ("action", k)
        td = env.step(td_in)
        check(td, [env.observation_spec, env.done_spec, env.reward_spec], "step")
    return True
