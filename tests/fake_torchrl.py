"""Minimal stand-ins for torchrl.envs.EnvBase and tensordict.TensorDict (neither is installed in the build
image): just enough surface for student_mechanism_design_b200.torchrl_env.make_env_class."""
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


class EnvBase:
    def __init__(self, device=None, batch_size=None):
        self.device, self.batch_size = device, batch_size

    def reset(self, tensordict=None, **kwargs):
        return self._reset(tensordict, **kwargs)

    def step(self, tensordict):
        nxt = self._step(tensordict)
        tensordict.set("next_marker", torch.zeros(self.batch_size))
        return nxt

    def set_seed(self, seed):
        return self._set_seed(seed)
