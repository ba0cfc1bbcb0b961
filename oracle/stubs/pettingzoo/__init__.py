"""Test-only stand-in for `pettingzoo` (reference: src/environment/base_env.py:2)."""


class ParallelEnv:
    pass
