"""Minimal stand-in for `jax` (TEST INFRASTRUCTURE, see ../README.md). Only `jax.numpy`,
`jax.scipy.special` and `jax.random.PRNGKey` are reachable from the reference's hot-path module."""
from . import numpy  # noqa: F401
from . import scipy  # noqa: F401


class _Random:
    @staticmethod
    def PRNGKey(seed):
        return int(seed)


random = _Random()
