"""`jax.scipy.special` stand-in (TEST INFRASTRUCTURE, see ../../README.md)."""
import torch as _t
from ..numpy import _TF


def logsumexp(x, axis=None):
    x = _TF(x)
    if axis is None:
        return _t.logsumexp(x.reshape(-1), dim=0)
    return _t.logsumexp(x, dim=axis)
