"""`jax.numpy` stand-in over torch float64 (TEST INFRASTRUCTURE, see ../README.md).

Semantics restated from the JAX API documentation (third-party, version unpinned by the reference's
environment.yml): `interp` = searchsorted(side='right') clipped to [1, n-1], linear inside, clamped to
fp[0]/fp[-1] outside, with a guard for dx ~ 0; differentiable with respect to x, xp and fp.
"""
import numpy as _np
import torch as _t

_F = _t.float64
pi = _np.pi
inf = _np.inf


def _T(x):
    if isinstance(x, _t.Tensor):
        return x if x.dtype == _F or not x.dtype.is_floating_point else x.to(_F)
    if hasattr(x, "to_numpy"):  # pandas Series (run_cosmo_fit.py:49 passes Series)
        x = x.to_numpy()
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], _t.Tensor):
        return _t.stack([_T(v) for v in x])
    a = _np.asarray(x)
    if a.dtype.kind in "iub":
        return _t.as_tensor(a)
    return _t.as_tensor(a, dtype=_F)


def _TF(x):
    x = _T(x)
    return x if x.dtype.is_floating_point else x.to(_F)


def array(x, dtype=None):
    return _T(x)


asarray = array


def zeros(n):
    return _t.zeros(n, dtype=_F)


def ones(n):
    return _t.ones(n, dtype=_F)


def where(c, a, b):
    c = _T(c)
    a = _TF(a)
    b = _TF(b)
    return _t.where(c, a, b)


def log(x):
    return _t.log(_TF(x))


def log1p(x):
    return _t.log1p(_TF(x))


def exp(x):
    return _t.exp(_TF(x))


def expm1(x):
    return _t.expm1(_TF(x))


def sqrt(x):
    return _t.sqrt(_TF(x))


def square(x):
    x = _TF(x)
    return x * x


def abs(x):  # noqa: A001
    return _t.abs(_TF(x))


def sum(x, axis=None):  # noqa: A001
    x = _TF(x)
    return x.sum() if axis is None else x.sum(dim=axis)


def cumsum(x, axis=None):
    x = _TF(x)
    return _t.cumsum(x, dim=0 if axis is None else axis)


def diff(x, axis=-1):
    return _t.diff(_TF(x), dim=axis)


def concatenate(xs, axis=0):
    return _t.cat([_TF(x) for x in xs], dim=axis)


def logaddexp(a, b):
    return _t.logaddexp(_TF(a), _TF(b))


def linspace(start, stop, num):
    # jnp.linspace: start*(1-s) + stop*s with s = i/(num-1), last point pinned to `stop`.
    start = _TF(start)
    stop = _TF(stop)
    s = _t.arange(num, dtype=_F) / (num - 1)
    out = start * (1 - s) + stop * s
    return _t.cat([out[:-1], stop.reshape(1)])


def interp(x, xp, fp):
    x = _TF(x)
    xp = _TF(xp)
    fp = _TF(fp)
    n = xp.shape[0]
    i = _t.clamp(_t.searchsorted(xp.detach(), x.detach().contiguous(), right=True), 1, n - 1)
    df = fp[i] - fp[i - 1]
    dx = xp[i] - xp[i - 1]
    delta = x - xp[i - 1]
    eps = float(_np.spacing(_np.finfo(_np.float64).eps))
    dx0 = _t.abs(dx) <= eps
    f = _t.where(dx0, fp[i - 1], fp[i - 1] + (delta / _t.where(dx0, _t.ones_like(dx), dx)) * df)
    f = _t.where(x < xp[0], fp[0], f)
    f = _t.where(x > xp[-1], fp[-1], f)
    return f
