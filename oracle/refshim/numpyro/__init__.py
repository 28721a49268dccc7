"""`numpyro` stand-in: a recording context instead of effect handlers (TEST INFRASTRUCTURE).

`sample(name, dist)` returns the value preset for that site in the active `Recorder`;
`deterministic` and `factor` record their value and return it.
"""
from . import distributions  # noqa: F401

_active = None


class Recorder:
    def __init__(self, site_values):
        self.site_values = dict(site_values)
        self.sampled = {}
        self.deterministic = {}
        self.factors = {}
        self.priors = {}

    def __enter__(self):
        global _active
        self._prev = _active
        _active = self
        return self

    def __exit__(self, *exc):
        global _active
        _active = self._prev
        return False


def sample(name, d):
    v = _active.site_values[name]
    _active.sampled[name] = v
    _active.priors[name] = d
    return v


def deterministic(name, v):
    _active.deterministic[name] = v
    return v


def factor(name, v):
    _active.factors[name] = v
    return v


def set_host_device_count(n):
    return None
