"""`numpyro.distributions` stand-in (TEST INFRASTRUCTURE): the three families the reference's model uses, with the
log-densities and the unconstraining transforms numpyro applies to them (third-party semantics restated; numpyro is not
installable here).  Values are torch float64 tensors so that `torch.autograd` differentiates the potential energy.

  TruncatedNormal(loc, scale, *, low=None, high=None)
      numpyro: TwoSided / LeftTruncated / RightTruncatedDistribution(Normal(loc, scale), ...):
      log_prob(x) = Normal.log_prob(x) - log(Phi((high-loc)/scale) - Phi((low-loc)/scale)), missing bound -> 1 / 0;
      support: interval(low, high), greater_than(low), less_than(high)
  Normal(loc, scale)      support: real
  Uniform(low, high)      log_prob = -log(high - low); support: interval(low, high)

`biject_to(support)` (numpyro.distributions.transforms):
  real              identity                                  log|J| = 0
  greater_than(lb)  x = lb + exp(u)                           log|J| = u
  interval(lb, ub)  x = lb + (ub - lb) * sigmoid(u)           log|J| = log(ub - lb) + log sigmoid(u) + log sigmoid(-u)
"""
import math

import torch


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.tensor(float(x), dtype=torch.float64)


def _Phi(z):
    return 0.5 * torch.erfc(-_t(z) / math.sqrt(2.0))


class _D:
    def __init__(self, *args, **kw):
        self.args = args
        self.kw = kw

    def __repr__(self):
        return f"{type(self).__name__}{self.args}{self.kw}"

    low = None
    high = None

    # ---- biject_to(self.support): unconstrained u -> (x, log|dx/du|)
    def unconstrain_transform(self, u):
        u = _t(u)
        lo, hi = self.low, self.high
        if lo is not None and hi is not None:
            s = torch.sigmoid(u)
            return lo + (hi - lo) * s, math.log(hi - lo) + torch.nn.functional.logsigmoid(u) + \
                torch.nn.functional.logsigmoid(-u)
        if lo is not None:
            return lo + torch.exp(u), u
        if hi is not None:
            return hi - torch.exp(u), u
        return u, torch.zeros((), dtype=torch.float64)


class Normal(_D):
    def __init__(self, loc=0.0, scale=1.0):
        super().__init__(loc, scale)
        self.loc, self.scale = float(loc), float(scale)

    def log_prob(self, x):
        z = (_t(x) - self.loc) / self.scale
        return -0.5 * z * z - math.log(self.scale) - 0.5 * math.log(2.0 * math.pi)


class TruncatedNormal(_D):
    def __init__(self, loc=0.0, scale=1.0, *, low=None, high=None):
        super().__init__(loc, scale, low=low, high=high)
        self.loc, self.scale = float(loc), float(scale)
        self.low = None if low is None else float(low)
        self.high = None if high is None else float(high)

    def log_prob(self, x):
        base = Normal(self.loc, self.scale).log_prob(x)
        p_hi = _Phi((self.high - self.loc) / self.scale) if self.high is not None else _t(1.0)
        p_lo = _Phi((self.low - self.loc) / self.scale) if self.low is not None else _t(0.0)
        return base - torch.log(p_hi - p_lo)


class Uniform(_D):
    def __init__(self, low=0.0, high=1.0):
        super().__init__(low, high)
        self.low, self.high = float(low), float(high)

    def log_prob(self, x):
        return -torch.log(_t(self.high - self.low)) + 0.0 * _t(x)
