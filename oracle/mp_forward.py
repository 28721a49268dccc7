"""Test infrastructure (only tests/ may import this): a THIRD, independent restatement of the hot path's forward
pass in 40-digit arithmetic (mpmath, plain Python loops), used to check that the golden vectors minted from the
reference source in fp64 (tests/golden/) are right to fp64 rounding, and — by finite differences at a step of 1e-18,
which 40 digits make exact — that the reference's gradients are.  SURVEY.md section 8c asks for exactly this
cross-validation ("mpmath 50-digit forward on a tiny catalog") because the reference ships no golden vectors.

Follows /root/reference/src/scripts/intensity_models.py line by line (cited below) and utils.py:3-8; third-party
semantics restated from SURVEY.md appendix A2 (jnp.interp) and A8 (logsumexp).  Nothing here is shared with
oracle/bump_oracle.py, oracle/bump_cpu.cpp or the CUDA code.
"""
from mpmath import mp, mpf

mp.dps = 40

MBH_MIN = mpf(5)          # :13
MTR = mpf(20)             # :41
WIDTH = mpf("0.05")       # :45
NM = 256                  # :92
MIN_BH = mpf(3)           # :97
MIN_CO = mpf(1)           # :98
MREF = mpf(30)            # :129
QREF = mpf(1)             # :192
ZMAX = mpf(100)           # :220
NZ = 1024                 # :221
C_H100 = mpf("2.99792")   # :239
NEG_INF = mpf("-inf")


def linspace(a, b, n):
    return [a + (b - a) * mpf(k) / (n - 1) for k in range(n - 1)] + [b]


def logaddexp(a, b):
    if a == NEG_INF:
        return b
    if b == NEG_INF:
        return a
    m = max(a, b)
    return m + mp.log(mp.exp(a - m) + mp.exp(b - m))


def logsumexp(xs):
    xs = [x for x in xs if x != NEG_INF]
    if not xs:
        return NEG_INF
    m = max(xs)
    return m + mp.log(sum(mp.exp(x - m) for x in xs))


def interp(x, xp, fp):
    """jnp.interp: i = clip(searchsorted(xp, x, side='right'), 1, n-1); linear; end-clamped (appendix A2)."""
    n = len(xp)
    if x < xp[0]:
        return fp[0]
    if x > xp[n - 1]:
        return fp[n - 1]
    lo, hi = 0, n          # number of knots <= x
    while lo < hi:
        mid = (lo + hi) // 2
        if xp[mid] <= x:
            lo = mid + 1
        else:
            hi = mid
    i = min(max(lo, 1), n - 1)
    return fp[i - 1] + (x - xp[i - 1]) / (xp[i] - xp[i - 1]) * (fp[i] - fp[i - 1])


class Cosmology:
    """FlatwCDMCosmology (:212-273)."""

    def __init__(self, h, Om, w):
        self.h, self.Om, self.w = h, Om, w
        lz = linspace(mpf(0), mp.log1p(ZMAX), NZ)                      # :229-230
        self.z = [mp.expm1(v) for v in lz]
        dH = C_H100 / h                                                # :237-239
        inv_e = [1 / mp.sqrt(Om * (1 + z) ** 3 + (1 - Om) * (1 + z) ** (3 * (1 + w))) for z in self.z]   # :253-256
        cum = [mpf(0)]                                                 # utils.py:3-8
        for k in range(NZ - 1):
            cum.append(cum[-1] + (self.z[k + 1] - self.z[k]) * (inv_e[k] + inv_e[k + 1]) / 2)
        self.dc = [dH * c for c in cum]                                # :231
        self.dl = [d * (1 + z) for d, z in zip(self.dc, self.z)]       # :232
        self.ddl = [d + dH * (1 + z) * ie for d, z, ie in zip(self.dc, self.z, inv_e)]   # :233
        self.dvc = [4 * mp.pi * d * d * dH * ie for d, ie in zip(self.dc, inv_e)]       # :235


class MassFunction:
    """LogDNDMPISN (:56-111) + LogDNDM (:113-151)."""

    def __init__(self, a, b, c, mpisn, mbhmax, sigma, fpl, pisn_grid=None):
        self.c, self.mbhmax, self.sigma = c, mbhmax, sigma
        self.top = mbhmax + 7 * sigma
        self.mbh = linspace(MIN_BH, self.top, NM)                      # :102
        if pisn_grid is None:
            mcomax = 2 * mbhmax - mpisn                                # :27-30
            mco_top = mcomax + mp.sqrt(4 * mbhmax * (mbhmax - mpisn))
            mco = linspace(MIN_CO, mco_top, NM)                        # :103
            alpha = 1 / (4 * (mpisn - mbhmax))                         # :22
            mu = [m if m < mpisn else mbhmax + alpha * (m - mcomax) ** 2 for m in mco]   # :15-25
            ell = [(-a if m < MTR else -b) * mp.log(m / MTR) for m in mco]               # :32-43
            const = mp.log(mp.sqrt(2 * mp.pi)) + mp.log(sigma)
            ldm = [mp.log(mco[j + 1] - mco[j]) for j in range(NM - 1)]
            grid = []
            for mb in self.mbh:                                        # :105-107
                lw = [ell[j] - ((mb - mu[j]) / sigma) ** 2 / 2 - const for j in range(NM)]
                grid.append(logsumexp([mp.log(mpf("0.5")) + logaddexp(lw[j + 1], lw[j]) + ldm[j]
                                       for j in range(NM - 1)]))
            pisn_grid = grid
        self.grid = pisn_grid
        self.log_pl_norm = mp.log(fpl) + interp(mbhmax, self.mbh, self.grid)   # :136
        self.log_norm = mpf(0)                                                 # :130
        self.log_norm = -(self(MREF) + mp.log(MREF))                           # :138

    def __call__(self, m):                                                     # :140-151
        ld = interp(m, self.mbh, self.grid)
        if m <= self.mbh[0] or m >= self.mbh[-1]:
            ld = NEG_INF
        turn = mp.log(2) - mp.log1p(mp.exp(-(m - self.mbhmax) / (WIDTH * self.mbhmax)))   # :45-54
        ld = logaddexp(ld, -self.c * mp.log(m / self.mbhmax) + self.log_pl_norm + turn)
        if m < MBH_MIN:
            ld = NEG_INF
        return ld + self.log_norm


class Rate:
    """LogDNDV (:153-173)."""

    def __init__(self, lam, kappa, zp):
        self.lam, self.kappa, self.zp = lam, kappa, zp
        self.log_norm = mpf(0)
        self.log_norm = -self(mpf(0))

    def __call__(self, z):
        return self.lam * mp.log1p(z) - mp.log1p(((1 + z) / (1 + self.zp)) ** self.kappa) + self.log_norm


def forward(theta, data, pisn_grid=None):
    """theta = (h, Om, w, a, b, c, mpisn, mbhmax, sigma, fpl, beta, lam, kappa, zp); data = the nine model arguments
    (intensity_models.py:357).  Returns loglike, log_mu_sel, neff_sel, neff[nobs] (:378-394,401) and the PISN grid
    (reusable when only parameters outside (a, b, mpisn, mbhmax, sigma) change)."""
    h, Om, w, a, b, c, mpisn, mbhmax, sigma, fpl, beta, lam, kappa, zp = (mpf(float(t)) if not isinstance(t, mpf) else t
                                                                          for t in theta)
    m1s, qs, dls, pdraw, m1s_sel, qs_sel, dls_sel, pdraw_sel, ndraw = data
    cos = Cosmology(h, Om, w)
    mf = MassFunction(a, b, c, mpisn, mbhmax, sigma, fpl, pisn_grid)
    rate = Rate(lam, kappa, zp)

    def log_weight(m1d, q, dl, pd):                                    # :378-381 / :385-388
        m1d, q, dl, pd = mpf(float(m1d)), mpf(float(q)), mpf(float(dl)), mpf(float(pd))
        z = interp(dl, cos.dl, cos.z)                                  # :272-273
        m1 = m1d / (1 + z)
        m2 = q * m1
        dens = mf(m1) + mf(m2) + beta * mp.log((m1 + m2) / (MREF * (1 + QREF))) + mp.log(m1) + rate(z)   # :202-210
        dvc = interp(z, cos.z, cos.dvc)                                # :264-265
        ddl = interp(z, cos.z, cos.ddl)                                # :267-268
        if dens == NEG_INF or dvc <= 0:
            return NEG_INF
        return dens - 2 * mp.log1p(z) + mp.log(dvc) - mp.log(ddl) - mp.log(pd)

    loglike = mpf(0)
    neff = []
    for i in range(len(m1s)):
        lw = [log_weight(m1s[i][j], qs[i][j], dls[i][j], pdraw[i][j]) for j in range(len(m1s[i]))]
        l1, l2 = logsumexp(lw), logsumexp([2 * x for x in lw])
        loglike += l1 - mp.log(len(lw))                                # :382-383
        neff.append(mp.exp(2 * l1 - l2))                               # :401
    lw = [log_weight(m1s_sel[k], qs_sel[k], dls_sel[k], pdraw_sel[k]) for k in range(len(m1s_sel))]
    nd = mpf(float(ndraw))
    log_mu = logsumexp(lw) - mp.log(nd)                                # :389
    log_mu2 = logsumexp([2 * x for x in lw]) - 2 * mp.log(nd)          # :392
    log_s2 = log_mu2 + mp.log1p(-mp.exp(2 * log_mu - mp.log(nd) - log_mu2))   # :393
    neff_sel = mp.exp(2 * log_mu - log_s2)                             # :394
    return {"loglike": loglike, "log_mu_sel": log_mu, "neff_sel": neff_sel, "neff": neff, "pisn_grid": mf.grid}
