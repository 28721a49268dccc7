"""Run the UNMODIFIED reference hot path in the build container (TEST INFRASTRUCTURE ONLY).

The reference module `/root/reference/src/scripts/intensity_models.py` is imported as-is; its absent
third-party imports (jax, numpyro, astropy) are satisfied by the torch-float64 stand-ins in
`oracle/refshim/` (see its README), and `np.NINF` (removed in NumPy 2, used at
intensity_models.py:144,145,149) is restored as an attribute.  `torch.autograd` then differentiates
straight through the reference's own `pop_cosmo_model` (intensity_models.py:357-406).

This file is only usable where `/root/reference` exists (the build container).  It is used by
`tests/golden/make_golden.py` to mint fixtures and by `tests/test_oracle_vs_reference.py` (skipped when
the reference tree is absent).  The product never imports it.
"""
import os
import sys

import numpy as np
import torch

REFERENCE_SCRIPTS = os.environ.get("BUMP_REFERENCE_SCRIPTS", "/root/reference/src/scripts")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")

# The 14 numpyro *sample sites* that enter the factors (intensity_models.py:282-309), reference order.
SAMPLE_SITES = ("h", "Om", "w", "a", "b", "c", "mpisn", "dmbhmax", "sigma", "beta", "log_fpl",
                "lam", "dkappa", "zp")

_module = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_SCRIPTS, "intensity_models.py"))


def load_reference():
    """Import the reference's `intensity_models` with the shims on sys.path (idempotent)."""
    global _module
    if _module is not None:
        return _module
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_SCRIPTS}")
    if not hasattr(np, "NINF"):
        np.NINF = -np.inf  # NumPy<2 name used by the reference
    for p in (REFERENCE_SCRIPTS, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import intensity_models  # noqa: E402  (the reference module, unmodified)

    # numpy.ndarray * torch.Tensor is not defined; hand the diagnostics grids over as tensors
    # (module *data*, used only by the output-only curves at intensity_models.py:403-406).
    intensity_models.coords = {k: torch.as_tensor(v, dtype=torch.float64)
                               for k, v in intensity_models.coords.items()}
    _module = intensity_models
    return _module


def run_pop_cosmo_model(sites, data, R_unit=0.0, grad=True):
    """Execute the reference `pop_cosmo_model` at the given sample-site values.

    sites: dict name -> float for SAMPLE_SITES.  data: tuple of the 9 positional model arguments.
    Returns dict with the factors, deterministics and (if grad) d loglike / d site and
    d log_mu_sel / d site as float64 arrays in SAMPLE_SITES order.
    """
    im = load_reference()
    import numpyro  # the shim

    leaves = {k: torch.tensor(float(sites[k]), dtype=torch.float64, requires_grad=grad) for k in SAMPLE_SITES}
    vals = dict(leaves)
    vals["R_unit"] = torch.tensor(float(R_unit), dtype=torch.float64)
    with numpyro.Recorder(vals) as rec:
        im.pop_cosmo_model(*data)
    nobs = np.asarray(data[0]).shape[0]
    loglike = rec.factors["loglike"]
    selfactor = rec.factors["selfactor"]
    log_mu_sel = -selfactor / nobs
    out = {
        "loglike": float(loglike),
        "selfactor": float(selfactor),
        "log_mu_sel": float(log_mu_sel),
        "neff_sel": float(rec.deterministic["neff_sel"]),
        "neff": rec.deterministic["neff"].detach().numpy().copy(),
        "R": float(rec.deterministic["R"]),
        "mbhmax": float(rec.deterministic["mbhmax"]),
        "fpl": float(rec.deterministic["fpl"]),
        "kappa": float(rec.deterministic["kappa"]),
        "mdNdmdVdt_fixed_qz": rec.deterministic["mdNdmdVdt_fixed_qz"].detach().numpy().copy(),
        "dNdqdVdt_fixed_mz": rec.deterministic["dNdqdVdt_fixed_mz"].detach().numpy().copy(),
        "dNdVdt_fixed_mq": rec.deterministic["dNdVdt_fixed_mq"].detach().numpy().copy(),
        "hz": rec.deterministic["hz"].detach().numpy().copy(),
    }
    if grad:
        lv = [leaves[k] for k in SAMPLE_SITES]
        g1 = torch.autograd.grad(loglike, lv, retain_graph=True, allow_unused=True)
        g2 = torch.autograd.grad(log_mu_sel, lv, allow_unused=True)
        out["dloglike_dsite"] = np.array([0.0 if g is None else float(g) for g in g1])
        out["dlog_mu_sel_dsite"] = np.array([0.0 if g is None else float(g) for g in g2])
    return out


ALL_SITES = SAMPLE_SITES + ("R_unit",)   # the 15 sample sites of pop_cosmo_model in the package's order


def run_potential(u, data):
    """The potential energy numpyro's NUTS integrates for the reference model, U(u) = -[sum_sites log_prob(x_i(u_i)) +
    sum_sites log|dx_i/du_i| + loglike + selfactor], and dU/du, at the unconstrained point `u` (15 numbers, ALL_SITES
    order).  The distributions and their parameters are whatever the reference's own mass_parameters /
    redshift_parameters / cosmo_parameters / pop_cosmo_model pass to `numpyro.sample` (intensity_models.py:281-311,398):
    a first, discarded pass records them; densities and transforms are numpyro's, restated in refshim."""
    im = load_reference()
    import numpyro  # the shim

    # pass 1: which distribution does each site have?  (values are placeholders inside every support)
    probe = {k: torch.tensor(v, dtype=torch.float64) for k, v in dict(
        h=0.7, Om=0.3, w=-1.0, a=2.0, b=1.0, c=4.0, mpisn=35.0, dmbhmax=5.0, sigma=2.0, beta=0.0, log_fpl=-2.0,
        lam=2.7, dkappa=2.9, zp=1.9, R_unit=0.0).items()}
    with numpyro.Recorder(probe) as rec0:
        im.pop_cosmo_model(*data)
    dists = rec0.priors
    assert set(dists) == set(ALL_SITES), sorted(dists)
    # pass 2: x(u), the prior terms and the factors as one differentiable expression of u
    uu = torch.tensor(np.asarray(u, dtype=np.float64), dtype=torch.float64, requires_grad=True)
    vals, lp, lj = {}, 0.0, 0.0
    for i, name in enumerate(ALL_SITES):
        x, logj = dists[name].unconstrain_transform(uu[i])
        vals[name] = x
        lp = lp + dists[name].log_prob(x)
        lj = lj + logj
    with numpyro.Recorder(vals) as rec:
        im.pop_cosmo_model(*data)
    prior_part = -(lp + lj)
    U = prior_part - rec.factors["loglike"] - rec.factors["selfactor"]
    gU, = torch.autograd.grad(U, uu, retain_graph=True)
    gP, = torch.autograd.grad(prior_part, uu)
    return {"U": float(U), "grad": gU.numpy().copy(), "prior_U": float(prior_part), "prior_grad": gP.numpy().copy(),
            "x": np.array([float(vals[k]) for k in ALL_SITES]), "loglike": float(rec.factors["loglike"]),
            "selfactor": float(rec.factors["selfactor"]), "R": float(rec.deterministic["R"]),
            "distributions": {k: repr(dists[k]) for k in ALL_SITES}}


FIXED_SITES = ("a", "b", "c", "mpisn", "dmbhmax", "sigma", "beta", "log_fpl", "lam", "dkappa", "zp")


def run_pop_model(sites, data, R_unit=0.0):
    """Execute the reference's fixed-cosmology `pop_model` (intensity_models.py:313-355) unmodified; data are the
    source-frame arguments (m1s, qs, zs, pdraw, m1s_sel, qs_sel, zs_sel, pdraw_sel, Ndraw).  Also returns the
    dVdzdt table the model built (:323-325) so that the other implementations can be fed the same numbers."""
    im = load_reference()
    import numpyro  # the shim

    leaves = {k: torch.tensor(float(sites[k]), dtype=torch.float64, requires_grad=True) for k in FIXED_SITES}
    vals = dict(leaves)
    vals["R_unit"] = torch.tensor(float(R_unit), dtype=torch.float64)
    with numpyro.Recorder(vals) as rec:
        im.pop_model(*data)
    nobs = np.asarray(data[0]).shape[0]
    loglike, selfactor = rec.factors["loglike"], rec.factors["selfactor"]
    log_mu_sel = -selfactor / nobs
    lv = [leaves[k] for k in FIXED_SITES]
    g1 = torch.autograd.grad(loglike, lv, retain_graph=True, allow_unused=True)
    g2 = torch.autograd.grad(log_mu_sel, lv, allow_unused=True)
    zinterp = np.expm1(np.linspace(np.log1p(0), np.log1p(100), 1024))
    from astropy.cosmology import Planck18  # the shim
    tab = 4 * np.pi * Planck18.differential_comoving_volume(zinterp).value / (1 + zinterp)
    return {"loglike": float(loglike), "selfactor": float(selfactor), "log_mu_sel": float(log_mu_sel),
            "neff_sel": float(rec.deterministic["neff_sel"]), "neff": rec.deterministic["neff"].detach().numpy().copy(),
            "R": float(rec.deterministic["R"]),
            "dloglike_dsite": np.array([0.0 if g is None else float(g) for g in g1]),
            "dlog_mu_sel_dsite": np.array([0.0 if g is None else float(g) for g in g2]),
            "dvdzdt_interp": tab}


def reference_tables(sites):
    """Reference cosmology tables and PISN table at the given site values (no grad)."""
    im = load_reference()
    mbhmax = sites["mpisn"] + sites["dmbhmax"]
    cosmo = im.FlatwCDMCosmology(sites["h"], sites["Om"], sites["w"])
    pisn = im.LogDNDMPISN(sites["a"], sites["b"], sites["mpisn"], mbhmax, sites["sigma"])
    f = lambda t: t.detach().numpy().copy()  # noqa: E731
    return {
        "zinterp": f(cosmo.zinterp), "dcinterp": f(cosmo.dcinterp), "dlinterp": f(cosmo.dlinterp),
        "ddlinterp": f(cosmo.ddlinterp), "dvcinterp": f(cosmo.dvcinterp),
        "mbh_grid": f(pisn.mbh_grid), "log_dN_grid": f(pisn.log_dN_grid),
    }
