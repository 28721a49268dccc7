"""Test helper: the host model interface (potential / evaluate) backed by the CPU oracle instead of the CUDA
library, so the priors + transforms + NUTS driver can be exercised on the real posterior without a GPU."""
import math

import numpy as np

from bumpcosmology_b200 import priors
from oracle import bump_oracle as bo


class OracleModel:
    def __init__(self, *data):
        self.data = data
        self.n_evals = 0

    def evaluate(self, x):
        x = np.asarray(x, dtype=np.float64)
        theta = priors.theta_from_sites(x)
        r = bo.evaluate(theta, self.data, grad=True)
        self.n_evals += 1
        nobs = r["nobs"]
        mu = math.exp(r["log_mu_sel"])
        return {"loglike": r["loglike"], "selfactor": -nobs * r["log_mu_sel"], "mbhmax": theta[7], "fpl": theta[9],
                "kappa": theta[12], "log_mu_sel": r["log_mu_sel"], "neff_sel": r["neff_sel"], "neff": r["neff"],
                "R": nobs / mu + math.sqrt(nobs) / mu * x[14],
                "dloglike_dsite": priors.grad_sites_from_theta(r["dloglike"], theta),
                "dselfactor_dsite": -nobs * priors.grad_sites_from_theta(r["dlog_mu_sel"], theta), "nobs": nobs}

    def potential(self, u):
        x, dx, lj, dlj = priors.constrain(u)
        lp, glp = priors.log_prior(x)
        ev = self.evaluate(x)
        logl = ev["loglike"] + ev["selfactor"]
        if not (math.isfinite(logl) and math.isfinite(lp)):
            return math.inf, np.zeros(priors.NSITES), ev
        g = glp.copy()
        g[:14] += ev["dloglike_dsite"] + ev["dselfactor_dsite"]
        return -(lp + lj + logl), -(g * dx + dlj), ev
