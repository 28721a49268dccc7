"""The potential energy NUTS integrates, against the golden minted from the reference's own prior declarations
(tests/golden/make_golden_potential.py: unmodified mass_parameters / redshift_parameters / cosmo_parameters /
pop_cosmo_model, /root/reference/src/scripts/intensity_models.py:281-311,357-401, with numpyro's densities and
biject_to transforms restated in oracle/refshim).

CPU: the prior + Jacobian part of priors.py (Python driver) and of csrc/bump_nuts.cpp (C++ driver).
GPU: the full potential and its 15-dimensional gradient from PopCosmoModel.potential and from the library's
bump_nuts_potential.  Tolerance 1e-10 relative (gradients: floor 1e-10 x gradient scale)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "potential_small.npz"))


def _close(a, b, floor=1.0, rtol=1e-10):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), floor)))


def test_golden_covers_all_fifteen_sites_in_package_order():
    from bumpcosmology_b200 import priors
    g = _golden()
    assert tuple(str(s) for s in g["site_names"]) == tuple(priors.SITE_NAMES)
    # the reference's own declarations, as recorded when the golden was minted
    decl = dict(ln.split(": ", 1) for ln in (str(s) for s in g["distributions"]))
    assert decl["sigma"].startswith("TruncatedNormal(2, 2)") and "'high': None" in decl["sigma"]
    assert decl["beta"].startswith("Normal(0, 2)") and decl["R_unit"].startswith("Normal(0, 1)")
    assert decl["log_fpl"].startswith("Uniform(")


def test_python_prior_terms_match_the_reference_potential():
    from bumpcosmology_b200 import priors
    g = _golden()
    for k, u in enumerate(g["u"]):
        x, dx, lpj, glp, dlj = priors.potential_terms(u)
        assert _close(x, g["ref_x"][k], rtol=1e-13)
        assert _close(-lpj, g["ref_prior_U"][k], rtol=1e-12)
        grad = -(np.array(glp) * np.array(dx) + np.array(dlj))
        assert _close(grad, g["ref_prior_grad"][k], rtol=1e-11)
        # the documented composition gives the same numbers
        x2, dx2, lj, dlj2 = priors.constrain(u)
        lp, glp2 = priors.log_prior(x2)
        assert _close(-(lp + lj), g["ref_prior_U"][k], rtol=1e-12)


def test_cxx_prior_terms_match_the_reference_potential():
    from bumpcosmology_b200 import _lib
    lib = _lib.load()
    g = _golden()
    for k, u in enumerate(g["u"]):
        u = np.ascontiguousarray(u, dtype=np.float64)
        x, grad, pu = np.empty(15), np.empty(15), np.zeros(1)
        _lib.check(lib.bump_nuts_prior_terms(_lib.as_dp(u), _lib.as_dp(x), _lib.as_dp(pu), _lib.as_dp(grad)))
        assert _close(x, g["ref_x"][k], rtol=1e-13)
        assert _close(pu[0], g["ref_prior_U"][k], rtol=1e-12)
        assert _close(grad, g["ref_prior_grad"][k], rtol=1e-11)


@pytest.mark.gpu
def test_full_potential_and_gradient_match_the_reference(golden_dir):
    from bumpcosmology_b200 import _lib, intensity_models as im
    g = _golden()
    c = np.load(os.path.join(golden_dir, str(g["catalog"])))
    data = (c["m1s_det"], c["qs"], c["dls"], c["pdraw"], c["m1s_det_sel"], c["qs_sel"], c["dls_sel"], c["pdraw_sel"],
            float(c["Ndraw"]))
    model = im.pop_cosmo_model(*data)
    lib = _lib.load()
    for k, u in enumerate(g["u"]):
        gs = max(1.0, float(np.max(np.abs(g["ref_grad"][k]))))
        # Python driver's potential
        U, grad, rec = model.potential(u)
        assert _close(U, g["ref_U"][k]), (k, U, g["ref_U"][k])
        assert _close(grad, g["ref_grad"][k], floor=gs), (k, np.max(np.abs(grad - g["ref_grad"][k])))
        det = model.deterministics(rec)
        assert _close(det["loglike"], g["ref_loglike"][k]) and _close(det["selfactor"], g["ref_selfactor"][k])
        assert _close(det["R"], g["ref_R"][k])
        # C++ driver's potential (what bump_nuts_chain integrates)
        uu = np.ascontiguousarray(u, dtype=np.float64)
        Uc, gc, rc = np.zeros(1), np.empty(15), np.empty(_lib.NUTS_NDET)
        _lib.check(lib.bump_nuts_potential(model.like._ctx, _lib.as_dp(uu), _lib.as_dp(Uc), _lib.as_dp(gc),
                                           _lib.as_dp(rc)))
        assert _close(Uc[0], g["ref_U"][k]) and _close(gc, g["ref_grad"][k], floor=gs)
        assert _close(rc[0], g["ref_loglike"][k]) and _close(rc[3], g["ref_R"][k])
    model.close()
