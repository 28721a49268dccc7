"""JAX binding of the hot path: an XLA typed-FFI custom call (csrc/bump_xla_ffi.cc -> `BumpLoglikeFfi`) wrapped in a
`custom_vjp`, and a drop-in `pop_cosmo_model` for numpyro that keeps the reference's sample sites, deterministics and
factors (/root/reference/src/scripts/intensity_models.py:357-401) and replaces only lines :374-394,401.

JAX is not installable in the image this repository is developed in, so nothing here runs in its test suite: the
module raises ImportError on import without jax/jaxlib/numpyro, and the parity of the path it calls is established
through the ctypes binding (tests/test_gpu_parity.py), which enters the library through the same `bump_eval*`
entry points.
"""
import ctypes
import os

try:
    import jax
    import jax.numpy as jnp
except ImportError as e:   # pragma: no cover
    raise ImportError("bumpcosmology_b200.jax_ffi needs jax + jaxlib (and numpyro for the model); "
                      "use bumpcosmology_b200.intensity_models / nuts for the JAX-free host path") from e

from .. import _lib
from ..likelihood import Hyperlikelihood

_FFI_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libbump_xla_ffi.so")
_registered = False


def _register():
    global _registered
    if _registered:
        return
    if not os.path.exists(_FFI_LIB):
        raise ImportError(f"{_FFI_LIB} is missing: build it with `python -m bumpcosmology_b200._build --xla-ffi`")
    _lib.load()                                             # libbump_b200.so first (RTLD_GLOBAL): the adaptor links to it
    lib = ctypes.CDLL(_FFI_LIB)
    jax.ffi.register_ffi_target("bump_loglike", jax.ffi.pycapsule(lib.BumpLoglikeFfi), platform="CUDA")
    _registered = True


def make_factors(*data, device=0):
    """Upload the catalog once (outside jit) and return `factors(theta) -> (loglike, log_mu_sel, neff_sel, neff)`,
    differentiable with respect to theta = (h, Om, w, a, b, c, mpisn, mbhmax, sigma, fpl, beta, lam, kappa, zp)."""
    _register()
    jax.config.update("jax_enable_x64", True)
    like = Hyperlikelihood(*data, device=device)
    n_out = _lib.OUT_HEADER + like.nobs
    call = jax.ffi.ffi_call("bump_loglike", jax.ShapeDtypeStruct((n_out,), jnp.float64))
    handle = int(like._ctx.value)

    def _raw(theta):
        return call(theta.astype(jnp.float64), ctx=handle)

    @jax.custom_vjp
    def factors(theta):
        out = _raw(theta)
        return out[_lib.OUT_LOGLIKE], out[_lib.OUT_LOG_MU_SEL], out[_lib.OUT_NEFF_SEL], out[_lib.OUT_HEADER:]

    def fwd(theta):
        out = _raw(theta)
        res = (out[_lib.OUT_DLOGLIKE:_lib.OUT_DLOGLIKE + 14], out[_lib.OUT_DLOG_MU:_lib.OUT_DLOG_MU + 14])
        return (out[_lib.OUT_LOGLIKE], out[_lib.OUT_LOG_MU_SEL], out[_lib.OUT_NEFF_SEL], out[_lib.OUT_HEADER:]), res

    def bwd(res, ct):          # the backward pass is the stored gradient times the cotangents of the two factors
        g_ll, g_mu = res
        return (ct[0] * g_ll + ct[1] * g_mu,)

    factors.defvjp(fwd, bwd)
    factors.likelihood = like   # keeps the context (and its device memory) alive as long as the function
    return factors


def pop_cosmo_model(m1s_det, qs, dls, pdraw, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw, factors=None):
    """numpyro model with the reference's signature and site names (intensity_models.py:357-401); the likelihood
    factors come from the CUDA custom call.  Bind the data once: `factors = make_factors(*args)` and pass it in (a
    model body runs under tracing and must not upload)."""
    import numpyro
    import numpyro.distributions as dist
    if factors is None:
        raise ValueError("pass factors=make_factors(*the same nine arguments): uploads happen outside the trace")
    nobs = jnp.asarray(m1s_det).shape[0]
    tn = dist.TruncatedNormal
    # cosmo_parameters (:304-311), mass_parameters (:281-296), redshift_parameters (:298-303)
    h = numpyro.sample("h", tn(0.7, 0.2, low=0.35, high=1.4))
    Om = numpyro.sample("Om", tn(0.3, 0.15, low=0.0, high=1.0))
    w = numpyro.sample("w", tn(-1.0, 0.25, low=-1.5, high=-0.5))
    a = numpyro.sample("a", tn(2.35, 2.0, low=-1.65, high=6.35))
    b = numpyro.sample("b", tn(1.9, 2.0, low=-2.1, high=5.9))
    c = numpyro.sample("c", tn(4.0, 2.0, low=0.0, high=8.0))
    mpisn = numpyro.sample("mpisn", tn(35.0, 5.0, low=20.0, high=50.0))
    dmbhmax = numpyro.sample("dmbhmax", tn(5.0, 2.0, low=0.5, high=11.0))
    mbhmax = numpyro.deterministic("mbhmax", mpisn + dmbhmax)
    sigma = numpyro.sample("sigma", tn(2.0, 2.0, low=1.0))
    beta = numpyro.sample("beta", dist.Normal(0.0, 2.0))
    log_fpl = numpyro.sample("log_fpl", dist.Uniform(jnp.log(1e-3), jnp.log(0.5)))
    fpl = numpyro.deterministic("fpl", jnp.exp(log_fpl))
    lam = numpyro.sample("lam", tn(2.7, 2.0, low=-1.3, high=6.7))
    dkappa = numpyro.sample("dkappa", tn(5.6 - 2.7, 2.0, low=1.0, high=9.6 - 2.7))
    kappa = numpyro.deterministic("kappa", lam + dkappa)
    zp = numpyro.sample("zp", tn(1.9, 1.0, low=0.0, high=3.9))
    theta = jnp.stack([h, Om, w, a, b, c, mpisn, mbhmax, sigma, fpl, beta, lam, kappa, zp])
    loglike, log_mu_sel, neff_sel, neff = factors(theta)
    numpyro.factor("loglike", loglike)                       # :383
    numpyro.factor("selfactor", -nobs * log_mu_sel)          # :390
    numpyro.deterministic("neff_sel", neff_sel)              # :394
    mu = jnp.exp(log_mu_sel)
    r_unit = numpyro.sample("R_unit", dist.Normal(0.0, 1.0)) # :398
    numpyro.deterministic("R", nobs / mu + jnp.sqrt(nobs) / mu * r_unit)   # :396-399
    numpyro.deterministic("neff", neff)                      # :401
