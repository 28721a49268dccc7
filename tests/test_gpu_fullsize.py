"""BASELINE.json's full sizes on the GPU.

* O4 shape (3.46e6 samples): direct parity with the torch oracle.
* FULL O5 (5000 x 10000 + 1e7) and FULL GWTC-3 (69 x 4096 + 2e5) shapes: every output (loglike, log_mu_sel, log_mu2,
  neff_sel, neff[nobs], both 14-parameter gradients) against the fused C++ CPU port (oracle/bump_cpu.cpp, pinned to
  the reference-minted goldens at 1e-10 in tests/test_cpu_port.py), all host threads, two theta each:
  /root/reference/src/scripts/intensity_models.py:378-394,401 at the headline sizes.
* w0-wa mode at wa != 0 on a 1 % sub-catalog of O5 against the torch oracle's CPL branch.
* size-independent properties at the O5 shape: shard additivity, permutation invariance, exact shifts under rescaling
  of pdraw, oracle parity of the per-event Neff on a random subset of events.

Tolerance: 1e-10 relative (north star, fp64), gradients with an absolute floor of 1e-10 x the gradient scale."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


@pytest.fixture(scope="module")
def o5_catalog():
    from bumpcosmology_b200.catalogs import make_catalog
    return make_catalog("o5")


def _assert_matches_cpu_port(cat, thetas):
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    from oracle.bump_cpu import CpuPort
    port = CpuPort(*cat.as_args())
    like = Hyperlikelihood(*cat.as_args())
    worst = {}
    for th in thetas:
        r = like(th)
        o = port.evaluate(th, nthreads=_host_threads())
        gs = max(1.0, float(np.max(np.abs(o["dloglike"]))))
        pairs = {"loglike": (r.loglike, o["loglike"], 1.0), "log_mu_sel": (r.log_mu_sel, o["log_mu_sel"], 1.0),
                 "log_mu2": (r.log_mu2, o["log_mu2"], 1.0), "neff_sel": (r.neff_sel, o["neff_sel"], 1.0),
                 "neff": (r.neff, o["neff"], 1.0), "dloglike": (r.dloglike[:14], o["dloglike"], gs),
                 "dlog_mu_sel": (r.dlog_mu_sel[:14], o["dlog_mu_sel"], 1.0)}
        assert r.nobs == cat.nobs and r.nsel == cat.nsel and r.neff.shape == (cat.nobs,)
        for k, (a, b, floor) in pairs.items():
            a, b = np.asarray(a, float), np.asarray(b, float)
            err = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
            worst[k] = max(worst.get(k, 0.0), err)
            assert err <= RTOL, (k, err, th)
    like.close()
    port.close()
    print("max rel err vs CPU port:", worst)


def test_o5_full_size_matches_cpu_port(o5_catalog):
    """The headline configuration itself: all 4 + 28 + 5000 outputs at 6.0e7 samples."""
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas
    _assert_matches_cpu_port(o5_catalog, (THETA_DEFAULT, draw_prior_thetas(1, seed=8)[0]))


def test_gwtc3_full_size_matches_cpu_port():
    """BASELINE.json config 1 at its full shape (69 x 4096 + 2e5), not a stand-in."""
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas, make_catalog
    _assert_matches_cpu_port(make_catalog("gwtc3"), (THETA_DEFAULT, draw_prior_thetas(1, seed=8)[0]))


def test_o5_wa_subcatalog_matches_oracle(o5_catalog):
    """BASELINE.json config 5 (w0-wa) at wa != 0: 1 % of the O5 catalog (50 events x 10000 + 1e5 injections) against
    the torch oracle's CPL branch (no reference counterpart: pinned to the reference only at wa = 0)."""
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    from oracle import bump_oracle as bo
    a = o5_catalog.as_args()
    sub = (a[0][:50], a[1][:50], a[2][:50], a[3][:50], a[4][:100_000], a[5][:100_000], a[6][:100_000],
           a[7][:100_000], a[8])
    like = Hyperlikelihood(*sub, wa=True)
    for th14, wa in ((THETA_DEFAULT, 0.35), (draw_prior_thetas(1, seed=8)[0], -0.6)):
        r = like(np.concatenate([th14, [wa]]))
        o = bo.evaluate(th14, sub, grad=True, wa=wa, event_chunk=16)
        gs = max(1.0, float(np.max(np.abs(o["dloglike"]))))
        assert _close(r.loglike, o["loglike"]) and _close(r.log_mu_sel, o["log_mu_sel"])
        assert _close(r.log_mu2, o["log_mu2"]) and _close(r.neff_sel, o["neff_sel"]) and _close(r.neff, o["neff"])
        assert _close(r.dloglike[:14], o["dloglike"], floor=gs) and _close(r.dlog_mu_sel[:14], o["dlog_mu_sel"])
        assert _close(r.dloglike[14], o["dloglike_dwa"], floor=gs) and _close(r.dlog_mu_sel[14], o["dlog_mu_sel_dwa"])
    like.close()


def _close(a, b, rtol=RTOL, floor=1.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), floor)))


def test_o4_shape_matches_oracle():
    import torch
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas, make_catalog
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    from oracle import bump_oracle as bo
    torch.set_num_threads(max(1, torch.get_num_threads()))
    cat = make_catalog("o4")
    like = Hyperlikelihood(*cat.as_args())
    for th in (THETA_DEFAULT, draw_prior_thetas(1, seed=8)[0]):
        r = like(th)
        o = bo.evaluate(th, cat.as_args(), grad=True, event_chunk=64)
        assert _close(r.loglike, o["loglike"]) and _close(r.log_mu_sel, o["log_mu_sel"])
        assert _close(r.log_mu2, o["log_mu2"]) and _close(r.neff_sel, o["neff_sel"]) and _close(r.neff, o["neff"])
        assert _close(r.dloglike, o["dloglike"], floor=max(1.0, float(np.max(np.abs(o["dloglike"])))))
        assert _close(r.dlog_mu_sel, o["dlog_mu_sel"])
    like.close()


def test_o5_shape_properties(o5_catalog):
    from bumpcosmology_b200.catalogs import THETA_DEFAULT
    from bumpcosmology_b200.likelihood import Hyperlikelihood, merge_partials, shard_catalog, unpack_header
    from oracle import bump_oracle as bo
    cat = o5_catalog
    args = cat.as_args()
    like = Hyperlikelihood(*args)
    r = like(THETA_DEFAULT)
    assert np.isfinite(r.logl) and np.all(np.isfinite(r.dlogl)) and r.nobs == 5000 and r.nsel == 10_000_000
    gscale = max(1.0, float(np.max(np.abs(r.dloglike))))
    # (1) determinism
    r2 = like(THETA_DEFAULT)
    assert r2.loglike == r.loglike and np.array_equal(r2.dloglike, r.dloglike)
    like.close()
    # (2) shard additivity: 3 shards evaluated separately, merged by the library's rank-ordered merge
    parts, neffs = [], []
    for rank in range(3):
        sh = Hyperlikelihood(*shard_catalog(args, rank, 3))
        p, ne = sh.partial(THETA_DEFAULT)
        parts.append(p)
        neffs.append(ne)
        sh.close()
    m = unpack_header(merge_partials(np.array(parts)), 14)
    assert _close(m["loglike"], r.loglike, rtol=1e-12) and _close(m["log_mu_sel"], r.log_mu_sel, rtol=1e-12)
    assert _close(m["dloglike"], r.dloglike, rtol=1e-11, floor=gscale)
    assert _close(np.concatenate(neffs), r.neff, rtol=1e-12)
    # (3) permutation invariance inside events and of the injections, and (4) exact shifts under pdraw -> c pdraw:
    #     loglike -> loglike - nobs log c, log_mu_sel -> log_mu_sel - log c2, gradients and Neff unchanged
    rng = np.random.default_rng(0)
    perm = rng.permutation(cat.nsamp)
    sperm = rng.permutation(cat.nsel)
    c1, c2 = 3.5, 0.25
    args2 = (args[0][:, perm], args[1][:, perm], args[2][:, perm], c1 * args[3][:, perm],
             args[4][sperm], args[5][sperm], args[6][sperm], c2 * args[7][sperm], args[8])
    like2 = Hyperlikelihood(*args2, sort=False)
    q = like2(THETA_DEFAULT)
    like2.close()
    assert _close(q.loglike, r.loglike - r.nobs * np.log(c1), rtol=1e-12)
    assert _close(q.log_mu_sel, r.log_mu_sel - np.log(c2), rtol=1e-12)
    assert _close(q.dloglike, r.dloglike, rtol=1e-11, floor=gscale) and _close(q.dlog_mu_sel, r.dlog_mu_sel, rtol=1e-11)
    assert _close(q.neff, r.neff, rtol=1e-11) and _close(q.neff_sel, r.neff_sel, rtol=1e-10)
    # (5) oracle parity of the per-event Neff on a random subset of events
    idx = np.sort(rng.choice(cat.nobs, 40, replace=False))
    sub = (args[0][idx], args[1][idx], args[2][idx], args[3][idx], args[4][:1000], args[5][:1000], args[6][:1000],
           args[7][:1000], args[8])
    o = bo.evaluate(THETA_DEFAULT, sub, grad=False)
    assert _close(r.neff[idx], o["neff"])
