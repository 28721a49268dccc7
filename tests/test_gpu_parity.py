"""Parity of the CUDA path (through the C ABI) with the golden vectors minted from the unmodified reference
and with the fp64 oracle.  Tolerance: 1e-10 relative (north star, fp64), with an absolute floor of 1e-10 times
the gradient scale for near-zero components."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"pop_cosmo_{name}.npz"))


def _data(g):
    return (g["m1s_det"], g["qs"], g["dls"], g["pdraw"], g["m1s_det_sel"], g["qs_sel"], g["dls_sel"],
            g["pdraw_sel"], float(g["Ndraw"]))


def _close(a, b, rtol=RTOL, floor=1.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= rtol * np.maximum(np.abs(b), floor)
    return bool(np.all(ok | same_inf))


def _sites_grad(g, th):
    from oracle import bump_oracle as bo
    return bo.grad_sites_from_theta(g, th)


@pytest.fixture(scope="module")
def hl():
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    return Hyperlikelihood


@pytest.mark.parametrize("name", ("tiny", "small"))
def test_cuda_matches_reference_goldens(golden_dir, hl, name):
    g = _load(golden_dir, name)
    like = hl(*_data(g))
    for k, th in enumerate(g["thetas"]):
        r = like(th)
        assert _close(r.loglike, g["ref_loglike"][k]), (k, r.loglike, g["ref_loglike"][k])
        assert _close(r.log_mu_sel, g["ref_log_mu_sel"][k])
        assert _close(r.selfactor, g["ref_selfactor"][k])
        assert _close(r.neff_sel, g["ref_neff_sel"][k])
        assert _close(r.neff, g["ref_neff"][k])
        scale = max(1.0, float(np.max(np.abs(g["ref_dloglike_dsite"][k]))))
        assert _close(_sites_grad(r.dloglike[:14], th), g["ref_dloglike_dsite"][k], floor=scale), k
        assert _close(_sites_grad(r.dlog_mu_sel[:14], th), g["ref_dlog_mu_sel_dsite"][k]), k
    like.close()


@pytest.mark.parametrize("name", ("tiny", "small"))
def test_cuda_tables_match_reference(golden_dir, hl, name):
    g = _load(golden_dir, name)
    like = hl(*_data(g))
    for k, th in enumerate(g["thetas"]):
        like(th)
        t = like.tables()
        for key in ("zinterp", "dlinterp", "ddlinterp", "dvcinterp"):
            assert _close(t[key], g["tab_" + key][k], rtol=1e-12, floor=1e-3), (k, key)
        # a log-density: absolute error is what propagates into the log-weights
        assert _close(t["log_dN_grid"], g["tab_log_dN_grid"][k], rtol=1e-11, floor=1.0), k
    like.close()


def test_cuda_table_tangents_match_oracle_autograd(golden_dir, hl):
    from oracle import bump_oracle as bo
    g = _load(golden_dir, "tiny")
    like = hl(*_data(g))
    for th in g["thetas"][:3]:
        like(th)
        t = like.tables()
        jac = bo.table_jacobians(th)
        sc = t["scalars"]
        for name, key in (("d_dl", "dlinterp"), ("d_ddl", "ddlinterp"), ("d_dvc", "dvcinterp")):
            for c, ith in enumerate((1, 2)):   # Om, w
                ref = jac[key][:, ith]
                assert _close(t[name][c], ref, rtol=1e-10, floor=max(1e-6, float(np.max(np.abs(ref))) * 1e-3))
        for c, ith in enumerate((3, 4, 6, 7, 8)):   # a, b, mpisn, mbhmax, sigma
            ref = jac["log_dN_grid"][:, ith]
            assert _close(t["d_log_dN_grid"][c], ref, rtol=1e-10, floor=max(1e-6, float(np.max(np.abs(ref))) * 1e-3))
        assert _close(sc[20:25], jac["log_pl_norm"][0, [3, 4, 6, 7, 8]], rtol=1e-10, floor=1e-3)
        assert _close(sc[25:32], jac["log_norm"][0, [3, 4, 5, 6, 7, 8, 9]], rtol=1e-10, floor=1e-3)
        assert _close(sc[32:34], jac["rate_log_norm"][0, [12, 13]], rtol=1e-10, floor=1e-3)
    like.close()


def test_cuda_matches_oracle_on_fresh_catalog(hl):
    """A catalog that is NOT in the fixtures: ragged sizes (odd nsamp / nsel exercise the sentinel padding,
    several tiles per event), 4 prior draws."""
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas, make_catalog
    from oracle import bump_oracle as bo
    cat = make_catalog("gwtc3", nobs=7, nsamp=3001, nsel=20001, seed=99)
    like = hl(*cat.as_args())
    for th in np.vstack([THETA_DEFAULT, draw_prior_thetas(4, seed=3)]):
        r = like(th)
        o = bo.evaluate(th, cat.as_args(), grad=True)
        assert _close(r.loglike, o["loglike"]) and _close(r.log_mu_sel, o["log_mu_sel"])
        assert _close(r.log_mu2, o["log_mu2"]) and _close(r.neff_sel, o["neff_sel"]) and _close(r.neff, o["neff"])
        assert _close(r.dloglike, o["dloglike"], floor=max(1.0, float(np.max(np.abs(o["dloglike"])))))
        assert _close(r.dlog_mu_sel, o["dlog_mu_sel"])
    like.close()


def test_wa_mode_matches_oracle(golden_dir, hl):
    from oracle import bump_oracle as bo
    g = _load(golden_dir, "tiny")
    like = hl(*_data(g), wa=True)
    for wa in (0.0, 0.4, -0.7):
        th = np.concatenate([g["thetas"][1], [wa]])
        r = like(th)
        o = bo.evaluate(th[:14], _data(g), grad=True, wa=wa)
        assert _close(r.loglike, o["loglike"]) and _close(r.log_mu_sel, o["log_mu_sel"])
        assert _close(r.dloglike[:14], o["dloglike"], floor=max(1.0, float(np.max(np.abs(o["dloglike"])))))
        assert _close(r.dlog_mu_sel[:14], o["dlog_mu_sel"])
        assert _close(r.dloglike[14], o["dloglike_dwa"], floor=max(1.0, abs(o["dloglike_dwa"])))
        assert _close(r.dlog_mu_sel[14], o["dlog_mu_sel_dwa"])
    like.close()


def test_edge_cases(hl):
    """Samples beyond the last d_L knot, exactly on knots, below mbh_min; an event split over tiles;
    one-sample events; a single injection."""
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
    from oracle import bump_oracle as bo
    cat = make_catalog("tiny", seed=5)
    args = list(cat.as_args())
    dls = args[2].copy()
    dls[0, :4] = [5e3, 1e4, 1e-4, 2e-3]          # beyond z = 100 and tiny distances
    m1 = args[0].copy()
    m1[1, :8] = 4.0                               # below mbh_min -> -inf weights
    args[2], args[0] = dls, m1
    like = hl(*args)
    o = bo.evaluate(THETA_DEFAULT, tuple(args), grad=True)
    r = like(THETA_DEFAULT)
    assert _close(r.loglike, o["loglike"]) and _close(r.neff, o["neff"])
    assert _close(r.dloglike, o["dloglike"], floor=max(1.0, float(np.max(np.abs(o["dloglike"])))))
    like.close()
    # one sample per event, one injection
    one = (args[0][:, :1], args[1][:, :1], args[2][:, 5:6], args[3][:, :1], args[4][:1], args[5][:1], args[6][:1],
           args[7][:1], 10.0)
    like = hl(*one)
    o = bo.evaluate(THETA_DEFAULT, one, grad=True)
    r = like(THETA_DEFAULT)
    assert _close(r.loglike, o["loglike"]) and _close(r.log_mu_sel, o["log_mu_sel"])
    assert _close(r.dlog_mu_sel, o["dlog_mu_sel"])
    finite = np.isfinite(o["neff"])
    assert _close(r.neff[finite], o["neff"][finite]) and np.all(np.isnan(r.neff[~finite]) | (r.neff[~finite] == 0))
    like.close()


@pytest.mark.parametrize("name,world", (("small", 3), ("tiny", 8)))
def test_partials_merge_equals_single_rank(hl, name, world):
    """Emulate several ranks on one GPU: shard, evaluate partials, merge on the host (same code as the device
    finalize) -> equals the unsharded evaluation to 1e-12.  ("tiny", 8): 5 events over 8 ranks leaves ranks with
    no event at all."""
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
    from bumpcosmology_b200.likelihood import merge_partials, shard_catalog, unpack_header
    cat = make_catalog(name, seed=17)
    full = hl(*cat.as_args())
    r = full(THETA_DEFAULT)
    parts, neffs = [], []
    for rank in range(world):
        sh = hl(*shard_catalog(cat.as_args(), rank, world))
        p, ne = sh.partial(THETA_DEFAULT)
        parts.append(p)
        neffs.append(ne)
        sh.close()
    m = unpack_header(merge_partials(np.array(parts)), 14)
    assert _close(m["loglike"], r.loglike, rtol=1e-12) and _close(m["log_mu_sel"], r.log_mu_sel, rtol=1e-12)
    assert _close(m["dloglike"], r.dloglike, rtol=1e-11, floor=max(1.0, float(np.max(np.abs(r.dloglike)))))
    assert _close(m["dlog_mu_sel"], r.dlog_mu_sel, rtol=1e-11)
    assert _close(np.concatenate(neffs), r.neff, rtol=1e-12)
    full.close()


def test_out_of_support_theta_gives_nan_not_error(hl):
    """mbhmax < mpisn (sqrt of a negative number in largest_mco) and a NaN parameter: the reference yields NaN,
    NUTS treats it as a divergent step; the library must not raise, and must recover on the next call."""
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
    cat = make_catalog("tiny")
    like = hl(*cat.as_args())
    good = like(THETA_DEFAULT)
    bad = THETA_DEFAULT.copy()
    bad[7] = bad[6] - 1.0
    r = like(bad)
    assert np.isnan(r.loglike) and np.all(np.isnan(r.dloglike))
    bad = THETA_DEFAULT.copy()
    bad[0] = np.nan
    assert np.isnan(like(bad).log_mu_sel)
    again = like(THETA_DEFAULT)
    assert again.loglike == good.loglike and np.array_equal(again.dloglike, good.dloglike)
    like.close()


def test_repeat_evaluations_are_bitwise_identical(hl):
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
    cat = make_catalog("small")
    like = hl(*cat.as_args())
    a, b = like(THETA_DEFAULT), like(THETA_DEFAULT)
    assert a.loglike == b.loglike and a.log_mu_sel == b.log_mu_sel
    assert np.array_equal(a.dloglike, b.dloglike) and np.array_equal(a.neff, b.neff)
    like.close()


def test_alternating_theta_never_sees_the_previous_evaluation(golden_dir, hl):
    """The prologue writes the theta-dependent scalars straight into the constant bank, and the stream kernel and the
    epilogue are programmatic dependents that may be scheduled before their predecessor has finished: a read that
    came too early, or a constant served from a cache line of the previous evaluation, would show up as the result
    of the OTHER theta (or a mix).  Back-to-back graph replays with theta changing every time, checked bitwise
    against each theta's first result."""
    g = _load(golden_dir, "small")
    like = hl(*_data(g))
    thetas = [np.array(t, dtype=float) for t in g["thetas"]]
    first = [like.raw(t).copy() for t in thetas]
    for k, t in enumerate(thetas):   # the first pass itself is right (against the reference's goldens)
        assert _close(first[k][0], g["ref_loglike"][k]), k
    for rep in range(300):
        k = (rep * 7 + rep // 3) % len(thetas)
        assert np.array_equal(like.raw(thetas[k]), first[k], equal_nan=True), (rep, k)
    like.close()


def test_host_model_mirror_matches_reference_sites_and_curves(golden_dir):
    """bumpcosmology_b200.intensity_models.pop_cosmo_model: same signature and site names as the reference;
    factors, deterministics, site gradients and the 128-point diagnostic curves against the goldens."""
    from bumpcosmology_b200 import intensity_models as im
    g = _load(golden_dir, "small")
    model = im.pop_cosmo_model(*_data(g))
    names = [str(s) for s in g["site_names"]]
    assert tuple(names) == im.LIKELIHOOD_SITES
    for k, th in enumerate(g["thetas"]):
        sites = dict(h=th[0], Om=th[1], w=th[2], a=th[3], b=th[4], c=th[5], mpisn=th[6], dmbhmax=th[7] - th[6],
                     sigma=th[8], beta=th[10], log_fpl=np.log(th[9]), lam=th[11], dkappa=th[12] - th[11], zp=th[13],
                     R_unit=0.25)
        r = model(sites, diagnostics=True)
        assert _close(r["loglike"], g["ref_loglike"][k]) and _close(r["selfactor"], g["ref_selfactor"][k])
        assert _close(r["R"], g["ref_R"][k]) and _close(r["neff_sel"], g["ref_neff_sel"][k])
        scale = max(1.0, float(np.max(np.abs(g["ref_dloglike_dsite"][k]))))
        assert _close(r["dloglike_dsite"], g["ref_dloglike_dsite"][k], floor=scale)
        for name in ("mdNdmdVdt_fixed_qz", "dNdqdVdt_fixed_mz", "dNdVdt_fixed_mq", "hz"):
            ref = g["ref_" + name][k]
            assert _close(r[name], ref, rtol=1e-9, floor=max(1e-300, float(np.max(np.abs(ref))) * 1e-6)), (k, name)
    model.close()


def test_fixed_cosmology_mode_matches_reference_pop_model(golden_dir):
    """BUMP_FLAG_FIXED_COSMO through the host mirror `pop_model(...)`: the golden is the unmodified reference
    `pop_model` (intensity_models.py:313-355)."""
    from bumpcosmology_b200 import intensity_models as im
    g = np.load(os.path.join(golden_dir, "pop_fixed_small.npz"))
    data = (g["m1s"], g["qs"], g["zs"], g["pdraw"], g["m1s_sel"], g["qs_sel"], g["zs_sel"], g["pdraw_sel"],
            float(g["Ndraw"]))
    model = im.pop_model(*data, dVdzdt_interp=g["dvdzdt_interp"])
    assert tuple(str(s) for s in g["site_names"]) == im.FIXED_SITES[:11]
    for k, th in enumerate(g["thetas"]):
        sites = dict(a=th[3], b=th[4], c=th[5], mpisn=th[6], dmbhmax=th[7] - th[6], sigma=th[8], beta=th[10],
                     log_fpl=np.log(th[9]), lam=th[11], dkappa=th[12] - th[11], zp=th[13], R_unit=-0.5)
        r = model(sites)
        assert _close(r["loglike"], g["ref_loglike"][k]) and _close(r["selfactor"], g["ref_selfactor"][k])
        assert _close(r["neff_sel"], g["ref_neff_sel"][k]) and _close(r["neff"], g["ref_neff"][k])
        assert _close(r["R"], g["ref_R"][k])
        scale = max(1.0, float(np.max(np.abs(g["ref_dloglike_dsite"][k]))))
        assert _close(r["dloglike_dsite"], g["ref_dloglike_dsite"][k], floor=scale)
        assert _close(r["dselfactor_dsite"], -r["nobs"] * g["ref_dlog_mu_sel_dsite"][k], floor=float(r["nobs"]))
    model.close()


def test_invalid_inputs_are_rejected_at_upload(hl):
    from bumpcosmology_b200._lib import BumpError
    from bumpcosmology_b200.catalogs import make_catalog
    cat = make_catalog("tiny")
    args = list(cat.as_args())
    bad = args[3].copy()
    bad[2, 5] = 0.0                      # pdraw = 0 -> log(pdraw) = -inf in the reference
    args[3] = bad
    with pytest.raises(BumpError, match="finite and strictly positive"):
        hl(*args)
    args = list(cat.as_args())
    bad = args[6].copy()
    bad[7] = np.nan
    args[6] = bad
    with pytest.raises(BumpError):
        hl(*args)
    with pytest.raises(ValueError):
        hl(args[0][:, :3], *cat.as_args()[1:])   # shape mismatch


def test_unsorted_upload_gives_the_same_answer(hl):
    """The locality sort at upload only permutes samples inside an event: results agree to rounding."""
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
    cat = make_catalog("small", seed=23)
    a = hl(*cat.as_args())(THETA_DEFAULT)
    b = hl(*cat.as_args(), sort=False)(THETA_DEFAULT)
    assert _close(a.loglike, b.loglike, rtol=1e-13) and _close(a.neff, b.neff, rtol=1e-12)
    assert _close(a.dloglike, b.dloglike, rtol=1e-12, floor=max(1.0, float(np.max(np.abs(b.dloglike)))))


def test_sampler_potential_fast_path_matches_evaluate(golden_dir):
    """PopCosmoModel.potential (raw library call + plain-float chain rule) against the documented composition
    priors.constrain / log_prior / evaluate, and against finite differences in unconstrained space."""
    from bumpcosmology_b200 import intensity_models as im, priors
    g = _load(golden_dir, "small")
    model = im.pop_cosmo_model(*_data(g))
    rng = np.random.default_rng(4)
    for _ in range(4):
        u = rng.uniform(-1, 1, priors.NSITES)
        U, grad, rec = model.potential(u)
        x, dx, lj, dlj = priors.constrain(u)
        lp, glp = priors.log_prior(x)
        ev = model.evaluate(x)
        gg = glp.copy()
        gg[:14] += ev["dloglike_dsite"] + ev["dselfactor_dsite"]
        assert _close(U, -(lp + lj + ev["loglike"] + ev["selfactor"]), rtol=1e-12)
        assert _close(grad, -(gg * dx + dlj), rtol=1e-11, floor=max(1.0, float(np.max(np.abs(grad)))))
        d = model.deterministics(rec)
        assert _close(d["R"], ev["R"], rtol=1e-12) and _close(d["neff"], ev["neff"], rtol=1e-12)
        for i in (3, 9, 11):   # a, beta, lam: smooth directions
            up, um = u.copy(), u.copy()
            up[i] += 1e-6
            um[i] -= 1e-6
            fd = (model.potential(up)[0] - model.potential(um)[0]) / 2e-6
            assert abs(fd - grad[i]) <= 1e-5 * max(1.0, abs(fd))
    model.close()


def test_concurrent_contexts_on_one_device_do_not_interfere(golden_dir, hl):
    """Parallel NUTS chains: one context per chain, each bound to its own constant-bank slot (4 per device), evaluated
    from concurrent host threads with different theta.  Every result must be bitwise what the same context returns
    alone; six contexts also cover two of them sharing a slot (those are chained, not overlapped)."""
    import threading
    g = _load(golden_dir, "small")
    thetas = g["thetas"]
    likes = [hl(*_data(g)) for _ in range(6)]
    alone = [likes[i].raw(thetas[i % len(thetas)]).copy() for i in range(len(likes))]
    errors = []

    def work(i):
        th = thetas[i % len(thetas)]
        for _ in range(200):
            out = likes[i].raw(th)
            if not np.array_equal(out, alone[i], equal_nan=True):
                errors.append(i)
                return

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(likes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for like in likes:
        like.close()
    assert not errors, f"contexts {sorted(set(errors))} saw another context's parameters"


def test_hubble_far_outside_the_prior_is_flagged_or_exact(hl):
    """The single-step d_L bin search is exact while the clamped end buckets of the search table hold at most two
    bins, i.e. for roughly 0.11 < h < 7 (prior support: 0.35 .. 1.4).  Inside that range the result must still match the
    oracle; beyond it the evaluation is flagged (NaN), never silently wrong."""
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
    from oracle import bump_oracle as bo
    cat = make_catalog("tiny")
    like = hl(*cat.as_args())
    for h in (0.15, 5.0):
        th = THETA_DEFAULT.copy()
        th[0] = h
        r = like(th)
        o = bo.evaluate(th, cat.as_args(), grad=True)
        assert _close(r.log_mu_sel, o["log_mu_sel"]) and _close(r.dlog_mu_sel, o["dlog_mu_sel"]), h
        if np.isfinite(o["loglike"]):
            assert _close(r.loglike, o["loglike"]), h
    for h in (0.02, 20.0):
        th = THETA_DEFAULT.copy()
        th[0] = h
        assert np.isnan(like(th).log_mu_sel), h
    like.close()


def test_device_math(hl):
    """The streaming kernel's own exp / reciprocal / log1p / 1/(1+x) against numpy in the ranges the kernel uses them:
    a few ulp (the kernel's arithmetic is sized for 1e-10 on sums of 6e7 terms; these are 1e-15 - 1e-14)."""
    import ctypes as C
    from bumpcosmology_b200 import _lib
    from bumpcosmology_b200.catalogs import make_catalog
    like = hl(*make_catalog("tiny").as_args())
    rng = np.random.default_rng(2)

    def probe(which, x):
        y = np.empty_like(x)
        _lib.check(like.lib.bump_debug_math(like._ctx, which, _lib.as_dp(x), x.shape[0], _lib.as_dp(y)))
        return y

    x = np.concatenate([rng.uniform(-60, 60, 200_000), rng.uniform(-300, 260, 100_000), rng.uniform(-1e-3, 1e-3, 1000)])
    rel = np.abs(probe(0, x) / np.exp(x) - 1)
    assert np.max(rel / np.maximum(1.0, np.abs(x))) < 4e-16, float(np.max(rel))      # |x| 1.1e-16 + polynomial
    xw = np.concatenate([rng.uniform(-690, 700, 200_000), [0.0, -0.0, 1e-300, 709.0, -690.0]])
    assert np.max(np.abs(probe(1, xw) / np.exp(xw) - 1)) < 1.5e-15
    # below about -700 the result saturates near 1e-304 instead of underflowing: callers treat it as zero
    assert np.all(probe(1, np.array([-705.0, -800.0, -1e4, -9e4])) < 1e-300)
    xr = np.exp(rng.uniform(np.log(1e-200), np.log(1e200), 300_000))
    yr = probe(2, xr)
    err = np.abs(yr * xr - 1)
    k = int(np.argmax(err))
    bad = np.flatnonzero(~(err < 5e-16))
    assert bad.size == 0, (float(err[k]), float(xr[k]), float(yr[k]), bad.size, bad[:8], bad[-4:], yr[bad[:4]],
                           probe(2, xr)[bad[:4]], probe(2, xr[bad[:4]]))
    xs = np.concatenate([rng.uniform(0, 0.00453, 200_000), [0.0, 0.00453]])
    assert np.max(np.abs(probe(3, xs) - np.log1p(xs))) < 1e-17 + 2.3e-16 * 0.00453
    assert np.max(np.abs(probe(4, xs) * (1 + xs) - 1)) < 4e-16
    like.close()
