"""The reference's fit driver (/root/reference/src/scripts/run_cosmo_fit.py) on the CUDA hot path.

    python -m bumpcosmology_b200.run_cosmo_fit --pe pe-samples.h5 --sel selection-samples.h5 --out trace_cosmo.npz

Same steps as the reference script: read the posterior-sample table (columns m1, q, z, wt, evt; draw_pe_samples.py:24)
and the found-injection table (m1, q, z, pdraw, ndraw; draw_selection_samples.py:15), convert to the detector frame
with the Jacobian (run_cosmo_fit.py:22-30, weighting.py:173-180), stack the per-event rows (:32-43), run
NUTS(dense_mass=True) with 1000 + 1000 steps on 4 chains and seed 1652819403 (:17-19,45-49), save the trace (:51-53).
Differences forced by the environment: the sampler is the library's C++ NUTS driver instead of numpyro (one context
and one host thread per chain), astropy's Planck18 is replaced by `inputs.FlatLCDM` (same H0, Om0; no radiation), and
the trace is a NumPy .npz (arviz / netCDF are not installed) — an arviz InferenceData is built as well when arviz is.
Tables may be HDF5 (needs PyTables, as in the reference), Parquet, CSV or .npz.
"""
import argparse
import json
import os
import time

import numpy as np

from . import inputs, priors

NMCMC, NCHAIN, RANDOM_SEED = 1000, 4, 1652819403   # run_cosmo_fit.py:17-19

# What the reference's trace holds besides the 15 sample sites: every numpyro.deterministic of pop_cosmo_model
# (intensity_models.py:288,294,301,394,399,401,403-406).  Downstream scripts read them by these names
# (e.g. dNdm_fitted.py:15 uses trace.posterior.mdNdmdVdt_fixed_qz).
REFERENCE_DETERMINISTICS = ("mbhmax", "fpl", "kappa", "neff_sel", "R", "neff", "mdNdmdVdt_fixed_qz",
                            "dNdqdVdt_fixed_mz", "dNdVdt_fixed_mq", "hz")
_CURVES = ("neff", "mdNdmdVdt_fixed_qz", "dNdqdVdt_fixed_mz", "dNdVdt_fixed_mq", "hz")


def read_table(path, key="samples"):
    """A table of named columns as {name: 1-D array}.  `.h5` / `.hdf5` are read like the reference does
    (`pd.read_hdf(path, 'samples')`)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        with np.load(path) as z:
            return {k: np.asarray(z[k]) for k in z.files}
    import pandas as pd
    if ext in (".h5", ".hdf5", ".hdf"):
        try:
            df = pd.read_hdf(path, key)
        except ImportError as e:
            raise ImportError(f"{path}: reading HDF5 needs PyTables (pip install tables); "
                              "Parquet, CSV and .npz tables work without it") from e
    elif ext in (".parquet", ".pq"):
        df = pd.read_parquet(path)
    elif ext in (".csv", ".txt"):
        df = pd.read_csv(path)
    else:
        raise ValueError(f"{path}: unknown table format (use .h5, .parquet, .csv or .npz)")
    return {c: df[c].to_numpy() for c in df.columns}


def fit(pe, sel, num_warmup=NMCMC, num_samples=NMCMC, num_chains=NCHAIN, seed=RANDOM_SEED, device=0, cosmo=None,
        native=True, deterministics=True):
    """Tables -> trace dict.  `pe`, `sel`: mappings of columns (pandas DataFrames work).  The trace holds the 15 sample
    sites (`posterior`) and, as `det_<name>`, every deterministic the reference's model registers
    (REFERENCE_DETERMINISTICS) plus the two factors."""
    from . import intensity_models as im, nuts
    args = inputs.model_arguments(pe, sel, cosmo)
    first = im.pop_cosmo_model(*args, device=device)            # one upload, one copy of the catalog in HBM ...
    models = [first] + [first.clone() for _ in range(num_chains - 1)]   # ... and one context per chain on it
    try:
        t0 = time.perf_counter()
        r = nuts.run_mcmc(models, num_warmup, num_samples, num_chains, seed=seed, native=native)
        wall = time.perf_counter() - t0
        # The vector-valued deterministics (per-event neff :401, the three 128-point rate curves :403-405, hz :406) are
        # not carried through the sampler: one more evaluation per kept draw recomputes them from the draw (the
        # PISN table and the cosmology come from the device's own tables for that theta).
        t0 = time.perf_counter()
        curves = posterior_deterministics(models[0], r["x"]) if deterministics else {}
        post_s = time.perf_counter() - t0
    finally:
        for m in models:
            m.close()
    trace = {"site_names": np.array(priors.SITE_NAMES), "posterior": r["x"],                      # [chain, draw, site]
             "ess_bulk": r["ess_bulk"], "rhat": r["rhat"], "wall_s": wall, "deterministics_s": post_s,
             "warmup_s": r["warmup_s"], "sampling_s": r["sampling_s"], "n_leapfrog": r["n_leapfrog_total"],
             "nobs": args[0].shape[0], "nsamp": args[0].shape[1], "nsel": len(args[4])}
    for k in r["chains"][0]["stats"]:
        trace["stat_" + k] = np.stack([np.asarray(c["stats"][k]) for c in r["chains"]])
    for k in r["chains"][0]["deterministic"]:                                                      # :288-301,394-401
        trace["det_" + k] = np.stack([np.asarray(c["deterministic"][k]) for c in r["chains"]])
    for k, v in curves.items():
        trace["det_" + k] = v
    return trace


def posterior_deterministics(model, x):
    """x: [chain, draw, 15] constrained draws -> {name: [chain, draw, len]} for the reference's vector-valued
    deterministics: neff [nobs] (intensity_models.py:401), mdNdmdVdt_fixed_qz, dNdqdVdt_fixed_mz, dNdVdt_fixed_mq
    [128] (:403-405) and hz [128] (:406)."""
    nc, nd = x.shape[:2]
    out = {}
    for c in range(nc):
        for d in range(nd):
            ev = model.evaluate(x[c, d], diagnostics=True)
            for k in _CURVES:
                v = np.asarray(ev[k], dtype=np.float64)
                if k not in out:
                    out[k] = np.empty((nc, nd) + v.shape)
                out[k][c, d] = v
    return out


def posterior_variables(trace):
    """{variable name: array [chain, draw, ...]} under the reference's names: what `az.from_numpyro(mcmc)` exposes as
    `trace.posterior` for the reference (run_cosmo_fit.py:51): sample sites + deterministics."""
    post = {str(n): trace["posterior"][:, :, i] for i, n in enumerate(trace["site_names"])}
    for k, v in trace.items():
        if k.startswith("det_"):
            post[k[4:]] = v
    return post


def to_inference_data(trace):
    """arviz.InferenceData with the reference's variable names (run_cosmo_fit.py:51), if arviz is installed."""
    import arviz as az
    stats = {k[5:]: v for k, v in trace.items() if k.startswith("stat_")}
    return az.from_dict(posterior=posterior_variables(trace), sample_stats=stats)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--pe", required=True, help="posterior samples: columns m1, q, z, wt, evt")
    ap.add_argument("--sel", required=True, help="found injections: columns m1, q, z, pdraw, ndraw")
    ap.add_argument("--out", default="trace_cosmo.npz")
    ap.add_argument("--nmcmc", type=int, default=NMCMC)
    ap.add_argument("--nchain", type=int, default=NCHAIN)
    ap.add_argument("--seed", type=int, default=RANDOM_SEED)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--python-driver", action="store_true", help="nuts.py instead of the library's C++ driver")
    a = ap.parse_args(argv)
    trace = fit(read_table(a.pe), read_table(a.sel), a.nmcmc, a.nmcmc, a.nchain, a.seed, a.device,
                native=not a.python_driver)
    np.savez_compressed(a.out, **trace)
    try:
        to_inference_data(trace).to_netcdf(os.path.splitext(a.out)[0] + ".nc")
    except ImportError:
        pass
    ess = trace["ess_bulk"][:14]
    print(json.dumps({"out": a.out, "nobs": int(trace["nobs"]), "nsamp": int(trace["nsamp"]), "nsel": int(trace["nsel"]),
                      "chains": a.nchain, "draws": a.nmcmc, "wall_s": round(float(trace["wall_s"]), 3),
                      "ess_min": float(ess.min()), "rhat_max": float(trace["rhat"][:14].max()),
                      "divergences": int(trace["stat_diverging"].sum()),
                      "posterior_mean": {n: round(float(trace["posterior"][:, :, i].mean()), 4)
                                         for i, n in enumerate(trace["site_names"])}}))
    return trace


if __name__ == "__main__":
    main()
