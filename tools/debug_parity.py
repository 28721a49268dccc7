"""Debug helper (not a test): print CUDA-vs-golden differences."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bumpcosmology_b200.likelihood import Hyperlikelihood
from oracle import bump_oracle as bo
np.set_printoptions(linewidth=200, precision=4)
for name in ("tiny", "small"):
    g = np.load(f"tests/golden/pop_cosmo_{name}.npz")
    data = (g["m1s_det"], g["qs"], g["dls"], g["pdraw"], g["m1s_det_sel"], g["qs_sel"], g["dls_sel"], g["pdraw_sel"], float(g["Ndraw"]))
    like = Hyperlikelihood(*data)
    print(name, like.plan())
    for k, th in enumerate(g["thetas"]):
        r = like(th)
        t = like.tables()
        def rel(a, b, floor=1.0):
            a, b = np.asarray(a, float), np.asarray(b, float)
            return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
        print(k, "loglike", r.loglike, g["ref_loglike"][k], "rel", rel(r.loglike, g["ref_loglike"][k]),
              "| log_mu", rel(r.log_mu_sel, g["ref_log_mu_sel"][k]), "| neff_sel", rel(r.neff_sel, g["ref_neff_sel"][k]),
              "| neff", rel(r.neff, g["ref_neff"][k]))
        for key in ("zinterp", "dlinterp", "ddlinterp", "dvcinterp", "log_dN_grid"):
            print("   tab", key, rel(t[key], g["tab_" + key][k], 1e-3), end="")
        print()
        gs = bo.grad_sites_from_theta(r.dloglike[:14], th); gm = bo.grad_sites_from_theta(r.dlog_mu_sel[:14], th)
        sc = max(1.0, np.max(np.abs(g["ref_dloglike_dsite"][k])))
        print("   dloglike relerr", np.abs(gs - g["ref_dloglike_dsite"][k]) / sc)
        print("   dlogmu   relerr", np.abs(gm - g["ref_dlog_mu_sel_dsite"][k]) / np.maximum(1, np.abs(g["ref_dlog_mu_sel_dsite"][k])))
        if k == 0:
            print("   dloglike", gs); print("   ref     ", g["ref_dloglike_dsite"][k])
            print("   dlogmu", gm); print("   ref   ", g["ref_dlog_mu_sel_dsite"][k])
    like.close()
