"""The CUDA path on the reference's own particle sets against the outputs of the reference's own code (tests/golden/ref_realdata.npz):
the body of tests/test_gpu_golden.py::test_full_step_on_the_reference_s_own_data without pytest / torch (starts in ~2 s)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abcsmc_b200 import api

g = np.load(os.path.join(ROOT, "tests", "golden", "ref_realdata.npz"))
ok_all = True
for tag in ("dengue_sqlite", "dengue_pp250"):
    met, par, target = np.asfortranarray(g[f"{tag}_met"]), np.asfortranarray(g[f"{tag}_par"]), g[f"{tag}_target"]
    r = api.particle_ranking_PLS(met, par, target, 0.5, top_n=0, return_info=True)
    order = r["order"].astype(np.int64)
    n_pp = met.shape[0] // 10
    th_new, th_old = np.asfortranarray(par[order[:n_pp]]), np.asfortranarray(par[order[n_pp:2 * n_pp]])
    numer = np.full(n_pp, np.prod(1.0 / (g[f"{tag}_prior_hi"] - g[f"{tag}_prior_lo"])))
    w = api.weight_predictive_prior(numer, th_new, th_old, np.full(n_pp, 1.0 / n_pp), g[f"{tag}_dv_next"])
    checks = {
        "ncomp": list(r["ncomp"]) == [int(v) for v in g[f"{tag}_ncomp"]],
        "order": bool(np.array_equal(order, g[f"{tag}_order"].astype(np.int64))),
        "dist": float(np.max(np.abs(r["dist"] - g[f"{tag}_dist"]) / np.abs(g[f"{tag}_dist"]))),
        "dv": float(np.max(np.abs(api.calculate_doubled_variance(th_new) - g[f"{tag}_dv"]) / g[f"{tag}_dv"])),
        "w": float(np.max(np.abs(w - g[f"{tag}_w_vs_next"]) / g[f"{tag}_w_vs_next"])),
    }
    good = checks["ncomp"] and checks["order"] and checks["dist"] < 1e-10 and checks["dv"] < 1e-10 and checks["w"] < 1e-10
    ok_all &= good
    print(tag, "PASS" if good else "FAIL", checks, "order mismatches:", int(np.sum(order != g[f"{tag}_order"].astype(np.int64))), flush=True)
sys.exit(0 if ok_all else 1)
