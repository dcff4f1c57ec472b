"""Where the reference's component selection is ill-posed: quantised metrics with few predictors. Hold-out rows that share a predictor row
give mathematically equal |D| = ||e_ref| - |e_alt|| with opposite signs whenever their residuals differ in sign, so the signed-rank sum
(pls.cpp:190-211: no tie correction, std::sort decides) turns on rounding-level differences between the two error columns, and p-values near
the 0.1 threshold flip the component count. Any two correct FP64 implementations (the reference on two Eigen builds included) can disagree
there. This script counts how often, on seeded cases: implementation A vs B = the CPU oracle vs the CUDA path (on a GPU box), or the
oracle vs the reference's own sources on the Eigen stand-in (where oracle/_ref exists: pass `ref`).
usage: python tools/quantised_fuzz.py [gpu|ref] [n_cases] [stdsort]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import oracle


def case(seed, quantised=True):
    rng = np.random.default_rng(seed)
    K = int(rng.integers(2, 5)); P = int(rng.integers(2, 6)); f = float(rng.choice([0.3, 0.5, 0.8]))
    N = int(rng.integers(max(int(np.ceil((K + 1) / f)) + 2, 12), 200))
    th = rng.uniform(size=(N, P))
    L = rng.normal(size=(P, K)) * np.linspace(1, 0.05, K)
    met = th @ L + 0.3 * np.tanh(th @ rng.normal(size=(P, K))) + 0.2 * rng.normal(size=(N, K))
    if quantised:
        met = np.round(met * 4) / 4
    return np.asfortranarray(met), np.asfortranarray(th), (0.5 * np.ones(P)) @ L, f


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "gpu"
    n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    if which == "gpu":
        from abcsmc_b200 import api
        if len(sys.argv) > 3 and sys.argv[3] == "stdsort":      # exact distance ties as std::sort leaves them (abcb200_set_tie_order 1)
            api.get_context(0).set_tie_order(api.TIES_STDSORT)
        other = lambda met, th, t, f: api.particle_ranking_PLS(met, th, t, f, return_info=True)
    else:
        import oracle.ref as ref
        other = lambda met, th, t, f: {"order": ref.particle_ranking_PLS(met, th, t, f)}
    for quantised in (True, False):
        ran = differ = count_differs = 0
        for seed in range(n_cases):
            met, th, t, f = case(seed, quantised)
            n_tr = int(round(met.shape[0] * f))
            if np.any(met.std(axis=0) == 0) or n_tr < met.shape[1] + 1:
                continue
            o = oracle.particle_ranking_PLS(met, th, t, f)
            try:
                g = other(met, th, t, f)
            except Exception as e:          # shapes the library refuses (fewer training rows than components)
                continue
            ran += 1
            same = np.array_equal(np.asarray(g["order"]).astype(np.int64), o["order"].astype(np.int64))
            differ += not same
            if "ncomp_used" in g:
                count_differs += int(g["ncomp_used"]) != int(o["ncomp_used"])
        print(f"{'quantised' if quantised else 'continuous'} metrics, oracle vs {which}: {ran} cases, {differ} orders differ"
              + (f", {count_differs} component counts differ" if which == "gpu" else ""))


if __name__ == "__main__":
    main()
