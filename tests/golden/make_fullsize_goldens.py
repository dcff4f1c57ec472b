#!/usr/bin/env python
"""Full-size golden vectors from the CPU oracle (oracle/abc_oracle.cpp) for the -m gpu parity tests.

    python tests/golden/make_fullsize_goldens.py C3            # N=250k, K=150, P=30 (~3 min, 6 GB)
    python tests/golden/make_fullsize_goldens.py T1M           # N=1M,   K=150, P=30 (~15 min, 25 GB)
    python tests/golden/make_fullsize_goldens.py C5 0.2        # N=200k, K=500, P=50 (~1 h, 25 GB; the full 1M cube is 100 GB)
    python tests/golden/make_fullsize_goldens.py C4rows        # 64 new rows of the 1M x 1M x 30 weight update (~1 min)

The oracle runs the reference's call sequence (src/AbcSmc.cpp:634-664, 1041-1066): particle_ranking_PLS -> truncate to
N_pp -> gather -> calculate_doubled_variance -> weight_predictive_prior against the previous predictive prior. What is kept
(tests/golden/fullsize_<name>.npz, a few hundred kB): the first N_pp entries of the order (bit-exact claim), component counts,
PRESS, the distances of the selected particles and of 4096 sampled ones, the sum of all distances, dv, weights, and the smallest
relative gap between neighbouring distances among the first N_pp+1 (how far the ranking is from a tie).
Inputs are regenerated from abcsmc_b200/synth.py (seeded splitmix64), so only outputs are stored.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def ranking_golden(name, scale):
    from abcsmc_b200 import synth
    import oracle
    cfg = synth.make_config(name, scale=scale)
    N, N_pp = cfg["N"], cfg["N_pp"]
    t0 = time.perf_counter()
    r = oracle.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    t_rank = time.perf_counter() - t0
    order = r["order"][:N_pp].astype(np.int64)
    sel = np.asfortranarray(cfg["params"][order, :])
    dv = oracle.calculate_doubled_variance(sel)
    t0 = time.perf_counter()
    w = oracle.weight_predictive_prior(np.ones(N_pp), sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"])
    t_w = time.perf_counter() - t0
    head = r["dist"][r["order"][:N_pp + 1].astype(np.int64)]
    gaps = np.diff(head) / head[1:]
    sidx = (np.arange(4096, dtype=np.int64) * 2654435761) % N
    out = dict(name=name, scale=scale, N=N, K=cfg["K"], P=cfg["P"], N_pp=N_pp, order_top=r["order"][:N_pp], ncomp=r["ncomp"],
               ncomp_used=r["ncomp_used"], press=r["press"], dist_top=head[:N_pp], dist_sample_idx=sidx, dist_sample=r["dist"][sidx],
               dist_sum=np.sum(r["dist"]), min_rel_gap_top=gaps.min(), dv=dv, w=w, oracle_seconds_rank=t_rank, oracle_seconds_weights=t_w)
    tag = name if scale == 1.0 else f"{name}_s{scale:g}"
    path = os.path.join(HERE, f"fullsize_{tag}.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: N={N} ncomp_used={r['ncomp_used']} min_rel_gap_top={gaps.min():.3e} rank {t_rank:.1f}s weights {t_w:.1f}s", flush=True)


def c4_rows_golden(n_rows=64):
    """64 new-particle rows (spread over the set) of the C4 update against all 1M old particles, un-normalised:
    w_i = numer_i / sum_j w_j prod_p phi(theta_ip - theta_jp; sqrt(dv_p)) (src/AbcUtil.cpp:556-581)."""
    from abcsmc_b200 import synth
    import oracle
    c = synth.CONFIGS["C4"]
    th_new, th_old, w_old, dv_old = synth.make_weight_case(c["N_new"], c["N_old"], c["P"], c["seed"])
    rows = (np.arange(n_rows, dtype=np.int64) * 15625 + 7) % c["N_new"]
    sub = np.asfortranarray(th_new[rows, :])
    t0 = time.perf_counter()
    w = oracle.weight_predictive_prior(np.ones(n_rows), sub, th_old, w_old, dv_old)     # L2-normalised over the 64 rows
    dt = time.perf_counter() - t0
    path = os.path.join(HERE, "fullsize_C4rows.npz")
    np.savez_compressed(path, rows=rows, w_normalised_over_rows=w, oracle_seconds=dt)
    print(f"{path}: {n_rows} rows x {c['N_old']} old particles in {dt:.1f}s", flush=True)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "C4rows":
        c4_rows_golden()
    else:
        ranking_golden(what, float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
