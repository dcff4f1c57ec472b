"""Seeded synthetic particle sets for the parity tests and bench.py.

Everything is derived from splitmix64 -> uniform -> Box-Muller, written out explicitly so the
buffers do not depend on numpy's or libstdc++'s distribution implementations.

Shapes follow BASELINE.json: C2 = 100k particles x (10 params, 20 metrics), top-N 1k;
C3 = 250k x (30, 150), top-N 5k; C4 = weight update 1M x 1M x 30; C5 = 1M x (50, 500); T1M = the north-star target 1M x (30, 150), top-N 10k.
"""
import numpy as np

_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)

CONFIGS = {
    "C2": dict(N=100_000, P=10, K=20, N_pp=1_000, seed=0xABC50002),
    "C3": dict(N=250_000, P=30, K=150, N_pp=5_000, seed=0xABC50003),
    "C4": dict(N_new=1_000_000, N_old=1_000_000, P=30, seed=0xABC50004),
    "C5": dict(N=1_000_000, P=50, K=500, N_pp=10_000, seed=0xABC50005),
    # BASELINE.json north_star's target shape: one 1M-particle SMC set with 30 params and 150 metrics
    "T1M": dict(N=1_000_000, P=30, K=150, N_pp=10_000, seed=0xABC50006),
}


def splitmix64(seed, n, stream=0):
    """n 64-bit outputs of splitmix64 started at `seed` (+ stream offset), vectorised."""
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + np.uint64(stream) * np.uint64(0xD1342543DE82EF95)
        z = base + np.arange(1, n + 1, dtype=np.uint64) * _GAMMA
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def uniform(seed, n, stream=0):
    """U[0,1) doubles with 53 random bits."""
    return (splitmix64(seed, n, stream) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(seed, n, stream=0):
    """Standard normals by Box-Muller (cosine branch) from two independent uniform streams."""
    u1 = ((splitmix64(seed, n, 2 * stream + 1000) >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)
    u2 = uniform(seed, n, 2 * stream + 1001)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def make_set(N, P, K, seed, noise=0.2):
    """One SMC set: theta ~ U(0,1)^P, metrics = theta L + 0.3 tanh(theta L2) + noise*eps with decaying
    factor strengths; target = noiseless metrics of theta* = 0.5. Returns Fortran-order arrays
    (params N x P, metrics N x K, target K)."""
    theta = np.asfortranarray(uniform(seed, N * P, 1).reshape(P, N).T)
    L = normal(seed, P * K, 2).reshape(P, K) * np.linspace(1.0, 0.05, K)[None, :]
    L2 = normal(seed, P * K, 3).reshape(P, K)
    eps = normal(seed, N * K, 4).reshape(K, N).T
    met = theta @ L + 0.3 * np.tanh(theta @ L2) + noise * eps
    tstar = np.full((1, P), 0.5)
    target = (tstar @ L + 0.3 * np.tanh(tstar @ L2)).ravel()
    return theta, np.asfortranarray(met), np.ascontiguousarray(target)


def make_prev_posterior(N_pp, P, seed):
    """A previous set's predictive prior: theta_old ~ N(0.5, 0.1^2) clipped to [0,1], positive L2-normalised
    weights, dv_old = 2 * sample variance (what calculate_doubled_variance would give)."""
    th = np.clip(0.5 + 0.1 * normal(seed, N_pp * P, 7).reshape(P, N_pp).T, 0.0, 1.0)
    w = 0.5 + uniform(seed, N_pp, 8)
    w = w / np.sqrt(np.sum(w * w))
    dv = 2.0 * th.var(axis=0, ddof=1)
    return np.asfortranarray(th), w, dv


def make_weight_case(N_new, N_old, P, seed):
    """Weight-update inputs (C4 shape): theta_new = resampled theta_old + N(0, dv_old), clipped to [0,1]."""
    th_old, w_old, dv_old = make_prev_posterior(N_old, P, seed)
    pick = (uniform(seed, N_new, 9) * N_old).astype(np.int64)
    th_new = th_old[pick, :] + normal(seed, N_new * P, 10).reshape(P, N_new).T * np.sqrt(dv_old)[None, :]
    th_new = np.asfortranarray(np.clip(th_new, 0.0, 1.0))
    return th_new, th_old, w_old, dv_old


def make_config(name, scale=1.0, seed_offset=0):
    """Inputs for one bench/parity step at a BASELINE.json config: the new set (params, metrics, target) and
    the previous set's predictive prior (theta_old, w_old, dv_old). `scale` shrinks N and N_pp for CPU tests;
    `seed_offset` gives an independent set of the same shape (bench replicas)."""
    c = CONFIGS[name]
    N = max(int(c["N"] * scale), 8 * c["K"])
    N_pp = max(int(c["N_pp"] * scale), 16)
    par, met, target = make_set(N, c["P"], c["K"], c["seed"] + seed_offset)
    th_old, w_old, dv_old = make_prev_posterior(N_pp, c["P"], c["seed"] + 1 + seed_offset)
    return dict(name=name, N=N, P=c["P"], K=c["K"], N_pp=N_pp, params=par, metrics=met, target=target,
                theta_old=th_old, w_old=w_old, dv_old=dv_old)
