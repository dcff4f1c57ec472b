"""Independent numpy/scipy formulation of the hot path, used ONLY to pin the C++ oracle
("validate the restatement ... against an independent numpy/scipy implementation").

It deliberately uses different building blocks from oracle/abc_oracle.cpp: numpy.linalg.eigh instead of
Jacobi, BLAS matmuls instead of explicit loops, scipy.stats.rankdata for the signed-rank statistic,
vectorised log-space Gaussian kernels for the weights.
"""
import numpy as np
from scipy.stats import rankdata


def zscore_cols(X):
    mu = X.mean(axis=0)
    sd = np.sqrt(((X - mu) ** 2).sum(axis=0) / (X.shape[0] - 1))
    return (X - mu) / sd, mu, sd


def kernel_pls(X, Y, A, gram=False):
    """Dayal & MacGregor (1997) modified kernel PLS #1 (gram=False) / #2 (gram=True)."""
    K, M = X.shape[1], Y.shape[1]
    XY = X.T @ Y
    XX = X.T @ X if gram else None
    W = np.zeros((K, A)); P = np.zeros((K, A)); R = np.zeros((K, A)); Q = np.zeros((M, A))
    T = np.zeros((X.shape[0], A))
    for i in range(A):
        if M == 1:
            w = XY[:, 0].copy()
        else:
            ev, evec = np.linalg.eigh(XY.T @ XY)
            w = XY @ evec[:, np.argmax(np.abs(ev))]
        w = w / np.sqrt(w @ w)
        r = w - R[:, :i] @ (P[:, :i].T @ w)
        if gram:
            xr = XX @ r; tt = r @ xr; p = xr / tt
        else:
            t = X @ r; tt = t @ t; p = X.T @ t / tt; T[:, i] = t
        q = XY.T @ r / tt
        XY = XY - np.outer(p, q) * tt
        W[:, i], P[:, i], R[:, i], Q[:, i] = w, p, r, q
    return dict(W=W, P=P, R=R, Q=Q, T=T)


def normalcdf(z):
    c = (0.196854, 0.115194, 0.000344, 0.019527)
    zs = abs(z)
    p = 0.5 / (1 + c[0] * zs + c[1] * zs ** 2 + c[2] * zs ** 3 + c[3] * zs ** 4) ** 4
    return p if z < 0 else 1.0 - p


def wilcoxon(e1, e2):
    d = np.abs(e1) - np.abs(e2)
    ranks = rankdata(np.abs(d), method="ordinal")
    n = d.size
    dd = float(np.sum(ranks * np.sign(d)))
    t = n * (n + 1) / 2.0
    v = (t - dd) / 2.0
    sv = np.sqrt(n * (n + 1) * (2 * n + 1) / 24.0)
    return 1.0 - normalcdf((v - t / 2.0) / sv)


def holdout_errors(m, Xte, Yte):
    """Error cube [y][n, c] via prefix sums over components (not the reference's per-c GEMM)."""
    Tte = Xte @ m["R"]
    A = m["R"].shape[1]
    E = np.empty((Yte.shape[1], Xte.shape[0], A))
    cur = Yte.copy()
    for c in range(A):
        cur = cur - np.outer(Tte[:, c], m["Q"][:, c])
        E[:, :, c] = cur.T
    return E


def optimal_num_components(E, alpha=0.1):
    press = (E ** 2).sum(axis=1)
    out = np.zeros(E.shape[0], dtype=np.int64)
    for y in range(E.shape[0]):
        ref = int(np.argmin(press[y]))
        best = ref
        for alt in range(ref):
            if wilcoxon(E[y, :, ref], E[y, :, alt]) > alpha:
                best = alt
                break
        out[y] = best + 1
    return out, press


def rank_pls(met, par, target, f=0.5):
    zm, mu, sd = zscore_cols(met)
    zp, _, _ = zscore_cols(par)
    obs = (target - mu) / sd
    ntr = int(np.floor(met.shape[0] * f + 0.5))
    m = kernel_pls(zm[:ntr], zp[:ntr], met.shape[1])
    E = holdout_errors(m, zm[ntr:], zp[ntr:])
    nc, press = optimal_num_components(E)
    c = int(nc.max())
    d = np.sqrt((((zm - obs) @ m["R"][:, :c]) ** 2).sum(axis=1))
    return dict(order=np.argsort(d, kind="stable"), dist=d, ncomp=nc, ncomp_used=c, press=press)


def doubled_variance(params):
    return 2.0 * params.var(axis=0, ddof=1)


def weights(numer, th_new, th_old, w_old, dv_old):
    """Log-space evaluation of src/AbcUtil.cpp:547-586 (valid for dv_old > 0)."""
    sig = np.sqrt(dv_old)
    out = np.empty(th_new.shape[0])
    logc = -np.sum(np.log(np.sqrt(2 * np.pi) * sig))
    for i0 in range(0, th_new.shape[0], 256):
        u = (th_new[i0:i0 + 256, None, :] - th_old[None, :, :]) / sig
        dens = np.exp(logc - 0.5 * (u ** 2).sum(axis=2)) @ w_old
        out[i0:i0 + 256] = numer[i0:i0 + 256] / dens
    return out / np.sqrt(np.sum(out ** 2))
