"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE): exact posterior moments of the reference's diamonds model.

Model: python/scripts/run_diamonds_lr_decay.py:24-40 -- b ~ N(0,1)^Kc, Intercept ~ StudentT(3, 8, 10),
sigma ~ half-StudentT(3, 0, 10), Y ~ N(Intercept + Xc b, sigma).  The likelihood is Gaussian-linear, so it depends
on the data only through the sufficient statistics (N, G = X1^T X1, h = X1^T Y, yy = Y^T Y), X1 = [1 | Xc], and

    beta | sigma ~ N( (G + sigma^2 P)^-1 h,  sigma^2 (G + sigma^2 P)^-1 ),        P = diag(0, 1, ..., 1)

(the Student-t prior of the Intercept, scale 10, is flat on the posterior's scale of ~2e-3 and is kept as a factor of the
sigma-marginal's weight only through its value at the conditional mean), which leaves a ONE-dimensional marginal
over s = log sigma that is integrated by quadrature.  Used to (i) recover a diamonds-equivalent data set from the
posteriordb reference draws shipped in the reference repo (tests/golden/make_golden.py) and (ii) give the samplers'
statistical tests exact targets in the flat order [Intercept, b, log sigma] (python/scripts/eval_diamonds.py:78-87).
"""
from __future__ import annotations

import numpy as np


def sufficient_stats(X, Y):
    X = np.asarray(X, np.float64)
    Y = np.asarray(Y, np.float64)
    Xc = X[:, 1:] - X[:, 1:].mean(0)
    X1 = np.column_stack([np.ones(len(Y)), Xc])
    return dict(n=len(Y), G=X1.T @ X1, h=X1.T @ Y, yy=float(Y @ Y))


def _t3_logpdf(x, loc, scale):
    from math import lgamma, log, pi

    z = (x - loc) / scale
    return -2.0 * np.log1p(z * z / 3.0) - (log(scale) + 0.5 * log(3.0) + 0.5 * log(pi) + lgamma(1.5) - lgamma(2.0))


def posterior_moments(stats, n_grid=801, width=9.0):
    """Exact (to quadrature error) posterior mean and covariance of q = [Intercept, b, log sigma].
    Returns dict(mean[d], cov[d,d], s_grid, s_weights)."""
    n, G, h, yy = stats["n"], np.asarray(stats["G"], np.float64), np.asarray(stats["h"], np.float64), float(stats["yy"])
    k = G.shape[0]
    P = np.eye(k)
    P[0, 0] = 0.0
    beta_ols = np.linalg.solve(G, h)
    rss0 = yy - h @ beta_ols
    s_hat = 0.5 * np.log(rss0 / (n - k))
    sd_s = 1.0 / np.sqrt(2.0 * (n - k))
    s = s_hat + np.linspace(-width, width, n_grid) * sd_s
    logw = np.empty(n_grid)
    means = np.empty((n_grid, k))
    covs = np.empty((n_grid, k, k))
    for i, si in enumerate(s):
        v = np.exp(2.0 * si)
        A = G / v + P
        L = np.linalg.cholesky(A)
        m = np.linalg.solve(A, h / v)
        means[i] = m
        covs[i] = np.linalg.inv(A)
        quad = yy / v - (h / v) @ m  # = min_beta [RSS(beta)/v + b^T b]
        logdet = 2.0 * np.log(np.diag(L)).sum()
        sigma = np.exp(si)
        logw[i] = (-n * si - 0.5 * quad - 0.5 * logdet            # likelihood x N(0,1) prior of b, beta integrated out
                   + _t3_logpdf(sigma, 0.0, 10.0) + si              # half-t prior of sigma + log-Jacobian of s = log sigma
                   + _t3_logpdf(m[0], 8.0, 10.0))                   # Intercept prior at the conditional mean
    w = np.exp(logw - logw.max())
    w /= w.sum()
    mean_b = w @ means
    cov_b = np.einsum("i,ijk->jk", w, covs) + np.einsum("i,ij,ik->jk", w, means - mean_b, means - mean_b)
    mean_s = w @ s
    var_s = w @ (s - mean_s) ** 2
    cross = np.einsum("i,ij->j", w * (s - mean_s), means - mean_b)
    d = k + 1
    mean = np.concatenate([mean_b, [mean_s]])
    cov = np.zeros((d, d))
    cov[:k, :k] = cov_b
    cov[k, k] = var_s
    cov[:k, k] = cov[k, :k] = cross
    return dict(mean=mean, cov=cov, s_grid=s, s_weights=w, e_sigma2=float(w @ np.exp(2 * s)))


def recover_stats_from_draws(draws, n=5000, iters=6):
    """Sufficient statistics (G, h, yy) of a diamonds data set whose posterior has the moments of `draws`
    ([S, 26] in the order [Intercept, b[24], log sigma]).  First-order estimate from the conditional-Gaussian structure
    (G = E[sigma^2] (Cov^-1 - P), h = (G + E[sigma^2] P) mean, RSS0 = (n - k - 3) E[sigma^2]) with the centred-design
    structure imposed (G_00 = n, G_0j = 0), then a few fixed-point corrections against `posterior_moments` so that the
    exact posterior of the recovered statistics reproduces the draws' mean / covariance of beta and E[sigma^2]."""
    x = np.asarray(draws, np.float64)
    k = x.shape[1] - 1
    beta, s = x[:, :k], x[:, k]
    m_t = beta.mean(0)
    C_t = np.cov(beta.T)
    C_t[0, 1:] = C_t[1:, 0] = 0.0            # centred predictors: Intercept is orthogonal to b in the likelihood
    e2_t = float(np.mean(np.exp(2.0 * s)))
    C_t[0, 0] = e2_t / n                      # G_00 = n exactly
    P = np.eye(k)
    P[0, 0] = 0.0
    A_eff, m_eff, e2_eff = np.linalg.inv(C_t), m_t.copy(), e2_t
    for it in range(iters + 1):
        G = e2_eff * (A_eff - P)
        G[0, 1:] = G[1:, 0] = 0.0
        G[0, 0] = n
        G = 0.5 * (G + G.T)
        h = (G + e2_eff * P) @ m_eff
        rss0 = (n - k - 3) * e2_eff
        yy = rss0 + h @ np.linalg.solve(G, h)
        stats = dict(n=n, G=G, h=h, yy=float(yy))
        if it == iters:
            break
        pm = posterior_moments(stats)
        k_ = k
        # multiplicative / additive corrections towards the targets
        C_now = pm["cov"][:k_, :k_]
        A_eff = A_eff + (np.linalg.inv(C_t) - np.linalg.inv(C_now))
        m_eff = m_eff + (m_t - pm["mean"][:k_])
        e2_eff = e2_eff * e2_t / pm["e_sigma2"]
    return stats


def dataset_from_stats(stats, seed=0, col_means=None):
    """A data set (X [n, k] with column 0 = ones, Y [n]) with EXACTLY the given sufficient statistics (up to round-off):
    Xc = Q R with Q orthonormal and orthogonal to the ones vector, R^T R = G_bb; Y = ybar + Xc b_ols + r, r orthogonal
    to [1, Q] with |r|^2 = RSS0.  `col_means` are added back to the predictors (the model centres them itself)."""
    n, G, h, yy = stats["n"], np.asarray(stats["G"], np.float64), np.asarray(stats["h"], np.float64), float(stats["yy"])
    k = G.shape[0]
    rng = np.random.default_rng(seed)
    M = np.column_stack([np.ones(n), rng.normal(size=(n, k))])
    Q, _ = np.linalg.qr(M)                      # Q[:,0] ~ ones; Q[:,1:k] ~ orthonormal, orthogonal to ones; Q[:,k] residual direction
    R = np.linalg.cholesky(G[1:, 1:]).T
    Xc = Q[:, 1:k] @ R
    beta_ols = np.linalg.solve(G, h)
    rss0 = yy - h @ beta_ols
    Y = beta_ols[0] + Xc @ beta_ols[1:] + np.sqrt(max(rss0, 0.0)) * Q[:, k]
    if col_means is None:
        col_means = np.linspace(0.2, 1.5, k - 1)
    X = np.column_stack([np.ones(n), Xc + np.asarray(col_means)[None, :]])
    return dict(X=X, Y=Y)
