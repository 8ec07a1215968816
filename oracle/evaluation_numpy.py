"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's sample-quality metrics
(/root/reference python/utils/evaluation.py), float32 like the reference's jnp arrays.  Parity unpinned: the
reference cannot run here (no JAX); the formulas are short enough to be read off the cited lines.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
import numpy as np


def _f32(a):
    a = np.asarray(a, dtype=np.float32)
    return a[:, None] if a.ndim == 1 else a


def sqdist(x, y):
    """((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)  (evaluation.py:220, :283)"""
    x, y = _f32(x), _f32(y)
    return ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1, dtype=np.float32)


def pth_moment_rmse(x, y, p=2.0):
    """evaluation.py:33-38"""
    x, y = _f32(x), _f32(y)
    return float(np.linalg.norm(np.mean(x ** np.float32(p), axis=0) - np.mean(y ** np.float32(p), axis=0)))


def gaussian_kernel(x, y, gamma):
    """evaluation.py:220-222"""
    return np.exp(-np.float32(gamma) * sqdist(x, y))


def mmd2_unbiased(x, y, gamma=1.0):
    """evaluation.py:246-263"""
    n, m = len(x), len(y)
    Kxx, Kyy, Kxy = gaussian_kernel(x, x, gamma), gaussian_kernel(y, y, gamma), gaussian_kernel(x, y, gamma)
    np.fill_diagonal(Kxx, 0)
    np.fill_diagonal(Kyy, 0)
    return float(Kxx.sum(dtype=np.float64) / (n * (n - 1)) + Kyy.sum(dtype=np.float64) / (m * (m - 1))
                 - 2 * Kxy.sum(dtype=np.float64) / (n * m))


def median_sqdist(y):
    """jnp.median(((y[:, None, :] - y[None, :, :]) ** 2).sum(-1))  (evaluation.py:283)"""
    return float(np.median(sqdist(y, y)))


def mmd_heuristic(x, y):
    """evaluation.py:279-294"""
    n, m = len(x), len(y)
    gamma = 4.0 / median_sqdist(y)
    Kxx, Kyy, Kxy = gaussian_kernel(x, x, gamma), gaussian_kernel(y, y, gamma), gaussian_kernel(x, y, gamma)
    mmd2 = Kxx.sum(dtype=np.float64) / n**2 + Kyy.sum(dtype=np.float64) / m**2 - 2 * Kxy.sum(dtype=np.float64) / (n * m)
    return float(np.sqrt(mmd2))


def distance_matrix(u, v, p=2.0):
    """scipy.spatial.distance_matrix(u, v, p) (evaluation.py:58)"""
    from scipy.spatial.distance import cdist  # (scipy.spatial.distance_matrix itself is deprecated since SciPy 1.18: same Minkowski-p matrix)

    return cdist(np.asarray(u, np.float64), np.asarray(v, np.float64), metric="minkowski", p=p)


def wasserstein_dist11_p(u, v, ord=2.0):
    """evaluation.py:58-62"""
    from scipy.optimize import linear_sum_assignment

    cm = distance_matrix(u, v, ord)
    r, c = linear_sum_assignment(cm)
    return float(cm[r, c].mean())


def wasserstein_1d(mu, nu, p=1.0):
    """evaluation.py:150-154"""
    diff = np.abs(np.sort(mu, axis=-1) - np.sort(nu, axis=-1))
    return np.mean(diff ** p, axis=-1) ** (1.0 / p)


def max_sliced_wasserstein_given_directions(mu, nu, directions, p=1.0):
    """evaluation.py:185-198 with the directions supplied (the reference draws them with jax.random.normal)."""
    directions = directions / np.linalg.norm(directions, axis=1, keepdims=True)
    return float(max(wasserstein_1d(mu @ dd, nu @ dd, p) for dd in directions))


def wasserstein_sinkhorn(u, v, epsilon=None, threshold=1e-3, max_iterations=2000, inner_iterations=10, return_info=False):
    """python/utils/evaluation.py:69-97 restated without OTT-JAX (a third-party dependency of the reference, not installable
    here): log-domain Sinkhorn in float64 for uniform weights and the Euclidean cost, with the defaults of
    `ott.solvers.linear.sinkhorn.Sinkhorn` as recalled (zero initial potentials; one iteration = g-update then f-update; the L1
    error of the row marginal is checked every `inner_iterations`; epsilon = 0.05 x mean cost when None) and OTT's
    `ent_reg_cost` of a balanced problem, sum a f + sum b g + eps (1 - sum P).  PARITY STATUS: unpinned against OTT."""
    from scipy.special import logsumexp

    Cm = distance_matrix(u, v, 2.0).astype(np.float32).astype(np.float64)  # the GPU solves the float32 matrix
    n, m = Cm.shape
    eps = 0.05 * Cm.mean() if epsilon is None else float(epsilon)
    la, lb = -np.log(n), -np.log(m)
    f, g = np.zeros(n), np.zeros(m)
    it, err, conv = 0, np.inf, False
    while it < max_iterations and not conv:
        for k in range(inner_iterations):
            if it >= max_iterations:
                break
            g = -eps * (logsumexp((f[:, None] - Cm) / eps, axis=0) + la)
            f_new = -eps * (logsumexp((g[None, :] - Cm) / eps, axis=1) + lb)
            if k == inner_iterations - 1 or it == max_iterations - 1:
                err = float(np.sum(np.abs(np.exp((f - f_new) / eps) - 1.0)) / n)
            f = f_new
            it += 1
        conv = err < threshold
    P_sum = np.exp((f[:, None] + g[None, :] - Cm) / eps).sum() / (n * m)
    val = f.mean() + g.mean() + eps * (1.0 - P_sum)
    if return_info:
        return float(val), dict(iterations=it, error=err, converged=conv, epsilon=eps, f=f, g=g)
    return float(val)
